// Seed lookup + per-read transcript vote (the reference's "sparse chaining", src/sparse_chaining.cpp:29-115).
//
// The sketch kernel leaves the selected hashes of a batch as one dense array per k-index.  From there:
//   lookup_kernel      thread per hash: the hash's list descriptor (see IndexTable): one 32-byte bitmap sector says
//                      whether it is a key and which, one word of the descriptor array follows for a hit.  It runs
//                      one k-index at a time over the whole batch, so that k's 31 + 25 MB are the only randomly
//                      accessed data in flight and stay in L2.
//   vote_bits_kernel   short reads, thread per read: bit-sliced counters over a 64-transcript window.
//   vote_long_kernel   warp per read on 64 shared-memory counters per k-index: long (multi-item) reads, and the
//                      short reads the bit-sliced kernel hands on.
//   vote_kernel        warp per read with a general transcript hash table in shared memory: whatever is left
//                      (lists outside the window that matter, three-range lists, more than 4 k values).
//   vote_overflow_kernel  the same on a per-worker global-memory scratch sized for every transcript of the index.
// Every tier produces, per read: per-k maximum over the transcripts seen, the double-precision
// `count < fraction*max` filter for every k, integer score = sum of counts, candidates ordered by (score desc,
// transcript asc), appended to the candidate store (its free tail; see sq_engine.cu).
#include <algorithm>

#include "sq_common.cuh"

namespace sq {

// the read's class sort key and list fingerprint (see ListHash) leave with its list: no later pass over the store
__device__ __forceinline__ void store_read_key(const VoteParams& P, uint32_t r, const ListHash& lh, uint32_t top) {
  const uint64_t g = P.read_base + r;
  P.rkey[g] = lh.key(top, P.key_hash_bits, g);
  reinterpret_cast<ulonglong2*>(P.rfp)[g] = make_ulonglong2(lh.h, lh.g);
}


size_t vote_smem_bytes(uint32_t nk);

static constexpr int kVoteWarps = 8;
static constexpr uint32_t kTabLog2 = 8;      // 256 transcript slots per warp
static constexpr uint32_t kTabMaxFill = 192;
static constexpr uint32_t kSetLog2 = 10;     // 1024 dedup-set slots per warp (aliased with the candidate buffer)
static constexpr uint32_t kSetMaxFill = 768;
static constexpr uint32_t kHashMul = 0x9E3779B1u;  // slot hashing of the per-read tables

struct Scratch {
  uint32_t* tkeys;  // transcript id per slot, SQ_EMPTY when free
  uint32_t* tcnt;   // [slot][nk] votes
  uint32_t* tlist;  // occupied slots in insertion order
  uint32_t* dset;   // per-(read,k) set of hashes already looked up
  unsigned long long* cand;  // sort buffer: (0x7FFFFFFF-score)<<32 | tid
  uint32_t* ctr;    // [0]=occupied slots, [1]=overflow flag, [2]=set fill, [3]=0xFFFFFFFF seen
  uint32_t tab_log2, tab_maxfill, set_log2, set_maxfill;
};

__device__ __forceinline__ bool set_insert(const Scratch& S, uint32_t h) {
  if (h == SQ_EMPTY) return atomicExch(&S.ctr[3], 1u) == 0u;
  const uint32_t mask = (1u << S.set_log2) - 1;
  uint32_t s = (h * kHashMul) >> (32 - S.set_log2);
  for (uint32_t tries = 0; tries <= mask; ++tries) {
    const uint32_t old = atomicCAS(&S.dset[s], SQ_EMPTY, h);
    if (old == SQ_EMPTY) {
      if (atomicAdd(&S.ctr[2], 1u) >= S.set_maxfill) S.ctr[1] = 1;
      return true;
    }
    if (old == h) return false;
    s = (s + 1) & mask;
  }
  S.ctr[1] = 1;
  return false;
}

__device__ __forceinline__ void table_vote(const Scratch& S, uint32_t t, uint32_t ki, uint32_t nk) {
  const uint32_t mask = (1u << S.tab_log2) - 1;
  uint32_t s = (t * kHashMul) >> (32 - S.tab_log2);
  for (uint32_t tries = 0; tries <= mask; ++tries) {
    const uint32_t old = atomicCAS(&S.tkeys[s], SQ_EMPTY, t);
    if (old == SQ_EMPTY) {
      const uint32_t n = atomicAdd(&S.ctr[0], 1u);
      if (n < S.tab_maxfill) S.tlist[n] = s; else S.ctr[1] = 1;
    }
    if (old == SQ_EMPTY || old == t) {
      atomicAdd(&S.tcnt[s * nk + ki], 1u);
      return;
    }
    s = (s + 1) & mask;
  }
  S.ctr[1] = 1;
}

// one 32-byte sector in a single 256-bit load (sm_100), not allocated in L1: a probe has no reuse there
__device__ __forceinline__ void ld_sector(const uint4* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}

// ------------------------------------------------------------------ seed lookup (sparse_chaining.cpp:61-63)
// pay[i] = descriptor of hsel[i] for the n selected hashes of k-index ki: the hash's bitmap sector says whether it
// is a key and which one (rank), the rank indexes the descriptor array.  Four independent probes per thread, no
// loops, no divergence beyond hit / miss.
static constexpr int kLookupBlock = 256;

__global__ void __launch_bounds__(kLookupBlock) lookup_kernel(const uint32_t* __restrict__ hsel, uint32_t* __restrict__ pay,
                                                              const uint32_t* __restrict__ n_ptr, const IndexTable tb,
                                                              unsigned long long* __restrict__ work) {
  const uint32_t n = *n_ptr;
  uint32_t hits = 0;
  for (uint32_t i0 = (blockIdx.x * kLookupBlock + threadIdx.x) * 4; i0 < n; i0 += gridDim.x * kLookupBlock * 4) {
    const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(hsel + i0));  // the arrays are padded to a multiple of 4
    const uint32_t hh[4] = {h4.x, h4.y, h4.z, h4.w};
    uint32_t w[4][8], rem[4];
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t sec = hh[u] / SQ_BMAP_BITS;
      rem[u] = hh[u] - sec * SQ_BMAP_BITS;
      in[u] = sec < tb.n_sectors && i0 + u < n;
      if (in[u]) ld_sector(tb.bmap + 2 * (size_t)sec, w[u]);
    }
    uint32_t d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      d[u] = SQ_EMPTY;
      if (in[u]) {
        const uint32_t wi = (rem[u] >> 5) + 1, bit = rem[u] & 31;
        uint32_t rank = w[u][0], cur = 0;
#pragma unroll
        for (uint32_t j = 1; j < 8; ++j) {
          rank += j < wi ? (uint32_t)__popc(w[u][j]) : 0u;
          cur = j == wi ? w[u][j] : cur;
        }
        if ((cur >> bit) & 1u) d[u] = __ldg(tb.desc + rank + __popc(cur & ((1u << bit) - 1u)));
      }
      hits += d[u] != SQ_EMPTY ? 1u : 0u;
    }
    *reinterpret_cast<uint4*>(pay + i0) = make_uint4(d[0], d[1], d[2], d[3]);
  }
  if (work) {
    __shared__ uint32_t s_hits;
    if (threadIdx.x == 0) s_hits = 0;
    __syncthreads();
    hits = __reduce_add_sync(0xFFFFFFFFu, hits);
    if (lane_id() == 0 && hits) atomicAdd(&s_hits, hits);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (s_hits) atomicAdd(work + 1, (unsigned long long)s_hits);
      if (blockIdx.x == 0) atomicAdd(work + 0, (unsigned long long)n);
    }
  }
}

// ---- list descriptors
__device__ __forceinline__ bool desc_inline(uint32_t d) { return (d >> 31) == 0; }
__device__ __forceinline__ uint32_t desc_base(const IndexTable& tb, uint32_t d) { return d & ((1u << tb.tbits) - 1u); }
// 64-bit membership mask of an inline list relative to its base (bit 0 = the base itself)
__device__ __forceinline__ unsigned long long desc_mask(const IndexTable& tb, uint32_t d) {
  return ((unsigned long long)(d >> tb.tbits) << 1) | 1ull;
}
__device__ __forceinline__ uint4 ld_hdr(const uint4* p) {  // one 16-byte list header, read-only path
  return __ldg(p);
}

__device__ __forceinline__ uint32_t items_of(uint32_t L) { return L == 0 ? 1u : (L + SQ_CHUNK - 1) / SQ_CHUNK; }

// warp-cooperative bitonic sort of cand[0..m), m a power of two >= 32
__device__ void warp_bitonic(unsigned long long* a, uint32_t m) {
  const uint32_t lane = lane_id();
  for (uint32_t k = 2; k <= m; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = lane; i < m; i += 32) {
        const uint32_t l = i ^ j;
        if (l > i) {
          const unsigned long long x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) { a[i] = y; a[l] = x; }
        }
      }
      __syncwarp();
    }
}

// Vote for read r.  Returns 0 = done, 1 = scratch overflow (nothing emitted; scratch is clean again).
__device__ int vote_read(const VoteParams& P, const Scratch& S, uint32_t r, uint32_t (&work)[3]) {
  const uint32_t lane = lane_id();
  const uint32_t nk = P.nk;
  const uint32_t item0 = P.item_start[r];
  const uint32_t n_it = P.item_start[r + 1] - item0;
  if (lane == 0) { S.ctr[0] = 0; S.ctr[1] = 0; }
  __syncwarp();

  for (uint32_t ki = 0; ki < nk; ++ki) {
    const IndexTable& tb = P.tab[ki];
    if (!tb.present) continue;
    // the sketch kernel removed the repeats inside an item: a hash can only come twice in a read of several
    // items (or in an item flagged SQ_CNT_RAW); then the hits go through a per-read set
    bool use_set = n_it > 1;
    if (lane == 0) { S.ctr[2] = 0; S.ctr[3] = 0; }
    __syncwarp();
    bool set_used = false;
    for (uint32_t g = 0; g < n_it; g += 32) {
      const uint32_t it = g + lane;
      const uint32_t craw = it < n_it ? P.cnt[(uint64_t)ki * P.n_items_ub + item0 + it] : 0u;
      const uint32_t c = craw & SQ_CNT_MASK;
      const uint32_t my_off = it < n_it ? P.hoff[(uint64_t)ki * P.n_items_ub + item0 + it] : 0u;
      if (__any_sync(0xFFFFFFFFu, (craw & SQ_CNT_RAW) != 0)) use_set = true;
      const uint32_t incl = warp_incl_scan(c);
      const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
      const uint32_t excl = incl - c;
      for (uint32_t f = 0; f < tot; f += 32) {
        const uint32_t idx = f + lane;
        const bool v = idx < tot;
        uint32_t j = 0;  // item (within this group) holding entry idx: number of lanes with incl <= idx
#pragma unroll
        for (int step = 16; step; step >>= 1) {
          const uint32_t t = __shfl_sync(0xFFFFFFFFu, incl, (j + step - 1) & 31);
          if (t <= idx) j += step;
        }
        const uint32_t exj = __shfl_sync(0xFFFFFFFFu, excl, j & 31);
        const uint32_t ofj = __shfl_sync(0xFFFFFFFFu, my_off, j & 31);
        const uint64_t at = (uint64_t)ki * P.hstride + ofj + (idx - exj);
        uint32_t d = SQ_EMPTY, h = 0;
        if (v) d = P.pay[at];
        bool first = d != SQ_EMPTY;
        if (use_set) {
          if (first) h = P.hsel[at];
          const uint32_t vm = __ballot_sync(0xFFFFFFFFu, first);
          const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h) & vm;
          first = first && (peers & ((1u << lane) - 1)) == 0;
          set_used = true;
          if (first) first = set_insert(S, h);
        }
        if (first) {
          if (desc_inline(d)) {
            const uint32_t base = desc_base(tb, d);
            uint32_t rest = d >> tb.tbits;
            table_vote(S, base, ki, nk);
            ++work[2];
            while (rest) {
              const uint32_t q = (uint32_t)__ffs((int)rest);
              rest &= rest - 1;
              table_vote(S, base + q, ki, nk);
              ++work[2];
            }
          } else {
            uint32_t off = __ldg(reinterpret_cast<const uint32_t*>(tb.lhdr + (d & 0x7FFFFFFFu)) + 3) + SQ_LIST_HDR;
            uint32_t t;
            do {
              t = __ldg(tb.postings + off++);
              table_vote(S, t & ~SQ_LAST, ki, nk);
              ++work[2];
            } while (!(t & SQ_LAST));
          }
        }
        __syncwarp();
        if (*(volatile uint32_t*)&S.ctr[1]) break;
      }
      if (*(volatile uint32_t*)&S.ctr[1]) break;
    }
    if (set_used) {  // leave the set empty for the next k / read
      const uint32_t n = 1u << S.set_log2;
      for (uint32_t i = lane; i < n; i += 32) S.dset[i] = SQ_EMPTY;
    }
    __syncwarp();
    if (*(volatile uint32_t*)&S.ctr[1]) break;
  }
  __syncwarp();

  if (*(volatile uint32_t*)&S.ctr[1]) {  // overflow: wipe the table completely, emit nothing
    const uint32_t n = 1u << S.tab_log2;
    for (uint32_t i = lane; i < n; i += 32) S.tkeys[i] = SQ_EMPTY;
    for (uint32_t i = lane; i < n * nk; i += 32) S.tcnt[i] = 0;
    const uint32_t ns = 1u << S.set_log2;
    for (uint32_t i = lane; i < ns; i += 32) S.dset[i] = SQ_EMPTY;
    __syncwarp();
    return 1;
  }

  // per-k maximum over the transcripts seen (sparse_chaining.cpp:76-82)
  const uint32_t n = S.ctr[0];
  uint32_t maxc[SQ_MAXK];
#pragma unroll
  for (int ki = 0; ki < SQ_MAXK; ++ki) maxc[ki] = 0;
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t s = S.tlist[i];
#pragma unroll
    for (int ki = 0; ki < SQ_MAXK; ++ki)
      if (ki < (int)nk) maxc[ki] = max(maxc[ki], S.tcnt[s * nk + ki]);
  }
  double thr[SQ_MAXK];
#pragma unroll
  for (int ki = 0; ki < SQ_MAXK; ++ki) {
#pragma unroll
    for (int d = 16; d; d >>= 1) maxc[ki] = max(maxc[ki], __shfl_xor_sync(0xFFFFFFFFu, maxc[ki], d));
    thr[ki] = P.fraction * (double)(int)maxc[ki];  // thresholds[i] = fraction * max_counts[i], :84-87
  }
  // filter + score (:90-105), collect sort keys, and clean the table slots as they are consumed.
  // The candidate buffer may alias the (now empty) dedup set, so it is re-emptied at the end.
  uint32_t nc = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    bool ok = false;
    uint32_t t = 0;
    int score = 0;
    if (i < n) {
      const uint32_t s = S.tlist[i];
      t = S.tkeys[s];
      ok = true;
#pragma unroll
      for (int ki = 0; ki < SQ_MAXK; ++ki)
        if (ki < (int)nk) {
          const int c = (int)S.tcnt[s * nk + ki];
          if ((double)c < thr[ki]) ok = false;  // counts_vec[i] < thresholds[i], :95
          score += c;
          S.tcnt[s * nk + ki] = 0;
        }
      S.tkeys[s] = SQ_EMPTY;
    }
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
    if (ok) {
      const uint32_t pos = nc + __popc(bal & ((1u << lane) - 1));
      S.cand[pos] = ((unsigned long long)(0x7FFFFFFFu - (uint32_t)score) << 32) | t;
    }
    nc += __popc(bal);
  }
  __syncwarp();
  // order: score descending (:108-109), ties by transcript id ascending (unspecified upstream)
  uint32_t padded = nc;
  if (nc > 1) {
    if (nc <= 32) {
      unsigned long long x = lane < nc ? S.cand[lane] : ~0ull;
#pragma unroll
      for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
          const unsigned long long y = __shfl_xor_sync(0xFFFFFFFFu, x, j);
          const bool up = (lane & k) == 0, lower = (lane & j) == 0;
          x = (lower == up) ? (x < y ? x : y) : (x < y ? y : x);
        }
      if (lane < nc) S.cand[lane] = x;
    } else {
      padded = 64;
      while (padded < nc) padded <<= 1;
      for (uint32_t i = nc + lane; i < padded; i += 32) S.cand[i] = ~0ull;
      __syncwarp();
      warp_bitonic(S.cand, padded);
    }
    __syncwarp();
  }
  // append to the candidate store
  unsigned long long sbase = 0;
  if (lane == 0) {
    sbase = nc ? atomicAdd(P.stage_cursor, (unsigned long long)nc) : 0ull;
    uint32_t kept = nc;
    if (sbase + nc > P.stage_cap) kept = 0;  // the host grows the store to the reported size and re-runs the vote
    P.read_loc[r] = make_uint2(P.stage_base + (uint32_t)sbase, kept);
    ListHash lh;  // rare tier: one lane folds the ordered list
    lh.init(nc);
    for (uint32_t i = 0; i < nc; ++i) {
      const unsigned long long key = S.cand[i];
      lh.add((uint32_t)key, (int32_t)(0x7FFFFFFFu - (uint32_t)(key >> 32)));
    }
    store_read_key(P, r, lh, nc ? (uint32_t)S.cand[0] : P.key_T);
  }
  sbase = __shfl_sync(0xFFFFFFFFu, sbase, 0);
  if (sbase + nc <= P.stage_cap)
    for (uint32_t i = lane; i < nc; i += 32) {
      const unsigned long long key = S.cand[i];
      P.stage[sbase + i] = make_uint2((uint32_t)key, 0x7FFFFFFFu - (uint32_t)(key >> 32));
    }
  __syncwarp();
  // the candidate buffer aliases the dedup set in the shared-memory tier: restore the empty pattern
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(S.cand);
    for (uint32_t i = lane; i < 2 * padded; i += 32) w[i] = SQ_EMPTY;
  }
  __syncwarp();
  return 0;
}

__global__ void __launch_bounds__(kVoteWarps * 32) vote_kernel(const __grid_constant__ VoteParams P) {
  extern __shared__ __align__(16) uint32_t smem[];
  if (*P.slow_count == 0) return;
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t nk = P.nk;
  const uint32_t tab = 1u << kTabLog2, set = 1u << kSetLog2;
  const uint32_t per_warp = tab + tab * nk + tab + set + 4;
  uint32_t* base = smem + warp * per_warp;
  Scratch S;
  S.dset = base;                                   // first: 8-byte aligned for the aliased u64 view
  S.cand = reinterpret_cast<unsigned long long*>(base);
  S.tkeys = base + set;
  S.tcnt = S.tkeys + tab;
  S.tlist = S.tcnt + tab * nk;
  S.ctr = S.tlist + tab;
  S.tab_log2 = kTabLog2; S.tab_maxfill = kTabMaxFill; S.set_log2 = kSetLog2; S.set_maxfill = kSetMaxFill;
  for (uint32_t i = lane; i < set; i += 32) S.dset[i] = SQ_EMPTY;
  for (uint32_t i = lane; i < tab; i += 32) S.tkeys[i] = SQ_EMPTY;
  for (uint32_t i = lane; i < tab * nk; i += 32) S.tcnt[i] = 0;
  __syncwarp();
  const uint32_t nwarps = gridDim.x * kVoteWarps;
  const uint32_t n_slow = *P.slow_count;
  uint32_t work[3] = {0, 0, 0};
  for (uint32_t i = blockIdx.x * kVoteWarps + warp; i < n_slow; i += nwarps) {
    const uint32_t r = P.slow_list[i];
    if (vote_read(P, S, r, work)) {
      if (lane == 0) {
        const uint32_t pos = atomicAdd(P.ovf_count, 1u);
        P.ovf_list[pos] = r;
        P.read_loc[r] = make_uint2(0u, 0u);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uint32_t v = work[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane == 0 && v && P.work) atomicAdd(P.work + i, (unsigned long long)v);
  }
}


// ------------------------------------------------------------------ bit-sliced path (short reads, 1..4 k values)
// The isoforms of a gene have neighbouring (internal) ids, so the posting lists a short read hits almost always
// fit a window of 32 transcripts, and nine list descriptors in ten carry the list itself (base + mask).  One
// THREAD votes for a read without any table of its own and -- for those nine in ten -- without touching memory
// beyond the read's descriptors: every hit adds 1 at the list's positions of a bit-sliced counter (per k-index NP
// planes of 32 bits = a count 0..2^NP-1 per window position, a ripple-carry over whole words).  Per-k maximum,
// the `count < ceil(fraction*max)` filter for every k, the score (sum over k, a bit-sliced adder) and the
// (score desc, transcript asc) order are word operations too.
//
// 32-bit hashes below the 5 % threshold collide: about one list in 70 joins the k-mers of two unrelated genes
// and one read in seven meets such a list (real annotations add paralogues).  A range that does not fit the
// window is counted as "outside": with n_out of them no transcript out there has more than n_out votes for that
// k, so when n_out is below some k's threshold (and not above any k's maximum) they can neither pass the filter
// nor move a maximum, and are ignored.  Lists of two ranges wait until the single-range ones have anchored the
// window.  Reads that still do not fit (several items, more hashes than the counters hold, a list that needs
// three ranges, outside ranges that could matter) are handed to the warp-per-read window kernel through mid_list.
static constexpr int kBitsBlock = 128;
static constexpr uint32_t kBitsMaxInd = 4;    // two-range lists per (read, k)

// bit-sliced counters: per k-index NP planes of 32 positions
template <int NK, int NP>
struct BitWin {
  uint32_t pl[NK][NP];
  uint32_t orm;   // positions with a count for some k
  uint32_t base;  // transcript id of position 0 (valid once orm != 0)
};

// Align a list part (first transcript `base`, 64-bit membership `mask`, bit 0 set) with the window, moving the
// window down when the part starts below it.  Returns the part's 32-bit mask in window coordinates, or 0 when
// it does not fit (nothing is modified then).
template <int NK, int NP>
__device__ __forceinline__ uint32_t win_place(BitWin<NK, NP>& W, uint32_t base, unsigned long long mask) {
  if (mask >> 32) return 0u;
  const uint32_t m = (uint32_t)mask;
  if (W.orm == 0) {
    W.base = base;
    return m;
  }
  if (base >= W.base) {
    const uint32_t d = base - W.base;
    if (d >= 32 || (d && (m >> (32 - d)))) return 0u;
    return m << d;
  }
  const uint32_t d = W.base - base;
  if (d >= 32 || (W.orm >> (32 - d))) return 0u;
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int b = 0; b < NP; ++b) W.pl[k][b] <<= d;
  W.orm <<= d;
  W.base = base;
  return m;
}

// add 1 at the positions of m for k-index K (a constant after unrolling); false (nothing changed) if a counter
// would pass 2^NP - 1
template <int NK, int NP>
__device__ __forceinline__ bool win_add(BitWin<NK, NP>& W, int K, uint32_t m) {
  uint32_t carry = m, npl[NP];
#pragma unroll
  for (int b = 0; b < NP; ++b) {
    npl[b] = W.pl[K][b] ^ carry;
    carry &= W.pl[K][b];
  }
  if (carry) return false;
#pragma unroll
  for (int b = 0; b < NP; ++b) W.pl[K][b] = npl[b];
  W.orm |= m;
  return true;
}

template <int NP>
__device__ __forceinline__ uint32_t planes_max(const uint32_t (&pl)[NP], uint32_t among) {  // MSB first
  uint32_t mx = 0, cand = among;
#pragma unroll
  for (int b = NP - 1; b >= 0; --b) {
    const uint32_t t = cand & pl[b];
    if (t) { cand = t; mx |= 1u << b; }
  }
  return mx;
}

// positions of `among` whose count is >= thr
template <int NP>
__device__ __forceinline__ uint32_t planes_at_least(const uint32_t (&pl)[NP], uint32_t among, uint32_t thr) {
  if (thr >> NP) return 0u;
  uint32_t gt = 0, eq = among;  // positions with count > / == the bits of thr seen so far
#pragma unroll
  for (int b = NP - 1; b >= 0; --b) {
    const uint32_t tbit = ((thr >> b) & 1u) ? ~0u : 0u;
    gt |= eq & pl[b] & ~tbit;
    eq &= ~(pl[b] ^ tbit);
  }
  return gt | eq;
}

template <int NP>
__device__ __forceinline__ uint32_t planes_equal(const uint32_t (&pl)[NP], uint32_t among, uint32_t c) {
  if (c >> NP) return 0u;
  uint32_t e = among;
#pragma unroll
  for (int b = 0; b < NP; ++b) e &= ((c >> b) & 1u) ? pl[b] : ~pl[b];
  return e;
}

// score planes = sum over k of the count planes (bit-sliced ripple adders); NS >= NP + 2 planes hold 4 x (2^NP - 1)
template <int NK, int NP, int NS>
__device__ __forceinline__ void planes_sum(const BitWin<NK, NP>& W, uint32_t (&s)[NS]) {
#pragma unroll
  for (int b = 0; b < NS; ++b) s[b] = b < NP ? W.pl[0][b] : 0u;
#pragma unroll
  for (int k = 1; k < NK; ++k) {
    uint32_t carry = 0;
#pragma unroll
    for (int b = 0; b < NS; ++b) {
      const uint32_t a = b < NP ? W.pl[k][b] : 0u;
      const uint32_t t = s[b] ^ a ^ carry;
      carry = (s[b] & a) | (s[b] & carry) | (a & carry);
      s[b] = t;
    }
  }
}

template <int NK, int NP>
__global__ void __launch_bounds__(kBitsBlock, NK * NP <= 10 ? 8 : 5) vote_bits_kernel(const __grid_constant__ VoteParams P) {
  __shared__ uint32_t s_ind[kBitsMaxInd][kBitsBlock];   // posting offsets of the two-range lists a (read, k) met
  __shared__ uint32_t s_work[2];
  constexpr uint32_t kMaxCount = (1u << NP) - 1;  // hashes per (read, k) the counters can take
  const uint32_t tx = threadIdx.x, lane = lane_id();
  const uint32_t r = blockIdx.x * kBitsBlock + tx;
  const bool valid = r < P.n_reads;
  if (tx < 2) s_work[tx] = 0;
  __syncthreads();
  bool defer = false;
  uint32_t wp = 0;
  BitWin<NK, NP> A;
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int b = 0; b < NP; ++b) A.pl[k][b] = 0;
  A.orm = 0;
  A.base = 0;
  uint32_t n_out[NK];
  uint32_t out_below_end = 0;  // largest end of a part counted as outside below the window
#pragma unroll
  for (int ki = 0; ki < NK; ++ki) n_out[ki] = 0;
  if (valid) {
    const uint32_t item0 = P.item_start[r];
    if (P.item_start[r + 1] - item0 != 1) defer = true;
    // counts, offsets and the first descriptors of every k-index are independent loads: all in flight at once
    uint32_t nn[NK], oo[NK], d0[NK][4];
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      nn[ki] = P.tab[ki].present ? (uint32_t)P.cnt[(uint64_t)ki * P.n_items_ub + item0] : 0u;
      oo[ki] = P.tab[ki].present ? P.hoff[(uint64_t)ki * P.n_items_ub + item0] : 0u;
    }
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      const uint32_t* pp = P.pay + (uint64_t)ki * P.hstride + oo[ki];
#pragma unroll
      for (int u = 0; u < 4; ++u) d0[ki][u] = (uint32_t)u < nn[ki] && nn[ki] <= kMaxCount ? __ldg(pp + u) : SQ_EMPTY;
    }
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      const IndexTable& tb = P.tab[ki];
      if (!tb.present || defer) continue;
      const uint32_t n = nn[ki];  // flagged counts are > kMaxCount
      if (n > kMaxCount) { defer = true; continue; }
      const uint32_t* pp = P.pay + (uint64_t)ki * P.hstride + oo[ki];
      uint32_t nind = 0;
      auto place_add = [&](uint32_t base, unsigned long long mask) {
        const uint32_t m = win_place(A, base, mask);
        if (m) {
          if (!win_add(A, ki, m)) defer = true;
          return;
        }
        // does not fit: it may be counted as "outside" only if it lies entirely beside the window's 32 ids -- above
        // them (the window only ever moves down), or below them now AND when the read is done (checked at the end)
        const uint32_t end = base + 64u - (uint32_t)__clzll((long long)mask);  // one past the last id of the part
        if (A.orm == 0) defer = true;  // wider than a window
        else if (base >= A.base + 32u) ++n_out[ki];
        else if (end <= A.base) { ++n_out[ki]; out_below_end = max(out_below_end, end); }
        else defer = true;
      };
      // One call site of place_add per loop: with several k values the kernel's code would otherwise pass the
      // instruction cache (6 000 instructions at three k values: two warps in three were waiting for instructions).
      auto vote = [&](uint32_t d) {
        if (d == SQ_EMPTY) return;
        uint32_t base;
        unsigned long long mask;
        if (desc_inline(d)) {
          mask = desc_mask(tb, d);
          base = desc_base(tb, d);
        } else {
          const uint4 hd = ld_hdr(tb.lhdr + (d & 0x7FFFFFFFu));
          if (hd.x == SQ_NOMASK) { defer = true; return; }
          if (hd.x >> 31) {  // two-range list: after the single-range ones (they anchor the window)
            if (nind < kBitsMaxInd) s_ind[nind++][tx] = hd.w; else defer = true;
            return;
          }
          mask = ((unsigned long long)hd.z << 32) | hd.y;
          base = hd.x;
        }
        wp += (uint32_t)__popcll(mask);
        place_add(base, mask);
      };
      // the read's descriptors are consecutive words: four loads in flight, then the votes (no dependent access
      // for an inline descriptor); the first four were loaded above
      uint32_t d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = d0[ki][u];
      for (uint32_t j = 0; j < n && !defer; j += 4) {
        if (j) {
#pragma unroll
          for (int u = 0; u < 4; ++u) d[u] = j + u < n ? __ldg(pp + j + u) : SQ_EMPTY;
        }
#pragma unroll(NK == 1 ? 4 : 1)
        for (int u = 0; u < 4; ++u) {
          if (!defer) vote(d[0]);
          d[0] = d[1]; d[1] = d[2]; d[2] = d[3];
        }
      }
      for (uint32_t i = 0; i < nind && !defer; ++i) {  // two-range lists: both ranges from the list's posting header
        const uint4* hp = reinterpret_cast<const uint4*>(tb.postings + s_ind[i][tx]);
        const uint4 hd = __ldg(hp), h2 = __ldg(hp + 1);
        wp += hd.x;
        // a list with both ranges outside the window still gives a transcript out there one vote at most
        const uint32_t before = n_out[ki];
        uint32_t b0 = hd.y & 0x7FFFFFFFu, b1 = h2.x;
        unsigned long long m0 = ((unsigned long long)hd.w << 32) | hd.z, m1 = ((unsigned long long)h2.z << 32) | h2.y;
#pragma unroll(NK == 1 ? 2 : 1)
        for (int h = 0; h < 2; ++h) {
          if (!defer) place_add(b0, m0);
          b0 = b1;
          m0 = m1;
        }
        if (n_out[ki] == before + 2) --n_out[ki];
      }
    }
  }
  // ---- per-k maximum over the positions (sparse_chaining.cpp:76-82), thresholds[i] = fraction * max_counts[i]
  // (:84-87), test (double)count < threshold (:95): for an integer count, count < x  <=>  count < ceil(x)
  uint32_t sa = 0;  // survivors
  uint32_t nc = 0;
  if (valid && !defer) {
    sa = A.orm;
    bool any_out = false, out_fails = false, out_above_max = false;
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      const uint32_t mx = planes_max(A.pl[ki], A.orm);
      const double t = ceil(P.fraction * (double)(int)mx);
      const uint32_t ithr = t >= 2147483647.0 ? 0x7FFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
      sa &= planes_at_least(A.pl[ki], A.orm, ithr);
      any_out |= n_out[ki] != 0;
      out_fails |= n_out[ki] < ithr;
      out_above_max |= n_out[ki] > mx;
    }
    // transcripts outside the window could matter, or the window moved onto a part counted as outside
    if (any_out && (!out_fails || out_above_max || out_below_end > A.base)) defer = true;
    else nc = (uint32_t)__popc(sa);
  }
  // hand reads that did not fit to the warp-per-read window kernel
  {
    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, valid && defer);
    if (dmask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.mid_count, (uint32_t)__popc(dmask));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (valid && defer) {
        P.mid_list[base + __popc(dmask & ((1u << lane) - 1))] = r;
        wp = 0;
        nc = 0;
      }
    }
  }
  // one staging allocation per warp
  const uint32_t incl = warp_incl_scan(nc);
  const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, incl, 31);
  unsigned long long sbase = 0;
  if (lane == 0 && wtot) sbase = atomicAdd(P.stage_cursor, (unsigned long long)wtot);
  sbase = __shfl_sync(0xFFFFFFFFu, sbase, 0);
  const bool fits = sbase + wtot <= P.stage_cap;
  sbase += incl - nc;
  if (valid) {
    P.read_loc[r] = make_uint2(P.stage_base + (uint32_t)sbase, (fits && !defer) ? nc : 0u);  // deferred reads are rewritten by the next kernel
    ListHash lh;
    lh.init(nc);
    uint32_t top = P.key_T;
    if (fits && nc) {
      // ---- score = sum of the counts over k (:100); order: score descending (:108-109), transcript ascending
      // inside a score
      constexpr int NS = NK == 1 ? NP : NP + 2;
      uint32_t SA[NS];
      planes_sum(A, SA);
      top = 0xFFFFFFFFu;
      while (sa) {
        const uint32_t c = planes_max(SA, sa);
        uint32_t e1 = planes_equal(SA, sa, c);
        sa &= ~e1;
        while (e1) {
          const uint32_t p = (uint32_t)__ffs((int)e1) - 1;
          e1 &= e1 - 1;
          P.stage[sbase] = make_uint2(A.base + p, c);
          lh.add(A.base + p, (int32_t)c);
          if (top == 0xFFFFFFFFu) top = A.base + p;
          ++sbase;
        }
      }
    }
    if (!defer) store_read_key(P, r, lh, top);  // a deferred read gets its key from the kernel that takes it
  }
  if (P.work) {
    wp = __reduce_add_sync(0xFFFFFFFFu, wp);
    if (lane == 0) {  // the last warp to arrive flushes the block's sum: no barrier at the exit
      atomicAdd(&s_work[0], wp);
      __threadfence_block();
      if (atomicAdd(&s_work[1], 1u) == kBitsBlock / 32 - 1) {
        __threadfence_block();
        const uint32_t v = *(volatile uint32_t*)&s_work[0];
        if (v) atomicAdd(P.work + 2, (unsigned long long)v);
      }
    }
  }
}

// ------------------------------------------------------------------ warp per read on window counters
// A long read (several items, hundreds of selected hashes) is voted by one WARP, and so are the short reads the
// bit-sliced kernel hands on (LISTED).  Most descriptors of an error-prone read are SQ_EMPTY; the hits are
// collected in shared memory as (hash, descriptor).  The sketch is a SET (include/sketch.h:15) and the sketch
// kernel only removed the repeats inside an item: in a read of several items a hit whose hash already occurred is
// dropped -- among the hits only, two small bitmaps tell which hits can have a twin at all.  The lists a read
// hits belong to one gene, i.e. one window of 64 internal transcript ids: the warp picks the window (the most
// common base among the lanes' first hits anchors it), and every list inside it adds 1 at its positions of a
// per-warp counter array in shared memory (64 positions per k-index, native shared-memory atomics).  Hits
// outside the window are hash collisions with unrelated genes; each can give a transcript there at most one
// vote, so with n_out of them no outside transcript has more than n_out votes: when that is below the
// threshold of some k-index (and not above any k's inside maximum) they can neither pass the filter nor move a
// maximum, and are ignored.  Everything else (a list straddling the window edge or needing three ranges, too
// many hits, outside hits that could matter) goes to the warp-per-read kernel with the general hash table
// (slow_list).  Maximum, filter, score and order over the 64 positions are done two positions per lane.
static constexpr int kLongWarps = 4;
static constexpr uint32_t kLongMaxHits = 512;

struct LongSmem {
  uint32_t hh[kLongMaxHits];  // hash of each hit
  uint32_t hd[kLongMaxHits];  // descriptor of each hit (SQ_EMPTY: dropped duplicate)
  uint32_t cnt[4][64];        // votes per k-index and window position
  uint32_t bm1[64], bm2[64];  // 2048-bit filters of the duplicate test
  unsigned long long cand[64];
};

template <int NK, bool LISTED>
__global__ void __launch_bounds__(kLongWarps * 32) vote_long_kernel(const __grid_constant__ VoteParams P) {
  extern __shared__ __align__(16) unsigned char long_smem_raw[];
  LongSmem& S = reinterpret_cast<LongSmem*>(long_smem_raw)[threadIdx.x >> 5];
  const uint32_t lane = lane_id();
  const uint32_t lt = (1u << lane) - 1;
  const uint32_t n_in = LISTED ? *P.mid_count : P.n_reads;
  uint32_t wp = 0;
  for (uint32_t ri = blockIdx.x * kLongWarps + (threadIdx.x >> 5); ri < n_in; ri += gridDim.x * kLongWarps) {
    const uint32_t r = LISTED ? P.mid_list[ri] : ri;
    const uint32_t item0 = P.item_start[r];
    const uint32_t n_it = P.item_start[r + 1] - item0;
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) { S.cnt[ki][lane] = 0; S.cnt[ki][lane + 32] = 0; }
    bool defer = false;
    bool have_win = false;
    uint32_t abase = 0;
    uint32_t n_out[NK];
    uint32_t tp = 0;
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      n_out[ki] = 0;
      const IndexTable& tb = P.tab[ki];
      if (!tb.present || defer) continue;
      // ---- collect the hits (items are walked 32 at a time, their descriptors flattened)
      uint32_t nh = 0;
      bool need_set = n_it > 1;
      for (uint32_t g = 0; g < n_it; g += 32) {
        const uint32_t it = g + lane;
        const uint32_t craw = it < n_it ? P.cnt[(uint64_t)ki * P.n_items_ub + item0 + it] : 0u;
        const uint32_t c = craw & SQ_CNT_MASK;
        const uint32_t my_off = it < n_it ? P.hoff[(uint64_t)ki * P.n_items_ub + item0 + it] : 0u;
        if (__any_sync(0xFFFFFFFFu, (craw & SQ_CNT_RAW) != 0)) need_set = true;
        const uint32_t incl = warp_incl_scan(c);
        const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t excl = incl - c;
        auto take = [&](uint32_t d, uint64_t at) {  // warp-wide: append the hits among 32 descriptors
          const uint32_t hit = __ballot_sync(0xFFFFFFFFu, d != SQ_EMPTY);
          if (d != SQ_EMPTY) {
            const uint32_t pos = nh + __popc(hit & lt);
            if (pos < kLongMaxHits) {
              S.hd[pos] = d;
              if (need_set) S.hh[pos] = __ldg(P.hsel + at);
            }
          }
          nh += __popc(hit);
        };
        // the items of a sketch warp lie back to back in the dense arrays: a group's descriptors are then one run
        // and are read as such (coalesced, two loads in flight); otherwise entry by entry through the item offsets
        const uint32_t first_off = __shfl_sync(0xFFFFFFFFu, my_off, 0);
        if (__all_sync(0xFFFFFFFFu, c == 0 || my_off == first_off + excl)) {
          const uint64_t run = (uint64_t)ki * P.hstride + first_off;
          for (uint32_t f = 0; f < tot; f += 64) {
            const uint32_t i0 = f + lane, i1 = f + 32 + lane;
            const uint32_t d0 = i0 < tot ? __ldg(P.pay + run + i0) : SQ_EMPTY;
            const uint32_t d1 = i1 < tot ? __ldg(P.pay + run + i1) : SQ_EMPTY;
            take(d0, run + i0);
            if (f + 32 < tot) take(d1, run + i1);
          }
        } else {
          for (uint32_t f = 0; f < tot; f += 32) {
            const uint32_t idx = f + lane;
            uint32_t j = 0;  // item (within this group) holding entry idx: number of lanes with incl <= idx
#pragma unroll
            for (int step = 16; step; step >>= 1) {
              const uint32_t t = __shfl_sync(0xFFFFFFFFu, incl, (j + step - 1) & 31);
              if (t <= idx) j += step;
            }
            const uint32_t exj = __shfl_sync(0xFFFFFFFFu, excl, j & 31);
            const uint32_t ofj = __shfl_sync(0xFFFFFFFFu, my_off, j & 31);
            const uint64_t at = (uint64_t)ki * P.hstride + ofj + (idx - exj);
            take(idx < tot ? __ldg(P.pay + at) : SQ_EMPTY, at);
          }
        }
      }
      if (nh > kLongMaxHits) { defer = true; continue; }
      __syncwarp();
      // ---- set semantics: a hit is dropped when an equal hash sits at a smaller index.  Filter 1 takes every
      // hit; a hit that finds its bit taken marks filter 2; only hits whose bit is in filter 2 can have a twin.
      if (need_set) {
        S.bm1[lane] = 0; S.bm1[lane + 32] = 0; S.bm2[lane] = 0; S.bm2[lane + 32] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < nh; i += 32) {
          const uint32_t b = (S.hh[i] * kHashMul) >> 21;
          if (atomicOr(&S.bm1[b >> 5], 1u << (b & 31)) & (1u << (b & 31))) atomicOr(&S.bm2[b >> 5], 1u << (b & 31));
        }
        __syncwarp();
        for (uint32_t i = lane; i < nh; i += 32) {
          const uint32_t h = S.hh[i], b = (h * kHashMul) >> 21;
          if (S.bm2[b >> 5] & (1u << (b & 31))) {
            bool dup = false;
            for (uint32_t j = 0; j < i; ++j) dup |= S.hh[j] == h;
            if (dup) S.hd[i] = SQ_EMPTY;
          }
        }
        __syncwarp();
      }
      // first range of a hit: (base, mask); SQ_NOMASK in base when the list needs three ranges
      auto first_range = [&](uint32_t d, uint32_t& base, unsigned long long& mask, uint32_t& post) {
        if (desc_inline(d)) {
          base = desc_base(tb, d);
          mask = desc_mask(tb, d);
          post = SQ_EMPTY;
        } else {
          const uint4 hd = ld_hdr(tb.lhdr + (d & 0x7FFFFFFFu));
          base = hd.x == SQ_NOMASK ? SQ_NOMASK : hd.x & 0x7FFFFFFFu;
          mask = ((unsigned long long)hd.z << 32) | hd.y;
          post = (hd.x != SQ_NOMASK && (hd.x >> 31)) ? hd.w : SQ_EMPTY;  // a second range follows in the posting header
        }
      };
      // ---- window: anchored at the most common base (in 64-id granules) among the lanes' first hits
      if (!have_win) {
        uint32_t b0 = SQ_EMPTY;
        for (uint32_t i = lane; i < nh && b0 == SQ_EMPTY; i += 32) {
          const uint32_t d = S.hd[i];
          if (d == SQ_EMPTY) continue;
          uint32_t base, post;
          unsigned long long mask;
          first_range(d, base, mask, post);
          if (base != SQ_NOMASK) b0 = base;
        }
        const uint32_t havem = __ballot_sync(0xFFFFFFFFu, b0 != SQ_EMPTY);
        if (havem) {
          const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b0 == SQ_EMPTY ? SQ_EMPTY : b0 >> 6);
          uint32_t best = b0 == SQ_EMPTY ? 0u : ((uint32_t)__popc(peers) << 8) | (31u - lane);  // most peers, lowest lane
#pragma unroll
          for (int d = 16; d; d >>= 1) best = max(best, __shfl_xor_sync(0xFFFFFFFFu, best, d));
          const uint32_t anchor = __shfl_sync(0xFFFFFFFFu, b0, 31u - (best & 255u));
          // window base = smallest range base within 64 ids below the anchor
          uint32_t mn = anchor;
          for (uint32_t i = lane; i < nh; i += 32) {
            const uint32_t d = S.hd[i];
            if (d == SQ_EMPTY) continue;
            uint32_t base, post;
            unsigned long long mask;
            first_range(d, base, mask, post);
            if (base == SQ_NOMASK) continue;
            if (base + 64 > anchor && base < mn) mn = base;
            if (post != SQ_EMPTY) {
              const uint32_t b2 = __ldg(tb.postings + post + 4);
              if (b2 + 64 > anchor && b2 < mn) mn = b2;
            }
          }
#pragma unroll
          for (int d = 16; d; d >>= 1) mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
          abase = mn;
          have_win = true;
        }
      }
      // ---- votes: every range of every remaining hit goes into the window counters, is counted as outside,
      // or (straddling the window edge / three-range list) sends the read to the general kernel
      uint32_t outside = 0;
      bool bad = false;
      auto range = [&](uint32_t b, unsigned long long m) {
        const uint32_t hi = b + 63u - (uint32_t)__clzll((long long)m);
        if (b >= abase && hi < abase + 64u) {
          uint32_t* c = &S.cnt[ki][b - abase];
          for (uint32_t lo = (uint32_t)m; lo; lo &= lo - 1) atomicAdd(c + __ffs((int)lo) - 1, 1u);
          for (uint32_t up = (uint32_t)(m >> 32); up; up &= up - 1) atomicAdd(c + 31 + __ffs((int)up), 1u);
        } else if (hi < abase || b >= abase + 64u) {
          ++outside;
        } else {
          bad = true;
        }
      };
      for (uint32_t i = lane; i < nh; i += 32) {
        const uint32_t d = S.hd[i];
        if (d == SQ_EMPTY) continue;
        uint32_t base, post;
        unsigned long long m1;
        first_range(d, base, m1, post);
        if (base == SQ_NOMASK) { bad = true; continue; }
        tp += (uint32_t)__popcll(m1);
        const uint32_t before = outside;
        range(base, m1);
        if (post != SQ_EMPTY) {
          const uint4 h2 = __ldg(reinterpret_cast<const uint4*>(tb.postings + post) + 1);
          const unsigned long long m2 = ((unsigned long long)h2.z << 32) | h2.y;
          tp += (uint32_t)__popcll(m2);
          range(h2.x, m2);
          if (outside == before + 2) --outside;  // both ranges outside: still one vote per transcript there
        }
      }
      if (__any_sync(0xFFFFFFFFu, bad)) defer = true;
      n_out[ki] = __reduce_add_sync(0xFFFFFFFFu, outside);
      __syncwarp();
    }
    // ---- per-k maximum (sparse_chaining.cpp:76-82), thresholds (:84-87), filter for every k (:90-98), score (:100)
    bool ok0 = false, ok1 = false;
    uint32_t s0 = 0, s1 = 0;
    if (!defer) {
      uint32_t c0[NK], c1[NK], ithr[NK];
      bool any_out = false, out_fails = false, out_above_max = false;
#pragma unroll
      for (int ki = 0; ki < NK; ++ki) {
        c0[ki] = S.cnt[ki][lane];
        c1[ki] = S.cnt[ki][lane + 32];
        const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, max(c0[ki], c1[ki]));
        const double t = ceil(P.fraction * (double)(int)mx);
        ithr[ki] = t >= 2147483647.0 ? 0x7FFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
        s0 += c0[ki];
        s1 += c1[ki];
        any_out |= n_out[ki] != 0;
        out_fails |= n_out[ki] < ithr[ki];
        out_above_max |= n_out[ki] > mx;
      }
      // transcripts outside the window have at most n_out votes per k: ignorable only if that cannot pass
      // some k's threshold and cannot raise any k's maximum
      if (any_out && (!out_fails || out_above_max)) defer = true;
      ok0 = s0 != 0;
      ok1 = s1 != 0;
#pragma unroll
      for (int ki = 0; ki < NK; ++ki) {
        ok0 &= c0[ki] >= ithr[ki];
        ok1 &= c1[ki] >= ithr[ki];
      }
    }
    if (defer) {
      if (lane == 0) {
        P.slow_list[atomicAdd(P.slow_count, 1u)] = r;
        P.read_loc[r] = make_uint2(0u, 0u);
      }
      __syncwarp();
      continue;
    }
    // ---- order: score descending (:108-109), transcript ascending inside a score; rank by counting
    const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, ok0), m1 = __ballot_sync(0xFFFFFFFFu, ok1);
    const uint32_t nc = (uint32_t)(__popc(m0) + __popc(m1));
    if (ok0) S.cand[__popc(m0 & lt)] = ((unsigned long long)(0x7FFFFFFFu - s0) << 32) | (abase + lane);
    if (ok1) S.cand[__popc(m0) + __popc(m1 & lt)] = ((unsigned long long)(0x7FFFFFFFu - s1) << 32) | (abase + lane + 32);
    unsigned long long sbase = 0;
    if (lane == 0) {
      sbase = nc ? atomicAdd(P.stage_cursor, (unsigned long long)nc) : 0ull;
      // (a list that does not fit: the host grows the store and re-runs the vote)
      P.read_loc[r] = make_uint2(P.stage_base + (uint32_t)sbase, sbase + nc <= P.stage_cap ? nc : 0u);
    }
    sbase = __shfl_sync(0xFFFFFFFFu, sbase, 0);
    __syncwarp();
    unsigned long long* sorted = reinterpret_cast<unsigned long long*>(S.bm1);  // 64 entries over bm1 + bm2 (done with)
    const bool fits = sbase + nc <= P.stage_cap;
    for (uint32_t i = lane; i < nc; i += 32) {
      const unsigned long long key = S.cand[i];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < nc; ++j) rank += S.cand[j] < key ? 1u : 0u;
      sorted[rank] = key;
      if (fits) {
        P.stage[sbase + rank] = make_uint2((uint32_t)key, 0x7FFFFFFFu - (uint32_t)(key >> 32));
      }
    }
    __syncwarp();
    if (lane == 0) {
      ListHash lh;
      lh.init(nc);
      for (uint32_t i = 0; i < nc; ++i) lh.add((uint32_t)sorted[i], (int32_t)(0x7FFFFFFFu - (uint32_t)(sorted[i] >> 32)));
      store_read_key(P, r, lh, nc ? (uint32_t)sorted[0] : P.key_T);
    }
    wp += tp;
    __syncwarp();
  }
  if (P.work) {
    wp = __reduce_add_sync(0xFFFFFFFFu, wp);
    if (lane == 0 && wp) atomicAdd(P.work + 2, (unsigned long long)wp);
  }
}

// every read of the batch to the general warp-per-read kernel (more than 4 k values)
__global__ void all_to_slow_kernel(const VoteParams P) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < P.n_reads) P.slow_list[r] = r;
  if (r == 0) *P.slow_count = P.n_reads;
}

// large-table path: one warp per worker, scratch in global memory
__global__ void __launch_bounds__(32) vote_overflow_kernel(const __grid_constant__ VoteParams P) {
  const uint32_t w = blockIdx.x;
  const uint32_t n = *P.ovf_count;
  if (w >= n) return;
  __shared__ uint32_t ctr[4];
  Scratch S;
  const size_t tab = (size_t)1 << P.big_cap_log2, set = (size_t)1 << P.big_set_log2;
  S.tkeys = P.big_keys + w * tab;
  S.tcnt = P.big_cnt + w * tab * P.nk;
  S.tlist = P.big_list + w * tab;
  S.dset = P.big_set + w * set;
  S.cand = P.big_cand + w * tab;
  S.ctr = ctr;
  S.tab_log2 = P.big_cap_log2;
  S.tab_maxfill = (uint32_t)tab;  // the table is sized so that every transcript fits
  S.set_log2 = P.big_set_log2;
  S.set_maxfill = (uint32_t)set;
  uint32_t work[3] = {0, 0, 0};
  for (uint32_t i = w; i < n; i += gridDim.x) {
    const uint32_t r = P.ovf_list[i];
    if (vote_read(P, S, r, work)) {
      if (lane_id() == 0) {
        atomicOr(P.flags, 2u);
        P.read_loc[r] = make_uint2(0u, 0u);
      }
    }
  }
}

__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

void launch_fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t s) {
  if (n == 0) return;
  const uint32_t grid = (uint32_t)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256);
  fill_u32_kernel<<<grid, 256, 0, s>>>(p, n, v);
}

size_t vote_smem_bytes(uint32_t nk) {
  const uint32_t tab = 1u << kTabLog2, set = 1u << kSetLog2;
  return (size_t)kVoteWarps * (tab + tab * nk + tab + set + 4) * sizeof(uint32_t);
}


// Per-device launch configuration (function attributes and occupancy-derived persistent grids).  Function
// attributes belong to the CURRENT device, so this runs once per engine, on the engine's device (sq_create):
// one process may drive several GPUs from several host threads.
template <int NK>
static cudaError_t long_occupancy(int* per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, vote_long_kernel<NK, false>, kLongWarps * 32,
                                                        sizeof(LongSmem) * kLongWarps);
}

cudaError_t vote_configure(uint32_t nk, VoteDeviceCfg* cfg) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if ((err = cudaDeviceGetAttribute(&cfg->sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return err;
  cfg->nk = nk;
  const size_t smem = vote_smem_bytes(nk);
  if ((err = cudaFuncSetAttribute(vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return err;
  int per_sm = 1;
  if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_kernel, kVoteWarps * 32, smem)) != cudaSuccess) return err;
  cfg->vote_grid = cfg->sm_count * (per_sm < 1 ? 1 : per_sm);
  per_sm = 1;
  switch (nk) {
    case 1: err = long_occupancy<1>(&per_sm); break;
    case 2: err = long_occupancy<2>(&per_sm); break;
    case 3: err = long_occupancy<3>(&per_sm); break;
    default: err = long_occupancy<4>(&per_sm); break;
  }
  if (err != cudaSuccess) return err;
  cfg->long_grid = cfg->sm_count * (per_sm < 1 ? 1 : per_sm);
  cfg->lookup_grid = cfg->sm_count * 8;
  return cudaSuccess;
}

void launch_lookup(const VoteParams& p, const VoteDeviceCfg& cfg, uint32_t ki, cudaStream_t s, uint64_t* launches) {
  if (!p.tab[ki].present) return;
  lookup_kernel<<<cfg.lookup_grid, kLookupBlock, 0, s>>>(p.hsel + (uint64_t)ki * p.hstride, p.pay + (uint64_t)ki * p.hstride,
                                                         p.hcursor + ki, p.tab[ki], p.work);
  if (launches) ++*launches;
}

template <int NK>
static cudaStream_t launch_tiers(const VoteParams& p, const VoteDeviceCfg& cfg, cudaStream_t s, cudaEvent_t ev_b,
                                 cudaStream_t tail, cudaEvent_t fork) {
  const size_t lsm = sizeof(LongSmem) * kLongWarps;
  const bool short_reads = p.force_tier != 1 && (uint64_t)(p.n_items_ub - p.n_reads) <= (uint64_t)p.n_reads + p.n_reads / 4;
  if (short_reads) {
    // mean read length up to ~320: the bit-sliced kernel takes every read it can, the window kernel the rest
    // (mid_list); the profiling events bracket the first, dominant kernel of the chain
    const uint32_t grid = (p.n_reads + kBitsBlock - 1) / kBitsBlock;
    if (p.count_bits <= 5) vote_bits_kernel<NK, 5><<<grid, kBitsBlock, 0, s>>>(p);
    else vote_bits_kernel<NK, 7><<<grid, kBitsBlock, 0, s>>>(p);
    if (ev_b) cudaEventRecord(ev_b, s);
    if (tail && fork) {  // the rest is a few latency-bound launches for ~2 % of the reads: off the main stream
      cudaEventRecord(fork, s);
      cudaStreamWaitEvent(tail, fork, 0);
      s = tail;
    }
    const uint32_t lgrid = (uint32_t)std::min<uint64_t>((uint64_t)cfg.long_grid, (p.n_reads / 16 + kLongWarps) / kLongWarps);
    vote_long_kernel<NK, true><<<lgrid, kLongWarps * 32, lsm, s>>>(p);
  } else {  // long reads span several items: one warp per read on the window counters
    const uint32_t lneed = (p.n_reads + kLongWarps - 1) / kLongWarps;
    const uint32_t lgrid = lneed < (uint32_t)cfg.long_grid ? lneed : (uint32_t)cfg.long_grid;
    vote_long_kernel<NK, false><<<lgrid, kLongWarps * 32, lsm, s>>>(p);
    if (ev_b) cudaEventRecord(ev_b, s);
    if (tail && fork) {  // the general kernel sees ~1 % of the reads and waits on their longest: off the main stream
      cudaEventRecord(fork, s);
      cudaStreamWaitEvent(tail, fork, 0);
      s = tail;
    }
  }
  return s;
}

cudaStream_t launch_vote(const VoteParams& p, const VoteDeviceCfg& cfg, cudaStream_t s, uint64_t* launches,
                         cudaEvent_t ev_a, cudaEvent_t ev_b, cudaStream_t tail, cudaEvent_t fork) {
  if (p.n_reads == 0) return s;
  const size_t smem = vote_smem_bytes(p.nk);
  uint32_t grid = (uint32_t)cfg.vote_grid;
  const uint32_t need = (p.n_reads + kVoteWarps - 1) / kVoteWarps;
  if (grid > need) grid = need;
  if (ev_a) cudaEventRecord(ev_a, s);
  switch (p.force_tier == 2 ? 99u : p.nk) {
    case 1: s = launch_tiers<1>(p, cfg, s, ev_b, tail, fork); break;
    case 2: s = launch_tiers<2>(p, cfg, s, ev_b, tail, fork); break;
    case 3: s = launch_tiers<3>(p, cfg, s, ev_b, tail, fork); break;
    case 4: s = launch_tiers<4>(p, cfg, s, ev_b, tail, fork); break;
    default:  // more than 4 k values: every read to the general kernel
      all_to_slow_kernel<<<(p.n_reads + 255) / 256, 256, 0, s>>>(p);
      if (ev_b) cudaEventRecord(ev_b, s);
      break;
  }
  vote_kernel<<<grid, kVoteWarps * 32, smem, s>>>(p);
  if (p.n_workers) vote_overflow_kernel<<<p.n_workers, 32, 0, s>>>(p);
  if (launches) *launches += 4;
  return s;
}

}  // namespace sq

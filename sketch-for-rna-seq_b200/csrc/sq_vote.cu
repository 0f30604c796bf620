// Seed lookup + per-read transcript vote (the reference's "sparse chaining", src/sparse_chaining.cpp:29-115).
//
// One warp per read.  For every k-index the warp walks the read's selected hashes (written by the sketch
// kernel), removes duplicates (the sketch is a SET, include/sketch.h:15), probes the GPU-resident bucketed
// hash table (one 32-byte sector per probe), walks the posting list of every hit and counts
// (transcript, k-index) votes in a per-warp shared-memory table.  Then: per-k maximum over the transcripts
// seen (shuffle reduction), the double-precision `count < fraction*max` filter for every k, integer score =
// sum of counts, candidates ordered by (score desc, transcript asc) and appended to the batch staging area.
// Reads whose tables do not fit in shared memory are queued for the large-table kernel, which runs the same
// code on a per-worker global-memory scratch sized for the worst case (every transcript of the index).
#include "sq_common.cuh"

namespace sq {

static constexpr int kVoteWarps = 8;
static constexpr uint32_t kTabLog2 = 8;      // 256 transcript slots per warp
static constexpr uint32_t kTabMaxFill = 192;
static constexpr uint32_t kSetLog2 = 10;     // 1024 dedup-set slots per warp (aliased with the candidate buffer)
static constexpr uint32_t kSetMaxFill = 768;
static constexpr uint32_t kHashMul = 0x9E3779B1u;

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

// one 32-byte table bucket in a single 256-bit load (sm_100), not allocated in L1: a probe has no reuse, and L1 is
// better spent on the reads' hash sectors (3.97 ms against 4.27 with allocation, 4.15 with two 128-bit loads)
__device__ __forceinline__ void ld_bucket(const uint4* p, uint4& a, uint4& c) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
               : "l"(p));
}

// look h up in the bucketed table; returns the posting offset or SQ_EMPTY
__device__ __forceinline__ uint32_t probe(const IndexTable& tb, uint32_t h) {
  uint32_t b = (h * kHashMul) >> tb.shift;
  for (uint32_t tries = 0; tries <= tb.mask; ++tries) {
    uint4 kk, oo;
    ld_bucket(tb.buckets + 2 * (size_t)b, kk, oo);
    if (kk.x == h && oo.x != SQ_EMPTY) return oo.x;
    if (kk.y == h && oo.y != SQ_EMPTY) return oo.y;
    if (kk.z == h && oo.z != SQ_EMPTY) return oo.z;
    if (kk.w == h && oo.w != SQ_EMPTY) return oo.w;
    if (oo.x == SQ_EMPTY || oo.y == SQ_EMPTY || oo.z == SQ_EMPTY || oo.w == SQ_EMPTY) return SQ_EMPTY;
    b = (b + 1) & tb.mask;
  }
  return SQ_EMPTY;
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
  const uint32_t L = P.len[r], boff = P.base_off[r] - P.bias;
  const uint32_t clen = (L + n_it - 1) / (n_it ? n_it : 1);
  if (lane == 0) { S.ctr[0] = 0; S.ctr[1] = 0; }
  __syncwarp();

  for (uint32_t ki = 0; ki < nk; ++ki) {
    const IndexTable& tb = P.tab[ki];
    if (!tb.present) continue;
    bool use_set = n_it > 1;
    if (lane == 0) { S.ctr[2] = 0; S.ctr[3] = 0; }
    __syncwarp();
    bool set_used = false;
    for (uint32_t g = 0; g < n_it; g += 32) {
      const uint32_t it = g + lane;
      const uint32_t c = it < n_it ? P.cnt[(uint64_t)ki * P.n_items_ub + item0 + it] : 0u;
      const uint32_t incl = warp_incl_scan(c);
      const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
      const uint32_t excl = incl - c;
      if (tot > 32) use_set = true;
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
        uint32_t h = 0;
        if (v) h = P.sel[(uint64_t)ki * P.slot_stride + boff + (uint64_t)(g + j) * clen + (idx - exj)];
        const uint32_t vm = __ballot_sync(0xFFFFFFFFu, v);
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, h) & vm;
        bool first = v && (peers & ((1u << lane) - 1)) == 0;
        if (use_set) {
          set_used = true;
          if (first) first = set_insert(S, h);
        }
        if (first) {
          uint32_t off = probe(tb, h);
          ++work[0];
          if (off != SQ_EMPTY) {
            ++work[1];
            off += SQ_LIST_HDR;  // skip the list header
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
  // append to the batch staging area
  unsigned long long sbase = 0;
  if (lane == 0) {
    sbase = nc ? atomicAdd(P.stage_cursor, (unsigned long long)nc) : 0ull;
    uint32_t kept = nc;
    if (sbase + nc > P.stage_cap) kept = 0;  // the host re-runs the vote with a staging area of the reported size
    P.read_soff[r] = (uint32_t)sbase;
    P.read_cnt[r] = kept;
  }
  sbase = __shfl_sync(0xFFFFFFFFu, sbase, 0);
  if (sbase + nc <= P.stage_cap)
    for (uint32_t i = lane; i < nc; i += 32) {
      const unsigned long long key = S.cand[i];
      P.stage_tid[sbase + i] = (uint32_t)key;
      P.stage_score[sbase + i] = (int32_t)(0x7FFFFFFFu - (uint32_t)(key >> 32));
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
        P.read_cnt[r] = 0;
        P.read_soff[r] = 0;
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


// ------------------------------------------------------------------ thread-per-read path (short reads)
// A short read has one item and a handful of selected hashes, so one THREAD votes for it: duplicates are
// removed by comparing with the earlier hashes of the same (read, k), every distinct hash is probed, and the
// posting lists of the hits are merged into a table (transcript -> per-k counts packed 8 bits each) that is
// kept sorted by transcript id.  Identical posting lists are stored once in the index, so hits with the same
// offset are walked once with a weight.  The table lives in shared memory, entry-major ([entry][thread]):
// a warp's accesses are bank-conflict free whatever entry each lane touches, and nothing spills to local
// memory.  Two instantiations run back to back: CAP=16 entries for every read, then CAP=48 for the reads
// that did not fit; what still does not fit (or has several items / more than kFastMaxHashes hashes for a
// k) goes to the warp-per-read kernel through slow_list.
static constexpr uint32_t kFastMaxHashes = 32;

template <typename CT, int CAP, int BLOCK, bool LISTED>  // CT: uint32_t for nk <= 4, unsigned long long for nk <= 8
__global__ void __launch_bounds__(BLOCK) vote_fast_kernel(const __grid_constant__ VoteParams P) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  CT* tc = reinterpret_cast<CT*>(fast_smem);                       // [CAP][BLOCK]
  uint32_t* tt = reinterpret_cast<uint32_t*>(tc + CAP * BLOCK);    // [CAP][BLOCK]
  __shared__ uint32_t s_warp[BLOCK / 32];
  __shared__ unsigned long long s_base;
  __shared__ uint32_t s_work[3];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, tx = threadIdx.x;
  const uint32_t gi = blockIdx.x * BLOCK + tx;
  const uint32_t n_in = LISTED ? *P.mid_count : P.n_reads;
  if (blockIdx.x * BLOCK >= n_in) return;
  const bool valid = gi < n_in;
  const uint32_t r = valid ? (LISTED ? P.mid_list[gi] : gi) : 0u;
  const uint32_t nk = P.nk;
  if (tx < 3) s_work[tx] = 0;

  uint32_t ntab = 0;
  bool defer = false;
  uint32_t wq = 0, wh = 0, wp = 0;
  // merge one posting list (ascending transcript ids) into the sorted table with weight `add`
  auto merge_list = [&](const IndexTable& tb, uint32_t off, CT add, uint32_t w, uint32_t room) {
    uint32_t p = 0, t;
    off += SQ_LIST_HDR;  // skip the list header
    do {
      t = __ldg(tb.postings + off++);
      const uint32_t tid = t & ~SQ_LAST;
      while (p < ntab && tt[p * BLOCK + tx] < tid) ++p;
      if (p < ntab && tt[p * BLOCK + tx] == tid) {
        tc[p * BLOCK + tx] += add;
      } else {
        if (ntab >= room) { defer = true; return; }
        for (uint32_t q = ntab; q > p; --q) {
          tt[q * BLOCK + tx] = tt[(q - 1) * BLOCK + tx];
          tc[q * BLOCK + tx] = tc[(q - 1) * BLOCK + tx];
        }
        tt[p * BLOCK + tx] = tid;
        tc[p * BLOCK + tx] = add;
        ++ntab;
      }
      wp += w;
    } while (!(t & SQ_LAST));
  };
  if (valid) {
    const uint32_t item0 = P.item_start[r];
    if (P.item_start[r + 1] - item0 != 1) defer = true;
    const uint32_t boff = P.base_off[r] - P.bias;
    for (uint32_t ki = 0; ki < nk && !defer; ++ki) {
      const IndexTable& tb = P.tab[ki];
      if (!tb.present) continue;
      const uint32_t n = P.cnt[(uint64_t)ki * P.n_items_ub + item0];
      if (n > kFastMaxHashes) { defer = true; break; }
      const uint32_t* hs = P.sel + (uint64_t)ki * P.slot_stride + boff;
      // (1) probe every distinct hash; park the posting offsets of the hits in this thread's column of the
      //     still unused top rows of the table (row CAP-1 downwards).  Without room the list is merged at once.
      uint32_t nh = 0;
      uint32_t* hit = tt + (CAP - 1) * BLOCK + tx;
      // the probes of one read are independent: pull every bucket towards L2/L1 before the dependent loop, so
      // the DRAM latencies overlap instead of adding up
      for (uint32_t j = 0; j < n; ++j) {
        const uint32_t b = (hs[j] * kHashMul) >> tb.shift;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.buckets + 2 * (size_t)b));
      }
      for (uint32_t j = 0; j < n && !defer; ++j) {
        const uint32_t h = hs[j];
        bool dup = false;
        for (uint32_t jj = 0; jj < j; ++jj) dup |= hs[jj] == h;
        if (dup) continue;
        const uint32_t off = probe(tb, h);
        ++wq;
        if (off == SQ_EMPTY) continue;
        ++wh;
        if (nh + ntab + 1 < (uint32_t)CAP) {
          hit[-(int)(nh * BLOCK)] = off;
          ++nh;
        } else {
          merge_list(tb, off, (CT)1 << (8 * ki), 1, (uint32_t)CAP - nh);
        }
      }
      // (2) group the parked offsets: equal offsets are the same list (weight = how many hits share it);
      //     the distinct ones are compacted to the first rows, their weights kept in the same rows of tc
      uint32_t nd = 0;
      CT* wgt = tc + (CAP - 1) * BLOCK + tx;
      for (uint32_t a = 0; a < nh; ++a) {
        const uint32_t off = hit[-(int)(a * BLOCK)];
        if (off == SQ_EMPTY) continue;
        uint32_t w = 1;
        for (uint32_t b = a + 1; b < nh; ++b)
          if (hit[-(int)(b * BLOCK)] == off) { ++w; hit[-(int)(b * BLOCK)] = SQ_EMPTY; }
        hit[-(int)(nd * BLOCK)] = off;
        wgt[-(int)(nd * BLOCK)] = (CT)w;
        ++nd;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.postings + off));
      }
      // (3) merge each distinct list once, last parked first, so the rows they occupy free up as the table grows
      for (uint32_t a = nd; a-- > 0 && !defer;) {
        const uint32_t w = (uint32_t)wgt[-(int)(a * BLOCK)];
        merge_list(tb, hit[-(int)(a * BLOCK)], (CT)w << (8 * ki), w, (uint32_t)CAP - a);
      }
    }
  }
  // hand reads that did not fit to the next tier (their work counters are recounted there)
  {
    uint32_t* list = LISTED ? P.slow_list : P.mid_list;
    uint32_t* count = LISTED ? P.slow_count : P.mid_count;
    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, valid && defer);
    if (dmask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(dmask));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (valid && defer) {
        list[base + __popc(dmask & ((1u << lane) - 1))] = r;
        wq = wh = wp = 0;
      }
    }
  }
  // per-k maximum (bytes of the packed word), threshold, filter, score; the surviving entries are sorted in
  // place (entry i is consumed before any slot <= i is overwritten): tc := 0x7FFFFFFF-score, tt := transcript
  uint32_t nc = 0;
  if (valid && !defer && ntab) {
    CT mx = 0;
    for (uint32_t i = 0; i < ntab; ++i) {
      const CT c = tc[i * BLOCK + tx];
      CT m2 = 0;
      for (uint32_t ki = 0; ki < nk; ++ki) {
        const CT a = (c >> (8 * ki)) & 255, b = (mx >> (8 * ki)) & 255;
        m2 |= (a > b ? a : b) << (8 * ki);
      }
      mx = m2;
    }
    for (uint32_t i = 0; i < ntab; ++i) {
      const CT c = tc[i * BLOCK + tx];
      const uint32_t tid = tt[i * BLOCK + tx];
      bool ok = true;
      uint32_t score = 0;
      for (uint32_t ki = 0; ki < nk; ++ki) {
        const int cc = (int)((c >> (8 * ki)) & 255), mm = (int)((mx >> (8 * ki)) & 255);
        if ((double)cc < P.fraction * (double)mm) ok = false;  // sparse_chaining.cpp:84-98
        score += (uint32_t)cc;
      }
      if (ok) {
        // insertion sort: score descending, transcript ascending
        const uint32_t inv = 0x7FFFFFFFu - score;
        uint32_t pos = nc++;
        while (pos > 0) {
          const uint32_t pi = (uint32_t)tc[(pos - 1) * BLOCK + tx], pt = tt[(pos - 1) * BLOCK + tx];
          if (pi < inv || (pi == inv && pt < tid)) break;
          tc[pos * BLOCK + tx] = (CT)pi;
          tt[pos * BLOCK + tx] = pt;
          --pos;
        }
        tc[pos * BLOCK + tx] = (CT)inv;
        tt[pos * BLOCK + tx] = tid;
      }
    }
  }
  // one staging allocation per block
  const uint32_t incl = warp_incl_scan(nc);
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (tx == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < BLOCK / 32; ++w) { const uint32_t c = s_warp[w]; s_warp[w] = tot; tot += c; }
    s_base = tot ? atomicAdd(P.stage_cursor, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (valid) {
    const unsigned long long sbase = s_base + s_warp[warp] + (incl - nc);
    const bool fits = sbase + nc <= P.stage_cap;
    P.read_soff[r] = (uint32_t)sbase;
    P.read_cnt[r] = (fits && !defer) ? nc : 0u;  // deferred reads are rewritten by the next tier
    if (fits)
      for (uint32_t i = 0; i < nc; ++i) {
        P.stage_tid[sbase + i] = tt[i * BLOCK + tx];
        P.stage_score[sbase + i] = (int32_t)(0x7FFFFFFFu - (uint32_t)tc[i * BLOCK + tx]);
      }
  }
  if (P.work) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      wq += __shfl_xor_sync(0xFFFFFFFFu, wq, d);
      wh += __shfl_xor_sync(0xFFFFFFFFu, wh, d);
      wp += __shfl_xor_sync(0xFFFFFFFFu, wp, d);
    }
    if (lane == 0) { atomicAdd(&s_work[0], wq); atomicAdd(&s_work[1], wh); atomicAdd(&s_work[2], wp); }
    __syncthreads();
    if (tx < 3 && s_work[tx]) atomicAdd(P.work + tx, (unsigned long long)s_work[tx]);
  }
}

// ------------------------------------------------------------------ 4 lanes per read (short reads)
// Middle ground between one thread and one warp per read: a QUAD of 4 lanes owns a read, a warp 8 reads.
// The read's hashes, the elements of each posting list and the table slots are dealt to the 4 lanes by
// index (j % 4), so those loops run ceil(n/4) times with little spread; the two phases whose work differs most
// from read to read (voting the posting elements, ranking the candidates) are dealt across the whole warp by
// a prefix sum over the 8 reads.  The per-warp shared memory is small enough (4.9 KB) for 44+ warps per SM.  (A warp-per-32-reads variant with every loop
// flattened over the tile by prefix sums was tried and was slower: 12 KB of tables per warp capped the
// occupancy at 25 %, see profiles/r01_notes.md.)  Tables are [slot][quad]: lane g scanning slots g, g+4, ... is bank-conflict free.
// Reads with several items, more than kQuadMaxHashes hashes for a k, more than kQuadMaxLists distinct lists
// or more than kQuadMaxFill distinct transcripts go to the warp-per-read kernel (slow_list).  nk <= 4.
static constexpr int kQuadWarps = 4;
static constexpr uint32_t kQuadSlots = 32;
static constexpr uint32_t kQuadMaxFill = 24;
static constexpr uint32_t kQuadMaxHashes = 16;
static constexpr uint32_t kQuadMaxLists = 8;

struct QuadSmem {
  uint32_t key[kQuadSlots][8];
  uint32_t cnt[kQuadSlots][8];
  uint32_t ct[kQuadMaxFill][8];        // surviving candidates: transcript
  uint32_t cs[kQuadMaxFill][8];        // surviving candidates: 0x7FFFFFFF - score
  uint32_t ho[kQuadMaxHashes][8];      // posting offset per hash (SQ_EMPTY: miss, 0xFFFFFFFE: duplicate hash)
  uint32_t lo[kQuadMaxLists][8];       // distinct posting lists: offset
  uint32_t llw[kQuadMaxLists][8];      // distinct posting lists: length (low 16 bits) | weight (high 16 bits)
  uint32_t fill[8];                    // distinct transcripts in each read's table
  uint16_t hh[kQuadMaxHashes][8];      // low 16 bits of each hash (duplicate pre-filter)
};

template <int NK, bool LISTED>  // LISTED: only the reads the bit-mask kernel left in mid_list
__global__ void __launch_bounds__(kQuadWarps * 32) vote_quad_kernel(const __grid_constant__ VoteParams P) {
  extern __shared__ __align__(16) unsigned char quad_smem_raw[];
  QuadSmem& S = reinterpret_cast<QuadSmem*>(quad_smem_raw)[threadIdx.x >> 5];
  const uint32_t lane = lane_id(), q = lane >> 2, g = lane & 3;
  const uint32_t qmask = 0xFu << (q * 4);
  constexpr uint32_t nk = NK;
  const uint32_t n_in = LISTED ? *P.mid_count : P.n_reads;
  const uint32_t n_oct = (n_in + 7) / 8;
  uint32_t wq = 0, wh = 0, wp = 0;

  for (uint32_t oct = blockIdx.x * kQuadWarps + (threadIdx.x >> 5); oct < n_oct; oct += gridDim.x * kQuadWarps) {
    const bool valid = oct * 8 + q < n_in;
    const uint32_t r = valid ? (LISTED ? P.mid_list[oct * 8 + q] : oct * 8 + q) : 0u;
#pragma unroll
    for (uint32_t sl = g; sl < kQuadSlots; sl += 4) { S.key[sl][q] = SQ_EMPTY; S.cnt[sl][q] = 0; }
    if (g == 0) S.fill[q] = 0;
    bool defer = false;
    uint32_t item0 = 0, boff = 0, fill = 0, tq = 0, th = 0;
    if (valid) {
      item0 = P.item_start[r];
      if (P.item_start[r + 1] - item0 != 1) defer = true;
      boff = P.base_off[r] - P.bias;
    }
    __syncwarp();
    for (uint32_t ki = 0; ki < nk; ++ki) {
      const IndexTable& tb = P.tab[ki];
      if (!tb.present) continue;
      uint32_t n = 0;
      if (valid && !defer) {
        n = P.cnt[(uint64_t)ki * P.n_items_ub + item0];
        if (n > kQuadMaxHashes) { defer = true; n = 0; }
      }
      const uint32_t* hs = P.sel + (uint64_t)ki * P.slot_stride + boff;
      // ---- probe: lane g takes hashes g, g+4, ...
      unsigned long long m1 = 0, m2 = 0;  // which of 64 buckets (two independent 6-bit fields) my hashes fall in
      bool maybe_dup = false;
      for (uint32_t j = g; j < n; j += 4) {
        const uint32_t h = hs[j];
        S.hh[j][q] = (uint16_t)h;
        S.ho[j][q] = probe(tb, h);
        const unsigned long long b1 = 1ull << (h & 63), b2 = 1ull << ((h >> 6) & 63);
        maybe_dup |= (m1 & b1) && (m2 & b2);
        m1 |= b1;
        m2 |= b2;
      }
      // the sketch is a set: a hash that already occurred earlier in the read must not vote twice.  Equal
      // hashes share both buckets, so the exact check runs only when some pair of hashes of the read does
      // (about 1 % of the reads).
#pragma unroll
      for (int d = 1; d <= 2; d <<= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, d), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, d);
        maybe_dup |= (m1 & o1) && (m2 & o2);
        if (d == 1) { m1 |= o1; m2 |= o2; }  // pairs (0,1) and (2,3) merged, then compared across
      }
      maybe_dup = __ballot_sync(0xFFFFFFFFu, maybe_dup) & qmask;
      __syncwarp();
      if (maybe_dup) {
        for (uint32_t j = g; j < n; j += 4) {
          const uint16_t h16 = S.hh[j][q];
          bool dup = false;
          for (uint32_t jj = 0; jj < j; ++jj)
            if (S.hh[jj][q] == h16) dup |= hs[jj] == hs[j];
          if (dup) S.ho[j][q] = 0xFFFFFFFEu;
        }
      }
      __syncwarp();
      // ---- group hits that share a posting list (equal offsets).  Lane g looks at hits g, g+4, ...: a hit is
      //      the FIRST of its list when no earlier hit has the same offset, and then its weight is the number of
      //      hits with that offset.  The firsts are compacted into (lo, llw) by a prefix inside the quad.
      uint32_t first_off[kQuadMaxHashes / 4], first_w[kQuadMaxHashes / 4];
      uint32_t my_first = 0;
#pragma unroll
      for (uint32_t m = 0; m < kQuadMaxHashes / 4; ++m) {
        const uint32_t j = g + 4 * m;
        first_w[m] = 0;
        first_off[m] = 0;
        if (j < n) {
          const uint32_t off = S.ho[j][q];
          if (off != 0xFFFFFFFEu) {  // not a duplicate hash
            ++tq;
            if (off != SQ_EMPTY) {
              ++th;
              bool seen = false;
              for (uint32_t jj = 0; jj < j; ++jj) seen |= S.ho[jj][q] == off;
              if (!seen) {
                uint32_t w = 1;
                for (uint32_t jj = j + 1; jj < n; ++jj) w += S.ho[jj][q] == off ? 1u : 0u;
                first_off[m] = off;
                first_w[m] = w;
                ++my_first;
              }
            }
          }
        }
      }
      uint32_t fpre = my_first;  // inclusive prefix of the firsts inside the quad
      {
        const uint32_t a1 = __shfl_up_sync(0xFFFFFFFFu, fpre, 1);
        if (g >= 1) fpre += a1;
        const uint32_t a2 = __shfl_up_sync(0xFFFFFFFFu, fpre, 2);
        if (g >= 2) fpre += a2;
      }
      uint32_t nd = __shfl_sync(0xFFFFFFFFu, fpre, q * 4 + 3);
      if (nd > kQuadMaxLists) { defer = true; nd = 0; }
      uint32_t nel = 0;  // posting elements of my lists
      if (nd) {
        uint32_t pos = fpre - my_first;
#pragma unroll
        for (uint32_t m = 0; m < kQuadMaxHashes / 4; ++m)
          if (first_w[m]) {
            const uint32_t len = __ldg(tb.postings + first_off[m]);  // header word: list length
            S.lo[pos][q] = first_off[m];
            S.llw[pos][q] = (first_w[m] << 16) | (len & 0xFFFFu);
            if (len > 0xFFFFu) nel = 0x40000000u;  // absurdly long list: leave the read to the warp kernel
            nel += len;
            ++pos;
          }
      }
      nel += __shfl_xor_sync(0xFFFFFFFFu, nel, 1);
      nel += __shfl_xor_sync(0xFFFFFFFFu, nel, 2);
      if (nel >= 0x40000000u) { defer = true; nel = 0; }
      // ---- vote: the posting elements of the warp's 8 reads are dealt to the 32 lanes evenly (prefix sum of the
      //      per-read element counts); each goes into its read's hash table with shared-memory atomics
      uint32_t eincl = g == 0 ? nel : 0;
#pragma unroll
      for (int d = 4; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, eincl, d);
        if ((int)lane >= d) eincl += t;
      }
      eincl = __shfl_sync(0xFFFFFFFFu, eincl, q * 4);  // inclusive prefix of my quad, on all its lanes
      const uint32_t etot = __shfl_sync(0xFFFFFFFFu, eincl, 28);
      __syncwarp();
      for (uint32_t e0 = 0; e0 < etot; e0 += 32) {
        const uint32_t e = e0 + lane;
        uint32_t qq = 0;  // owner quad = number of quads whose inclusive prefix is <= e
#pragma unroll
        for (int step = 4; step; step >>= 1) {
          const uint32_t t = __shfl_sync(0xFFFFFFFFu, eincl, ((qq + step - 1) & 7) * 4);
          if (t <= e) qq += step;
        }
        const uint32_t qel = __shfl_sync(0xFFFFFFFFu, nel, (qq & 7) * 4);
        const uint32_t qex = __shfl_sync(0xFFFFFFFFu, eincl, (qq & 7) * 4) - qel;
        if (e < etot) {
          uint32_t idx = e - qex, i = 0, lw = S.llw[0][qq];
          while (idx >= (lw & 0xFFFFu)) { idx -= lw & 0xFFFFu; lw = S.llw[++i][qq]; }
          const uint32_t tid = __ldg(tb.postings + S.lo[i][qq] + SQ_LIST_HDR + idx) & ~SQ_LAST;
          const uint32_t add = (lw >> 16) << (8 * ki);
          uint32_t sl = (tid * kHashMul) >> 27;
          uint32_t tries = 0;
          for (; tries < kQuadSlots; ++tries) {
            const uint32_t old = atomicCAS(&S.key[sl][qq], SQ_EMPTY, tid);
            if (old == SQ_EMPTY) atomicAdd(&S.fill[qq], 1u);
            if (old == SQ_EMPTY || old == tid) { atomicAdd(&S.cnt[sl][qq], add); break; }
            sl = (sl + 1) & (kQuadSlots - 1);
          }
          if (tries == kQuadSlots) S.fill[qq] = 1000;  // table full
          wp += lw >> 16;
        }
      }
      __syncwarp();
    }
    fill = S.fill[q];
    if (fill > kQuadMaxFill) defer = true;  // too many distinct transcripts for the table
    // ---- per-k maximum over the table: lane g scans slots g, g+4, ...; combine inside the quad
    uint32_t mx = 0;
#pragma unroll
    for (uint32_t sl = g; sl < kQuadSlots; sl += 4) {
      const uint32_t c = S.cnt[sl][q];
      uint32_t m2 = 0;
#pragma unroll
      for (int ki = 0; ki < NK; ++ki) {
        const uint32_t a = (c >> (8 * ki)) & 255, b = (mx >> (8 * ki)) & 255;
        m2 |= (a > b ? a : b) << (8 * ki);
      }
      mx = m2;
    }
#pragma unroll
    for (int d = 1; d <= 2; d <<= 1) {
      const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, mx, d);
      uint32_t m2 = 0;
#pragma unroll
      for (int ki = 0; ki < NK; ++ki) {
        const uint32_t a = (o >> (8 * ki)) & 255, b = (mx >> (8 * ki)) & 255;
        m2 |= (a > b ? a : b) << (8 * ki);
      }
      mx = m2;
    }
    // thresholds[i] = fraction * max_counts[i] (:84-87), test (double)count < threshold (:95); for an integer
    // count, count < x  <=>  count < ceil(x)
    uint32_t ithr[NK];
#pragma unroll
    for (int ki = 0; ki < NK; ++ki) {
      const double t = ceil(P.fraction * (double)(int)((mx >> (8 * ki)) & 255));
      ithr[ki] = t >= 2147483647.0 ? 0x7FFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
    }
    // ---- survivors of my slots, compacted into the quad's candidate list (positions by a 4-lane prefix)
    uint32_t mine = 0;  // bit m set: slot g+4m passes
    uint32_t my_n = 0;
    if (valid && !defer) {
#pragma unroll
      for (uint32_t m = 0; m < kQuadSlots / 4; ++m) {
        const uint32_t sl = g + 4 * m;
        const uint32_t c = S.cnt[sl][q];
        bool ok = S.key[sl][q] != SQ_EMPTY;
#pragma unroll
        for (int ki = 0; ki < NK; ++ki)
          if (((c >> (8 * ki)) & 255) < ithr[ki]) ok = false;
        if (ok) { mine |= 1u << m; ++my_n; }
      }
    }
    uint32_t pre = my_n;  // inclusive prefix inside the quad
    {
      const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, pre, 1);
      if (g >= 1) pre += a;
      const uint32_t b = __shfl_up_sync(0xFFFFFFFFu, pre, 2);
      if (g >= 2) pre += b;
    }
    const uint32_t nc = __shfl_sync(0xFFFFFFFFu, pre, q * 4 + 3);
    {
      uint32_t pos = pre - my_n;
      for (uint32_t m = 0; m < kQuadSlots / 4; ++m)
        if (mine & (1u << m)) {
          const uint32_t sl = g + 4 * m;
          const uint32_t c = S.cnt[sl][q];
          uint32_t score = 0;
#pragma unroll
          for (int ki = 0; ki < NK; ++ki) score += (c >> (8 * ki)) & 255;
          S.ct[pos][q] = S.key[sl][q];
          S.cs[pos][q] = 0x7FFFFFFFu - score;
          ++pos;
        }
    }
    if (valid && !defer) { wq += tq; wh += th; }  // every lane counted its own hashes
    // hand reads that did not fit (long reads, many hashes, many lists, many transcripts) to the
    // warp-per-read kernel: they are the heavy ones, a whole warp suits them better than one thread
    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, valid && defer && g == 0);
    if (dmask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.slow_count, (uint32_t)__popc(dmask));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (valid && defer && g == 0) P.slow_list[base + __popc(dmask & ((1u << lane) - 1))] = r;
    }
    // one staging allocation per warp (8 reads): exclusive prefix of nc over the quads
    uint32_t qincl = g == 0 ? nc : 0;
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, qincl, d);
      if ((int)lane >= d) qincl += t;
    }
    qincl = __shfl_sync(0xFFFFFFFFu, qincl, q * 4);      // inclusive prefix of my quad, on all its lanes
    const uint32_t wtot = __shfl_sync(0xFFFFFFFFu, qincl, 28);
    unsigned long long wbase = 0;
    if (lane == 0 && wtot) wbase = atomicAdd(P.stage_cursor, (unsigned long long)wtot);
    wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
    const bool fits = wbase + wtot <= P.stage_cap;
    const unsigned long long rbase = wbase + (qincl - nc);
    if (valid && g == 0) {
      P.read_soff[r] = (uint32_t)rbase;
      P.read_cnt[r] = (fits && !defer) ? nc : 0u;
    }
    __syncwarp();
    // ---- order (score desc, transcript asc).  The candidates of the warp's 8 reads are dealt to the 32 lanes
    //      evenly (lists differ a lot in size); each is ranked against its own read's list and written at its
    //      rank, which is its position in the ordered output.
    if (fits)
      for (uint32_t e0 = 0; e0 < wtot; e0 += 32) {
        const uint32_t e = e0 + lane;
        uint32_t qq = 0;  // owner quad = number of quads whose inclusive prefix is <= e
#pragma unroll
        for (int step = 4; step; step >>= 1) {
          const uint32_t t = __shfl_sync(0xFFFFFFFFu, qincl, ((qq + step - 1) & 7) * 4);
          if (t <= e) qq += step;
        }
        const uint32_t qn = __shfl_sync(0xFFFFFFFFu, nc, (qq & 7) * 4);
        const uint32_t qex = __shfl_sync(0xFFFFFFFFu, qincl, (qq & 7) * 4) - qn;
        if (e < wtot) {
          const uint32_t idx = e - qex;
          const uint32_t inv = S.cs[idx][qq], tid = S.ct[idx][qq];
          uint32_t rank = 0;
          for (uint32_t j = 0; j < qn; ++j) {
            const uint32_t pi = S.cs[j][qq], pt = S.ct[j][qq];
            rank += (pi < inv || (pi == inv && pt < tid)) ? 1u : 0u;
          }
          P.stage_tid[wbase + qex + rank] = tid;
          P.stage_score[wbase + qex + rank] = (int32_t)(0x7FFFFFFFu - inv);
        }
      }
    __syncwarp();
  }
  if (P.work) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      wq += __shfl_xor_sync(0xFFFFFFFFu, wq, d);
      wh += __shfl_xor_sync(0xFFFFFFFFu, wh, d);
      wp += __shfl_xor_sync(0xFFFFFFFFu, wp, d);
    }
    if (lane == 0) {
      if (wq) atomicAdd(P.work + 0, (unsigned long long)wq);
      if (wh) atomicAdd(P.work + 1, (unsigned long long)wh);
      if (wp) atomicAdd(P.work + 2, (unsigned long long)wp);
    }
  }
}

// ------------------------------------------------------------------ bit-mask path (short reads, one k)
// The isoforms of a gene have neighbouring ids, so the posting lists a short read hits almost always fit a
// window of 64 transcripts.  The index keeps, per distinct list, the window base and a 64-bit membership mask,
// and for one-k indexes a second, "direct" hash table whose entries carry that header next to the key: a probe
// is one 32-byte bucket and nothing else.  One THREAD votes for a read without any table of its own: every hit
// adds 1 at the list's positions of a bit-sliced counter (5 planes of 64 bits = a count 0..31 per window
// position, a ripple-carry over whole words).  Per-position maximum, the `count < ceil(fraction*max)` filter
// and the (score desc, transcript asc) order are word operations too.  The block first deals its reads to the
// threads in order of hash count, so that the probe loop of a warp diverges little.
//
// 32-bit hashes below the 5 % threshold collide: about one list in 70 joins the k-mers of two unrelated genes
// and one read in seven meets such a list (real annotations add paralogues).  A list header therefore holds
// up to two (base, mask) ranges, and the thread keeps a second, 2-plane window (counts up to 3) for an id
// range away from the first; the two windows never overlap, so maximum, filter and order stay word operations.
// Reads that still do not fit (several items, > 16 hashes, > 4 two-range lists, a list that needs three
// ranges, a third id range, a distant count above 3) are handed to the 4-lanes-per-read kernel through mid_list.
static constexpr int kBitsBlock = 128;
static constexpr uint32_t kBitsMaxHashes = 16;
static constexpr uint32_t kBitsMaxInd = 4;  // two-range lists per read

// bit-sliced counters: NP planes of 64 positions
template <int NP>
struct BitWin {
  unsigned long long pl[NP];
  unsigned long long orm;  // positions with a count
  uint32_t base;
  bool have;
};

// make room for a list part (window base `base`, membership `mask`) in W without touching the id range of the
// other window O; on success `mask` is aligned to W.  Nothing is modified on failure.
template <int NP, int NO>
__device__ __forceinline__ bool win_place(BitWin<NP>& W, const BitWin<NO>& O, uint32_t base, unsigned long long& mask) {
  if (!W.have || base < W.base) {
    if (W.have) {
      const uint32_t d = W.base - base;
      if (d >= 64 || (W.orm >> (64 - d)) != 0) return false;
    }
    if (O.have && base < O.base + 64 && O.base < base + 64) return false;  // ranges would overlap
    if (W.have) {
      const uint32_t d = W.base - base;
#pragma unroll
      for (int b = 0; b < NP; ++b) W.pl[b] <<= d;
      W.orm <<= d;
    }
    W.base = base;
    W.have = true;
    return true;
  }
  const uint32_t d2 = base - W.base;
  if (d2 >= 64 || (d2 && (mask >> (64 - d2)) != 0)) return false;
  mask <<= d2;
  return true;
}

// add weight w at the positions of mask; false if a counter would pass 2^NP - 1
template <int NP>
__device__ __forceinline__ bool win_add(BitWin<NP>& W, unsigned long long mask, uint32_t w) {
  if (w >> NP) return false;
  unsigned long long carry = 0, npl[NP];
#pragma unroll
  for (int b = 0; b < NP; ++b) {
    const unsigned long long a = ((w >> b) & 1u) ? mask : 0ull;
    npl[b] = W.pl[b] ^ a ^ carry;
    carry = (W.pl[b] & a) | (W.pl[b] & carry) | (a & carry);
  }
  if (carry) return false;
#pragma unroll
  for (int b = 0; b < NP; ++b) W.pl[b] = npl[b];
  W.orm |= mask;
  return true;
}

template <int NP>
__device__ __forceinline__ uint32_t win_max(const BitWin<NP>& W) {  // MSB first
  uint32_t mx = 0;
  unsigned long long cand = W.orm;
#pragma unroll
  for (int b = NP - 1; b >= 0; --b) {
    const unsigned long long t = cand & W.pl[b];
    if (t) { cand = t; mx |= 1u << b; }
  }
  return mx;
}

template <int NP>
__device__ __forceinline__ unsigned long long win_at_least(const BitWin<NP>& W, uint32_t thr) {
  if (thr >> NP) return 0ull;
  unsigned long long gt = 0, eq = W.orm;  // positions with count > / == the bits of thr seen so far
#pragma unroll
  for (int b = NP - 1; b >= 0; --b) {
    const unsigned long long tbit = ((thr >> b) & 1u) ? ~0ull : 0ull;
    gt |= eq & W.pl[b] & ~tbit;
    eq &= ~(W.pl[b] ^ tbit);
  }
  return gt | eq;
}

template <int NP>
__device__ __forceinline__ unsigned long long win_equal(const BitWin<NP>& W, unsigned long long among, uint32_t c) {
  if (c >> NP) return 0ull;
  unsigned long long e = among;
#pragma unroll
  for (int b = 0; b < NP; ++b) e &= ((c >> b) & 1u) ? W.pl[b] : ~W.pl[b];
  return e;
}

// first bucket of the direct table already loaded: the entry of h, or .y == SQ_DIRECT_EMPTY when h is not a key
__device__ __forceinline__ uint4 direct_resolve(const IndexTable& tb, uint32_t h, uint32_t b, uint4 a, uint4 c) {
  for (uint32_t tries = 0;; ++tries) {
    if (a.x == h && a.y != SQ_DIRECT_EMPTY) return a;
    if (c.x == h && c.y != SQ_DIRECT_EMPTY) return c;
    if (a.y == SQ_DIRECT_EMPTY || c.y == SQ_DIRECT_EMPTY || tries >= tb.dmask) break;
    b = (b + 1) & tb.dmask;  // full bucket without the key: next one (rare)
    ld_bucket(tb.direct + 2 * (size_t)b, a, c);
  }
  return make_uint4(h, SQ_DIRECT_EMPTY, 0u, 0u);
}

__global__ void __launch_bounds__(kBitsBlock, 8) vote_bits_kernel(const __grid_constant__ VoteParams P) {
  __shared__ uint32_t s_h[kBitsMaxHashes][kBitsBlock];  // the read's hashes so far (exact duplicate check)
  __shared__ uint32_t s_ind[kBitsMaxInd][kBitsBlock];   // posting offsets of the two-range lists it met
  __shared__ uint32_t s_work[4];
  __shared__ uint32_t s_hist[kBitsMaxHashes + 2];
  __shared__ uint32_t s_perm[kBitsBlock];
  const uint32_t tx = threadIdx.x, lane = lane_id(), warp = tx >> 5;
  uint32_t r = blockIdx.x * kBitsBlock + tx;
  bool valid = r < P.n_reads;
  const IndexTable& tb = P.tab[0];
  if (tx < 4) s_work[tx] = 0;
  bool defer = false;
  uint32_t wq = 0, wh = 0, wp = 0;
  BitWin<5> A = {{0, 0, 0, 0, 0}, 0, 0, false};  // the read's own gene: counts up to 31
  BitWin<2> B = {{0, 0}, 0, 0, false};           // a second, distant id range (hash collisions, paralogs): up to 3
  unsigned long long sa = 0, sb = 0;             // survivors
  uint32_t mx = 0, ithr = 0, nc = 0;
  uint32_t n = 0;
  // Every per-thread loop below runs for the longest lane of its warp, and the hash count of a read varies
  // 2..14: deal the block's reads to the threads in order of hash count (counting sort in shared memory), so
  // that a warp holds reads of similar size.  `r` is the read this thread works for from here on.
  if (tx < kBitsMaxHashes + 2) s_hist[tx] = 0;
  __syncthreads();
  {
    uint32_t key = kBitsMaxHashes + 1;  // beyond the batch, several items or too many hashes: no work, go last
    if (valid && tb.present) {
      const uint32_t item0 = P.item_start[r];
      if (P.item_start[r + 1] - item0 == 1) key = min((uint32_t)P.cnt[item0], kBitsMaxHashes + 1);
    }
    // The read's hashes go into this thread's shared-memory column now (one or two 256-bit loads when the slot
    // is 32-byte aligned, as it is for reads packed at a fixed stride): their DRAM latency passes behind the
    // deal's barriers, and the vote loop below never waits on global memory for a hash.
    if (key <= kBitsMaxHashes && key) {
      const uint32_t* hp = P.sel + (P.base_off[r] - P.bias);
      if ((reinterpret_cast<uintptr_t>(hp) & 31) == 0 && P.len[r] >= kBitsMaxHashes) {  // the slot is len words long
        for (uint32_t j0 = 0; j0 < key; j0 += 8) {
          uint4 a, c;
          asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                       : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
                       : "l"(hp + j0));
          s_h[j0 + 0][tx] = a.x; s_h[j0 + 1][tx] = a.y; s_h[j0 + 2][tx] = a.z; s_h[j0 + 3][tx] = a.w;
          s_h[j0 + 4][tx] = c.x; s_h[j0 + 5][tx] = c.y; s_h[j0 + 6][tx] = c.z; s_h[j0 + 7][tx] = c.w;
        }
      } else {
        for (uint32_t j = 0; j < key; ++j) s_h[j][tx] = __ldg(hp + j);
      }
    }
    const uint32_t rank = atomicAdd(&s_hist[key], 1u);
    __syncthreads();
    if (tx == 0) {
      uint32_t acc = 0;
      for (uint32_t i = 0; i < kBitsMaxHashes + 2; ++i) { const uint32_t c = s_hist[i]; s_hist[i] = acc; acc += c; }
    }
    __syncthreads();
    s_perm[s_hist[key] + rank] = tx;
    __syncthreads();
  }
  const uint32_t col = s_perm[tx];  // the thread that loaded the hashes of the read this thread votes for
  r = blockIdx.x * kBitsBlock + col;
  valid = r < P.n_reads;
  if (valid && tb.present) {
    const uint32_t item0 = P.item_start[r];
    if (P.item_start[r + 1] - item0 != 1) defer = true;
    n = defer ? 0u : (uint32_t)P.cnt[item0];
    if (n > kBitsMaxHashes) { defer = true; n = 0; }
  }
  if (valid && tb.present) {
    // ---- probe the read's hashes two at a time (independent loads in flight) and vote hit by hit.  After
    // the deal above the lanes of a warp have about the same number of hashes, so this loop diverges little.
    uint32_t m1 = 0, m2 = 0, nind = 0;
    auto seen = [&](uint32_t h, uint32_t upto) {  // is h one of the read's first `upto` hashes?
      const uint32_t b1 = 1u << (h & 31), b2 = 1u << ((h >> 5) & 31);
      bool dup = false;
      if ((m1 & b1) && (m2 & b2))  // an equal hash would share both filter bits: exact check (rare)
        for (uint32_t jj = 0; jj < upto; ++jj) dup |= s_h[jj][col] == h;
      m1 |= b1;
      m2 |= b2;
      return dup;
    };
    auto vote = [&](const uint4& e) {
      ++wq;
      if (e.y == SQ_DIRECT_EMPTY) return;
      ++wh;
      if (e.y == SQ_NOMASK) { defer = true; return; }
      if (e.y >> 31) {  // two-range list: after the single-range ones (they fix window A)
        if (nind < kBitsMaxInd) s_ind[nind++][tx] = e.z; else defer = true;
        return;
      }
      unsigned long long mask = ((unsigned long long)e.w << 32) | e.z;
      wp += (uint32_t)__popcll(mask);
      if (win_place(A, B, e.y, mask)) win_add(A, mask, 1u);
      else if (!(win_place(B, A, e.y, mask) && win_add(B, mask, 1u))) defer = true;
    };
    for (uint32_t j = 0; j < n && !defer; j += 2) {
      const bool two = j + 1 < n;
      const uint32_t h1 = s_h[j][col], h2 = two ? s_h[j + 1][col] : 0u;
      const bool v1 = !seen(h1, j);
      const bool v2 = two && !seen(h2, j + 1);
      const uint32_t b1 = (h1 * kHashMul) >> tb.dshift, b2 = (h2 * kHashMul) >> tb.dshift;
      uint4 a1 = make_uint4(0, 0, 0, 0), c1 = a1, a2 = a1, c2 = a1;
      if (v1) ld_bucket(tb.direct + 2 * (size_t)b1, a1, c1);
      if (v2) ld_bucket(tb.direct + 2 * (size_t)b2, a2, c2);
      if (v1) vote(direct_resolve(tb, h1, b1, a1, c1));
      if (v2 && !defer) vote(direct_resolve(tb, h2, b2, a2, c2));
    }
    for (uint32_t i = 0; i < nind && !defer; ++i) {  // two-range lists: both ranges from the list header
      const uint4* hp = reinterpret_cast<const uint4*>(tb.postings + s_ind[i][tx]);
      const uint4 hd = __ldg(hp), h2 = __ldg(hp + 1);
      unsigned long long mask = ((unsigned long long)hd.w << 32) | hd.z;
      unsigned long long mask2 = ((unsigned long long)h2.z << 32) | h2.y;
      wp += hd.x;
      if (win_place(A, B, hd.y & 0x7FFFFFFFu, mask)) win_add(A, mask, 1u);
      else if (!(win_place(B, A, hd.y & 0x7FFFFFFFu, mask) && win_add(B, mask, 1u))) { defer = true; break; }
      if (win_place(A, B, h2.x, mask2)) win_add(A, mask2, 1u);
      else if (!(win_place(B, A, h2.x, mask2) && win_add(B, mask2, 1u))) { defer = true; break; }
    }
    if (!defer && (A.orm | B.orm)) {
      // ---- maximum over the positions (sparse_chaining.cpp:76-82)
      mx = max(win_max(A), win_max(B));
      // thresholds[i] = fraction * max_counts[i] (:84-87), test (double)count < threshold (:95); for an
      // integer count, count < x  <=>  count < ceil(x)
      const double t = ceil(P.fraction * (double)(int)mx);
      ithr = t >= 2147483647.0 ? 0x7FFFFFFFu : (t <= 0.0 ? 0u : (uint32_t)t);
      sa = win_at_least(A, ithr);
      sb = win_at_least(B, ithr);
      nc = (uint32_t)(__popcll(sa) + __popcll(sb));
    }
  }
  // hand reads that did not fit to the 4-lanes-per-read kernel
  {
    const uint32_t dmask = __ballot_sync(0xFFFFFFFFu, valid && defer);
    if (dmask) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(P.mid_count, (uint32_t)__popc(dmask));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (valid && defer) {
        P.mid_list[base + __popc(dmask & ((1u << lane) - 1))] = r;
        wq = wh = wp = 0;
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
    P.read_soff[r] = (uint32_t)sbase;
    P.read_cnt[r] = (fits && !defer) ? nc : 0u;  // deferred reads are rewritten by the next kernel
    if (fits && nc) {
      // ---- order: score descending (:108-109), transcript ascending inside a score; the two windows
      // cover disjoint id ranges, so inside a score the lower window goes first
      const uint32_t lowest = ithr > 1 ? ithr : 1;
      const bool b_first = B.have && B.base < A.base;
      for (uint32_t c = mx; c >= lowest; --c) {
        unsigned long long e1 = win_equal(A, sa, c), e2 = win_equal(B, sb, c);
        uint32_t base1 = A.base, base2 = B.base;
        if (b_first) {
          const unsigned long long te = e1; e1 = e2; e2 = te;
          base1 = B.base; base2 = A.base;
        }
        while (e1) {
          const uint32_t p = (uint32_t)__ffsll((long long)e1) - 1;
          e1 &= e1 - 1;
          P.stage_tid[sbase] = base1 + p;
          P.stage_score[sbase] = (int32_t)c;
          ++sbase;
        }
        while (e2) {
          const uint32_t p = (uint32_t)__ffsll((long long)e2) - 1;
          e2 &= e2 - 1;
          P.stage_tid[sbase] = base2 + p;
          P.stage_score[sbase] = (int32_t)c;
          ++sbase;
        }
      }
    }
  }
  if (P.work) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      wq += __shfl_xor_sync(0xFFFFFFFFu, wq, d);
      wh += __shfl_xor_sync(0xFFFFFFFFu, wh, d);
      wp += __shfl_xor_sync(0xFFFFFFFFu, wp, d);
    }
    if (lane == 0) {  // the last warp to arrive flushes the block's sums: no barrier at the exit
      atomicAdd(&s_work[0], wq);
      atomicAdd(&s_work[1], wh);
      atomicAdd(&s_work[2], wp);
      __threadfence_block();
      if (atomicAdd(&s_work[3], 1u) == kBitsBlock / 32 - 1) {
        __threadfence_block();
        for (int i = 0; i < 3; ++i) {
          const uint32_t v = *(volatile uint32_t*)&s_work[i];
          if (v) atomicAdd(P.work + i, (unsigned long long)v);
        }
      }
    }
  }
}

template <typename CT>
static cudaStream_t launch_fast_tiers(const VoteParams& p, cudaStream_t s, cudaEvent_t ev_b, cudaStream_t tail,
                                      cudaEvent_t fork) {
  constexpr int capA = 16, blkA = 256, capB = 48, blkB = 128;
  constexpr size_t smA = (size_t)capA * blkA * (4 + sizeof(CT)), smB = (size_t)capB * blkB * (4 + sizeof(CT));
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(vote_fast_kernel<CT, capA, blkA, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    cudaFuncSetAttribute(vote_fast_kernel<CT, capB, blkB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
    attr = true;
  }
  if (p.nk <= 4) {
    // 4-lanes-per-read kernel first (persistent grid), what does not fit goes to the CAP=48 thread tier
    static int quad_grid[5] = {0, 0, 0, 0, 0};
    const int nkq = p.nk < 4 ? (int)p.nk : 4;
    const size_t qsm = sizeof(QuadSmem) * kQuadWarps;
    if (!quad_grid[nkq]) {
      int dev = 0, sms = 0, per_sm = 1;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      switch (nkq) {
        case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_quad_kernel<1, true>, kQuadWarps * 32, qsm); break;
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_quad_kernel<2, false>, kQuadWarps * 32, qsm); break;
        case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_quad_kernel<3, false>, kQuadWarps * 32, qsm); break;
        default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_quad_kernel<4, false>, kQuadWarps * 32, qsm); break;
      }
      quad_grid[nkq] = sms * (per_sm < 1 ? 1 : per_sm);
    }
    const uint32_t need = ((p.n_reads + 7) / 8 + kQuadWarps - 1) / kQuadWarps;
    const uint32_t qgrid = need < (uint32_t)quad_grid[nkq] ? need : (uint32_t)quad_grid[nkq];
    switch (nkq) {
      case 1:
        if (p.tab[0].direct && (uint64_t)(p.n_items_ub - p.n_reads) <= (uint64_t)p.n_reads + p.n_reads / 4) {
          // one k, mean read length up to ~320: the bit-mask kernel takes every read it can, the quad kernel
          // the rest (mid_list); the profiling events bracket the first, dominant kernel of the chain
          vote_bits_kernel<<<(p.n_reads + kBitsBlock - 1) / kBitsBlock, kBitsBlock, 0, s>>>(p);
          if (ev_b) cudaEventRecord(ev_b, s);
          if (tail && fork) {  // the rest is a few latency-bound launches for ~2 % of the reads: off the main stream
            cudaEventRecord(fork, s);
            cudaStreamWaitEvent(tail, fork, 0);
            s = tail;
          }
          vote_quad_kernel<1, true><<<qgrid, kQuadWarps * 32, qsm, s>>>(p);
        } else {  // long reads span several items: straight to the quad kernel
          vote_quad_kernel<1, false><<<qgrid, kQuadWarps * 32, qsm, s>>>(p);
          if (ev_b) cudaEventRecord(ev_b, s);
        }
        break;
      case 2: vote_quad_kernel<2, false><<<qgrid, kQuadWarps * 32, qsm, s>>>(p); break;
      case 3: vote_quad_kernel<3, false><<<qgrid, kQuadWarps * 32, qsm, s>>>(p); break;
      default: vote_quad_kernel<4, false><<<qgrid, kQuadWarps * 32, qsm, s>>>(p); break;
    }
    if (ev_b && nkq != 1) cudaEventRecord(ev_b, s);
  } else {
    // more than 4 k values: 64-bit packed counters, thread-per-read tiers (16 then 48 table entries)
    vote_fast_kernel<CT, capA, blkA, false><<<(p.n_reads + blkA - 1) / blkA, blkA, smA, s>>>(p);
    if (ev_b) cudaEventRecord(ev_b, s);
    vote_fast_kernel<CT, capB, blkB, true><<<(p.n_reads + blkB - 1) / blkB, blkB, smB, s>>>(p);
  }
  return s;
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
        P.read_cnt[r] = 0;
        P.read_soff[r] = 0;
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

cudaStream_t launch_vote(const VoteParams& p, cudaStream_t s, uint64_t* launches, cudaEvent_t ev_a, cudaEvent_t ev_b,
                         cudaStream_t tail, cudaEvent_t fork) {
  if (p.n_reads == 0) return s;
  static int sm_count = 0, configured_nk = -1;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t smem = vote_smem_bytes(p.nk);
  if (configured_nk != (int)p.nk) {
    cudaFuncSetAttribute(vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured_nk = (int)p.nk;
  }
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vote_kernel, kVoteWarps * 32, smem);
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = (uint32_t)(sm_count * per_sm);
  const uint32_t need = (p.n_reads + kVoteWarps - 1) / kVoteWarps;
  if (grid > need) grid = need;
  if (ev_a) cudaEventRecord(ev_a, s);
  s = p.nk <= 4 ? launch_fast_tiers<uint32_t>(p, s, ev_b, tail, fork)
                : launch_fast_tiers<unsigned long long>(p, s, ev_b, tail, fork);
  vote_kernel<<<grid, kVoteWarps * 32, smem, s>>>(p);
  vote_overflow_kernel<<<p.n_workers, 32, 0, s>>>(p);
  if (launches) *launches += 4;
  return s;
}

}  // namespace sq

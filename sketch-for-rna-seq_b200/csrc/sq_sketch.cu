// Rolling ntHash2 forward hash + FracMinHash threshold filter on 2-bit packed reads (sm_100a).
//
// Replaces nthash::NtHash::roll()/get_forward_hash() at its call sites (reference src/sketch.cpp:31-33)
// and the threshold filter of createSketch_FracMinhash_direct (src/sketch.cpp:24-39).
//
// Work decomposition: one thread per "item" = up to SQ_CHUNK consecutive window-end positions of one read
// (a 150 bp read is one item, a 10 kb read ~40 items).  A warp's 32 items are contiguous in the packed
// stream, so the warp stages its whole span in shared memory with coalesced 128-bit loads and every lane
// then rolls through its own bases from shared memory.  Per step the lane does one 8-byte table lookup
// (seed[in] ^ rot_k(seed[out]), 16 entries, conflict-free in shared memory), one funnel shift (33-bit
// rotate in two-word form) and two XORs.  Hashes <= threshold go to the lane's column of a shared-memory
// staging area; when the warp is through with a k, each lane removes the duplicates of its item (the sketch is
// a set), the warp reserves one contiguous region of the batch's dense output with a single atomic and
// writes its 32 items there back to back, coalesced.  Downstream kernels (lookup, vote) thus stream 4 bytes
// per selected hash instead of touching one sparse sector per read.
#include "sq_common.cuh"

namespace sq {

static constexpr int kSketchBlock = 128;
static constexpr int kStageWordsMax = 32 * 24;  // per warp: 32 lanes x (256+~100 bases)/16 words + slack

__device__ __forceinline__ uint32_t items_of(uint32_t L) { return L == 0 ? 1u : (L + SQ_CHUNK - 1) / SQ_CHUNK; }

static constexpr uint32_t kLutWords = 96;  // 48 uint2 per k

__device__ __forceinline__ uint2 lds_v2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t code_at(const uint32_t* wp, uint32_t pos) {
  return (wp[pos >> 4] >> ((pos & 15) * 2)) & 3;
}

// where a lane's selected hashes go: its column of the shared-memory staging area (entry e at sp0 + 128*e; `sp`
// walks it, `n` counts the selected hashes that did not fit any more), or -- for the rare lane that selected
// more than the area holds and rolls a second time -- straight to its reserved place in global memory (`n`
// counts them all)
struct Emitter {
  uint32_t sp, sp_end, n;
  uint32_t* gout;
};
template <bool GLOBAL>
__device__ __forceinline__ void emit_checked(Emitter& E, uint32_t x) {
  if (GLOBAL) {
    E.gout[E.n++] = x;
  } else if (E.sp < E.sp_end) {
    sts_u32(E.sp, x);
    E.sp += 128;
  } else {
    ++E.n;
  }
}
// staging variant without the capacity test: the caller made sure 16 more entries fit
template <bool GLOBAL>
__device__ __forceinline__ void emit_room(Emitter& E, uint32_t x) {
  if (GLOBAL) {
    asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, 4, %0;\n\tst.global.u32 [a], %2;\n\t}" ::"l"(E.gout), "r"(E.n), "r"(x) : "memory");
    ++E.n;
  } else {
    sts_u32(E.sp, x);
    E.sp += 128;
  }
}

// every window end in [c0, c1) of the read at bases [boff, boff+L): hashes <= thr are emitted in window order.
// Neither phase cares where the read starts inside a packed word: the bases are brought to bit 0 with one funnel
// shift per 16 (the word pair is carried from block to block), so every lane of a warp runs the same number of
// unrolled blocks whatever its alignment, and only the last (end - pos) mod 16 steps go one at a time.
template <bool GLOBAL>
__device__ __forceinline__ void roll_item(const uint32_t* wp, uint32_t tb, uint32_t k, uint32_t L, uint32_t boff,
                                          uint32_t c0, uint32_t c1, uint32_t thr, Emitter& E) {
  const uint32_t e_first = max(c0, k - 1);  // first window end that lies in this item
  if (L < k || e_first >= c1) return;
  const uint32_t end = boff + c1;
  const uint32_t last_w = (end - 1) >> 4;  // last packed word this item reads (words up to it belong to the read)
  uint32_t x = 0, y = 0;                   // lane value s: x = s[31:0], y = s[32:1]
  uint32_t pos = boff + e_first - (k - 1);
  // ---- fill the first window (k bases, no output): 16 bases per funnel shift, two bases per table lookup (a
  //      nibble of the packed word), a last single base when the count is odd
  {
    const uint32_t sh = 2 * (pos & 15);
    uint32_t iw = pos >> 4, lo = wp[iw];
    for (uint32_t left = k; left;) {
      const uint32_t hi = iw + 1 <= last_w ? wp[iw + 1] : 0u;
      uint32_t w = __funnelshift_r(lo, hi, sh);
      lo = hi;
      ++iw;
      const uint32_t n = min(left, 16u);
      left -= n;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (2 * i + 1 < (int)n) {  // warp-uniform: k is
          const uint2 d = lds_v2(tb + 128 + ((w >> (4 * i)) & 15) * 8);
          const uint32_t nx = __funnelshift_l(y, x, 2) ^ d.x;
          y = __funnelshift_l(y, x, 1) ^ d.y;
          x = nx;
        }
      if (n & 1) {
        const uint2 d = lds_v2(tb + 256 + ((w >> (2 * (n - 1))) & 3) * 8);
        const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
        y = x ^ d.y;
        x = nx;
      }
    }
    pos += k;
  }
  if (x <= thr) emit_checked<GLOBAL>(E, x);
  if (pos >= end) return;
  // ---- roll: pos = absolute index of the incoming base, the outgoing one is k behind
  const uint32_t sh_i = 2 * (pos & 15), sh_o = 2 * ((pos - k) & 15);
  uint32_t iw = pos >> 4, ow = (pos - k) >> 4;
  uint32_t ilo = wp[iw], olo = wp[ow];
  auto careful = [&](uint32_t n) {  // n <= 16 steps, one at a time, with the staging capacity test
    const uint32_t ihi = iw + 1 <= last_w ? wp[iw + 1] : 0u, ohi = ow + 1 <= last_w ? wp[ow + 1] : 0u;
    uint32_t win = __funnelshift_r(ilo, ihi, sh_i), wout = __funnelshift_r(olo, ohi, sh_o);
    for (uint32_t j = 0; j < n; ++j) {
      const uint2 d = lds_v2(tb + ((win & 3) << 5) + ((wout & 3) << 3));
      const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
      y = x ^ d.y;
      x = nx;
      if (x <= thr) emit_checked<GLOBAL>(E, x);
      win >>= 2;
      wout >>= 2;
    }
    pos += n;
    ++iw;
    ++ow;
    ilo = ihi;
    olo = ohi;
  };
  while (pos + 16 <= end) {
    if (!GLOBAL && E.sp + 16 * 128 > E.sp_end) {  // the staging column may fill up inside this block
      careful(16);
      continue;
    }
    const uint32_t ihi = iw + 1 <= last_w ? wp[iw + 1] : 0u, ohi = wp[ow + 1];  // ow + 1 <= iw: inside the read
    const uint32_t win = __funnelshift_r(ilo, ihi, sh_i), wout = __funnelshift_r(olo, ohi, sh_o);
    ilo = ihi;
    olo = ohi;
    ++iw;
    ++ow;
    const uint32_t xe = ((win & 0x33333333u) << 2) | (wout & 0x33333333u);
    const uint32_t xo = (win & 0xCCCCCCCCu) | ((wout >> 2) & 0x33333333u);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      {
        const uint32_t a = (j == 0 ? (xe << 3) & 0x78u : (xe >> (4 * j - 3)) & 0x78u) | tb;
        const uint2 d = lds_v2(a);
        const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
        y = x ^ d.y;
        x = nx;
        if (x <= thr) emit_room<GLOBAL>(E, x);
      }
      {
        const uint32_t a = (j == 0 ? (xo << 3) & 0x78u : (xo >> (4 * j - 3)) & 0x78u) | tb;
        const uint2 d = lds_v2(a);
        const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
        y = x ^ d.y;
        x = nx;
        if (x <= thr) emit_room<GLOBAL>(E, x);
      }
    }
    pos += 16;
  }
  if (pos < end) careful(end - pos);
}

__global__ void __launch_bounds__(kSketchBlock) sketch_kernel(const __grid_constant__ SketchParams p) {
  extern __shared__ __align__(128) uint32_t smem[];
  // layout: [nk] lookup tables of 384 bytes (128-byte aligned), one read staging area per warp, one output
  // staging area per warp ([entry][lane])
  uint2* lut = reinterpret_cast<uint2*>(smem);
  for (uint32_t i = threadIdx.x; i < p.nk * 48; i += blockDim.x) lut[i] = p.lut[i / 48].e[i % 48];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  // both areas are sized by the launch for the batch at hand (mean read length, scale factor): what a block does not
  // take lets more blocks live on the SM, and this kernel needs the warps to keep its integer pipe fed
  uint32_t* stage = smem + p.nk * kLutWords + warp * p.stage_words;
  uint32_t* ostage = smem + p.nk * kLutWords + (kSketchBlock / 32) * p.stage_words + warp * (p.cap * 32);
  // statistics: warps add their selected-hash counts here, the last one to arrive flushes (no exit barrier)
  __shared__ uint32_t s_sel, s_arrived;
  __shared__ __align__(8) unsigned long long s_bar[kSketchBlock / 32];  // one mbarrier per warp (used once: phase 0)
  if (threadIdx.x == 0) {
    s_sel = 0;
    s_arrived = 0;
    for (int w = 0; w < kSketchBlock / 32; ++w)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar[w])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto arrive = [&](uint32_t v) {  // called by one lane per warp
    if (!p.stats) return;
    if (v) atomicAdd(&s_sel, v);
    __threadfence_block();
    if (atomicAdd(&s_arrived, 1u) == kSketchBlock / 32 - 1) {
      __threadfence_block();
      const uint32_t t = *(volatile uint32_t*)&s_sel;
      if (t) atomicAdd(p.stats, (unsigned long long)t);
    }
  };

  const uint32_t n_items = p.item_start[p.n_reads];
  const uint32_t item = blockIdx.x * kSketchBlock + threadIdx.x;
  const bool valid = item < n_items;
  if (__all_sync(0xFFFFFFFFu, !valid)) {
    if (lane == 0) arrive(0);
    return;
  }

  uint32_t L = 0, boff = 0, c0 = 0, c1 = 0, nit = 1;
  if (valid) {
    const uint32_t r = p.item_read[item];
    const uint32_t ci = item - p.item_start[r];
    L = p.len[r];
    boff = p.base_off[r] - p.bias;
    nit = items_of(L);
    const uint32_t clen = (L + nit - 1) / nit;
    c0 = ci * clen;
    c1 = min(L, c0 + clen);
  }
  // words this lane will touch: bases [max(c0-(kmax-1),0), c1) of its read
  const uint32_t back = min(c0, p.kmax - 1);
  uint32_t wf = valid && c1 > 0 ? (boff + c0 - back) >> 4 : 0xFFFFFFFFu;
  uint32_t wl = valid && c1 > 0 ? (boff + c1 - 1) >> 4 : 0u;
  uint32_t wmin = wf, wmax = wl;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    wmin = min(wmin, __shfl_xor_sync(0xFFFFFFFFu, wmin, d));
    wmax = max(wmax, __shfl_xor_sync(0xFFFFFFFFu, wmax, d));
  }
  const uint32_t* wp = p.packed;
  if (wmin != 0xFFFFFFFFu) {
    const uint32_t base4 = wmin & ~3u;
    const uint32_t nw = wmax - base4 + 1;
    if (((nw + 3) & ~3u) <= p.stage_words) {
      const uint32_t n4 = (nw + 3) >> 2;
      const uint4* src = reinterpret_cast<const uint4*>(p.packed + base4);
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        // the warp's span is one contiguous piece of the packed stream: one bulk copy (TMA engine, no register
        // round trip, no per-lane load/store instructions) that signals the warp's mbarrier when the bytes landed
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[warp]);
        if (lane == 0) {
          const uint32_t bytes = n4 * 16;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           (uint32_t)__cvta_generic_to_shared(stage)),
                       "l"(src), "r"(bytes), "r"(bar)
                       : "memory");
        }
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done)
                       : "r"(bar)
                       : "memory");
      } else {  // a caller's device buffer that is not 16-byte aligned
        uint32_t* dst = stage;
        for (uint32_t i = lane; i < nw; i += 32) dst[i] = __ldg(p.packed + base4 + i);
        __syncwarp();
      }
      wp = stage - base4;  // generic pointer: word index w lives at stage[w - base4]
    }
  }

  const uint32_t thr = p.threshold;
  const uint32_t col = (uint32_t)__cvta_generic_to_shared(ostage) + lane * 4;  // entry e of this lane: col + 128*e
  uint32_t n_sel = 0;
  for (uint32_t ki = 0; ki < p.nk; ++ki) {
    const uint32_t k = p.ks[ki];
    // shared-memory byte address of this k's tables; 128-byte aligned, so (index*8) can be OR-ed in
    const uint32_t tb = (uint32_t)__cvta_generic_to_shared(lut + ki * 48);
    Emitter E;
    E.sp = col;
    E.sp_end = col + p.cap * 128;
    E.n = 0;
    E.gout = nullptr;
    if (valid) roll_item<false>(wp, tb, k, L, boff, c0, c1, thr, E);
    uint32_t c = (E.sp - col) / 128 + E.n;  // staged + those that did not fit
    n_sel += c;
    const bool raw = c > p.cap;  // did not fit: this lane rolls again, straight to global memory
    __syncwarp();
    // ---- the sketch is a set: drop the repeats inside the item (a two-word filter says when to look at all).
    // A read of several items is voted through a per-read set of hashes anyway (the same hash may sit in two of
    // its items): its items keep their repeats.
    if (p.dedup && !raw && c > 1 && nit == 1) {
      if (c <= 32) {
        // pass 1, the same for every lane: which entries MAY repeat an earlier one (both filter bits already set)
        uint32_t m1 = 0, m2 = 0, sus = 0;
        for (uint32_t i = 0; i < c; ++i) {
          const uint32_t h = lds_u32(col + 128 * i);
          const uint32_t b1 = 1u << (h & 31), b2 = 1u << ((h >> 5) & 31);
          if ((m1 & b1) && (m2 & b2)) sus |= 1u << i;
          m1 |= b1;
          m2 |= b2;
        }
        // pass 2: the few suspects are compared with their predecessors, last suspect first, so that the entry a
        // removal moves down (the last one) is already settled.  The warp runs this loop as often as its worst lane
        // has suspects (about once), not once per position at which some lane has one.
        while (sus) {
          const uint32_t i = 31u - (uint32_t)__clz((int)sus);
          sus &= ~(1u << i);
          const uint32_t h = lds_u32(col + 128 * i);
          bool dup = false;
          for (uint32_t j = 0; j < i; ++j) dup |= lds_u32(col + 128 * j) == h;
          if (dup) {  // the last entry takes its place (order inside a set does not matter)
            --c;
            if (i < c) sts_u32(col + 128 * i, lds_u32(col + 128 * c));
          }
        }
      } else {
        uint32_t m1 = 0, m2 = 0;
        for (uint32_t i = 0; i < c; ++i) {
          const uint32_t h = lds_u32(col + 128 * i);
          const uint32_t b1 = 1u << (h & 31), b2 = 1u << ((h >> 5) & 31);
          if ((m1 & b1) && (m2 & b2)) {
            bool dup = false;
            for (uint32_t j = 0; j < i; ++j) dup |= lds_u32(col + 128 * j) == h;
            if (dup) {
              --c;
              if (i < c) sts_u32(col + 128 * i, lds_u32(col + 128 * c));
              --i;
              continue;
            }
          }
          m1 |= b1;
          m2 |= b2;
        }
      }
    }
    // ---- one region of the dense output per warp, the 32 items back to back
    const uint32_t incl = warp_incl_scan(c);
    const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint32_t base = 0;
    if (lane == 0 && tot) base = atomicAdd(p.cursor + ki, tot);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    const uint32_t mine = base + incl - c;
    if (valid) {
      p.hoff[(uint64_t)ki * p.n_items_ub + item] = mine;
      p.cnt[(uint64_t)ki * p.n_items_ub + item] = (uint16_t)(c | (raw && p.dedup ? SQ_CNT_RAW : 0u));
    }
    // each lane copies its own entries: the warp's stores of one round land in one ~1 KB stretch of the region
    // (far fewer instructions than dealing the outputs to the lanes, and this kernel is issue-bound)
    uint32_t* out = p.hsel + (uint64_t)ki * p.hstride;
    if (!raw)
      for (uint32_t i = 0; i < c; ++i) out[mine + i] = lds_u32(col + 128 * i);
    if (raw) {
      E.n = 0;
      E.gout = out + mine;
      roll_item<true>(wp, tb, k, L, boff, c0, c1, thr, E);
    }
    __syncwarp();
  }
  {
    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, n_sel);
    if (lane == 0) arrive(tot);
  }
}

uint32_t sketch_stage_words_max() { return kStageWordsMax; }

size_t sketch_smem_bytes(uint32_t nk, uint32_t cap, uint32_t stage_words) {
  return (nk * kLutWords + (kSketchBlock / 32) * (stage_words + cap * 32)) * sizeof(uint32_t);
}

cudaError_t sketch_configure() {  // per device (sq_create): the staging area may pass the default 48 KB
  return cudaFuncSetAttribute(sketch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sketch_smem_bytes(SQ_MAXK, SQ_CHUNK, kStageWordsMax));
}

void launch_sketch(const SketchParams& p, cudaStream_t s, uint64_t* launches) {
  if (p.n_items_ub == 0) return;
  const uint32_t grid = (p.n_items_ub + kSketchBlock - 1) / kSketchBlock;
  sketch_kernel<<<grid, kSketchBlock, sketch_smem_bytes(p.nk, p.cap, p.stage_words), s>>>(p);
  if (launches) ++*launches;
}

// ---------- items: split reads into chunks of <= SQ_CHUNK window ends ----------

// items per read; also the batch's k-mer count sum_r sum_k max(len_r - k + 1, 0) and base count (stats)
__global__ void __launch_bounds__(256) items_count_kernel(const uint32_t* __restrict__ len, uint32_t n,
                                                          uint32_t* __restrict__ nit, KList ks,
                                                          unsigned long long* __restrict__ stats) {
  // grid-stride: a few hundred blocks, so the two counters see one atomic per block, not one per warp
  __shared__ unsigned long long s_km[8], s_nb[8];
  unsigned long long km = 0, nb = 0;
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const uint32_t L = len[r];
    nit[r] = items_of(L);
    nb += L;
    for (uint32_t i = 0; i < ks.nk; ++i) km += L >= ks.k[i] ? L - ks.k[i] + 1 : 0;
  }
  if (stats) {
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      km += __shfl_xor_sync(0xFFFFFFFFu, km, d);
      nb += __shfl_xor_sync(0xFFFFFFFFu, nb, d);
    }
    if (lane_id() == 0) { s_km[threadIdx.x >> 5] = km; s_nb[threadIdx.x >> 5] = nb; }
    __syncthreads();
    if (threadIdx.x == 0) {
      km = nb = 0;
      for (int w = 0; w < 8; ++w) { km += s_km[w]; nb += s_nb[w]; }
      if (nb) {
        atomicAdd(stats + 0, km);
        atomicAdd(stats + 1, nb);
      }
    }
  }
}

__global__ void items_expand_kernel(const uint32_t* __restrict__ item_start, uint32_t n, uint32_t n_items_ub,
                                    uint32_t* __restrict__ item_read) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t b = item_start[r], e = min(item_start[r + 1], n_items_ub);
  for (uint32_t i = b; i < e; ++i) item_read[i] = r;
}

__global__ void padded_len_kernel(const uint32_t* __restrict__ len, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) out[r] = (len[r] + 3u) & ~3u;
}

// base_off[r] for reads packed back to back, each starting at the next multiple of 4 bases
void launch_derive_offsets(const uint32_t* len, uint32_t n_reads, uint32_t* tmp, uint32_t* base_off, uint32_t* scan_tmp,
                           cudaStream_t s, uint64_t* launches) {
  if (!n_reads) return;
  padded_len_kernel<<<(n_reads + 255) / 256, 256, 0, s>>>(len, n_reads, tmp);
  if (launches) ++*launches;
  launch_exclusive_scan(tmp, base_off, n_reads, scan_tmp, s, launches);
}

// equal-length reads packed back to back, each starting at the next multiple of 4 bases: nothing to scan.  A read
// of up to SQ_CHUNK bases is one item, so the item tables are written here too (item i = read i) and the three
// item kernels are not needed.
__global__ void fixed_layout_kernel(uint32_t* __restrict__ len, uint32_t* __restrict__ base_off, uint32_t n,
                                    uint32_t L, uint32_t stride, uint32_t* __restrict__ item_start,
                                    uint32_t* __restrict__ item_read, KList ks, unsigned long long* __restrict__ stats) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) {
    len[r] = L;
    if (item_read) item_read[r] = r;
  }
  if (r <= n) {
    base_off[r] = r * stride;
    if (item_start) item_start[r] = r;
  }
  if (r == 0 && stats && item_start) {
    unsigned long long km = 0;
    for (uint32_t i = 0; i < ks.nk; ++i) km += L >= ks.k[i] ? L - ks.k[i] + 1 : 0;
    atomicAdd(stats + 0, km * n);
    atomicAdd(stats + 1, (unsigned long long)L * n);
  }
}

// with_items: also write item_start / item_read (read_len <= SQ_CHUNK) and add the batch to the k-mer / base counters
void launch_fixed_layout(uint32_t* len, uint32_t* base_off, uint32_t n_reads, uint32_t read_len, uint32_t* item_start,
                         uint32_t* item_read, const KList& ks, unsigned long long* stats, cudaStream_t s,
                         uint64_t* launches) {
  fixed_layout_kernel<<<n_reads / 256 + 1, 256, 0, s>>>(len, base_off, n_reads, read_len, (read_len + 3u) & ~3u, item_start,
                                                        item_read, ks, stats);
  if (launches) ++*launches;
}

void launch_items(const uint32_t* len, uint32_t n_reads, uint32_t* nit, uint32_t* item_start, uint32_t* item_read,
                  uint32_t n_items_ub, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches, const KList& ks,
                  unsigned long long* stats) {
  if (n_reads == 0) return;
  const uint32_t grid = (n_reads + 255) / 256;
  items_count_kernel<<<grid < 1184 ? grid : 1184, 256, 0, s>>>(len, n_reads, nit, ks, stats);
  if (launches) ++*launches;
  launch_exclusive_scan(nit, item_start, n_reads, scan_tmp, s, launches);
  items_expand_kernel<<<grid, 256, 0, s>>>(item_start, n_reads, n_items_ub, item_read);
  if (launches) ++*launches;
}

}  // namespace sq

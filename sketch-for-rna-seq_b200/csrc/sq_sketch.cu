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
// rotate in two-word form) and two XORs; hashes <= threshold are appended to the item's output slot.
#include "sq_common.cuh"

namespace sq {

static constexpr int kSketchBlock = 128;
static constexpr int kStageWordsPerWarp = 32 * 24;  // 32 lanes x (256+~100 bases)/16 words + slack

__device__ __forceinline__ uint32_t items_of(uint32_t L) { return L == 0 ? 1u : (L + SQ_CHUNK - 1) / SQ_CHUNK; }

// sequential reader of 2-bit codes starting at an arbitrary base position
struct BaseReader {
  const uint32_t* wp;
  uint32_t w;
  uint32_t pos;
  __device__ __forceinline__ void init(const uint32_t* p, uint32_t start) {
    wp = p;
    pos = start;
    w = wp[pos >> 4] >> ((pos & 15) * 2);
  }
  __device__ __forceinline__ uint32_t next() {
    if ((pos & 15) == 0) w = wp[pos >> 4];
    uint32_t c = w & 3;
    w >>= 2;
    ++pos;
    return c;
  }
};

__global__ void __launch_bounds__(kSketchBlock) sketch_kernel(const __grid_constant__ SketchParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  // layout: [nk][20] uint2 lookup tables, then one staging area per warp
  uint2* lut = reinterpret_cast<uint2*>(smem);
  const uint32_t lut_words = p.nk * 40;
  for (uint32_t i = threadIdx.x; i < p.nk * 20; i += blockDim.x) lut[i] = p.lut[i / 20].e[i % 20];
  uint32_t* stage = smem + ((lut_words + 3) & ~3u) + (threadIdx.x >> 5) * kStageWordsPerWarp;
  __syncthreads();

  const uint32_t n_items = p.item_start[p.n_reads];
  const uint32_t item = blockIdx.x * kSketchBlock + threadIdx.x;
  const bool valid = item < n_items;
  if (__all_sync(0xFFFFFFFFu, !valid)) return;

  uint32_t L = 0, boff = 0, c0 = 0, c1 = 0;
  if (valid) {
    const uint32_t r = p.item_read[item];
    const uint32_t ci = item - p.item_start[r];
    L = p.len[r];
    boff = p.base_off[r] - p.bias;
    const uint32_t nit = items_of(L);
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
    if (nw <= (uint32_t)kStageWordsPerWarp) {
      const uint32_t n4 = (nw + 3) >> 2;
      const uint4* src = reinterpret_cast<const uint4*>(p.packed + base4);
      uint4* dst = reinterpret_cast<uint4*>(stage);
      for (uint32_t i = lane_id(); i < n4; i += 32) dst[i] = __ldg(src + i);
      __syncwarp();
      wp = stage - base4;  // generic pointer: word index w lives at stage[w - base4]
    }
  }
  if (!valid) return;

  const uint32_t thr = p.threshold;
  for (uint32_t ki = 0; ki < p.nk; ++ki) {
    const uint32_t k = p.ks[ki];
    const uint2* lk = lut + ki * 20;
    uint32_t n_out = 0;
    uint32_t* out = p.sel + (uint64_t)ki * p.slot_stride + boff + c0;
    const uint32_t e_first = max(c0, k - 1);  // first window end that lies in this item
    if (L >= k && e_first < c1) {
      // fill the first window: bases e_first-(k-1) .. e_first
      BaseReader in;
      in.init(wp, boff + e_first - (k - 1));
      uint32_t x = 0, y = 0;  // lane value s: x = s[31:0], y = s[32:1]
      for (uint32_t j = 0; j < k; ++j) {
        const uint2 d = lk[16 + in.next()];
        const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
        y = x ^ d.y;
        x = nx;
      }
      if (x <= thr) out[n_out++] = x;
      uint32_t e = e_first + 1;
      BaseReader ob;
      ob.init(wp, boff + e - k);
      // scalar steps until the incoming base is word aligned
      while (e < c1 && ((boff + e) & 15) != 0) {
        const uint2 d = lk[in.next() * 4 + ob.next()];
        const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
        y = x ^ d.y;
        x = nx;
        if (x <= thr) out[n_out++] = x;
        ++e;
      }
      // aligned blocks of 16 steps: one incoming word, the outgoing stream re-aligned by a funnel shift
      const uint32_t q = k >> 4, dr = k & 15;
      const char* lkb = reinterpret_cast<const char*>(lk);
      while (e + 16 <= c1) {
        const uint32_t iw = (boff + e) >> 4;
        const uint32_t win = wp[iw];
        uint32_t wout = wp[iw - q];
        if (dr) wout = __funnelshift_r(wp[iw - q - 1], wout, 32 - 2 * dr);
        const uint32_t xe = ((win & 0x33333333u) << 2) | (wout & 0x33333333u);
        const uint32_t xo = (win & 0xCCCCCCCCu) | ((wout >> 2) & 0x33333333u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          {
            const uint32_t a = j == 0 ? (xe << 3) & 0x78u : (xe >> (4 * j - 3)) & 0x78u;
            const uint2 d = *reinterpret_cast<const uint2*>(lkb + a);
            const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
            y = x ^ d.y;
            x = nx;
            if (x <= thr) out[n_out++] = x;
          }
          {
            const uint32_t a = j == 0 ? (xo << 3) & 0x78u : (xo >> (4 * j - 3)) & 0x78u;
            const uint2 d = *reinterpret_cast<const uint2*>(lkb + a);
            const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
            y = x ^ d.y;
            x = nx;
            if (x <= thr) out[n_out++] = x;
          }
        }
        e += 16;
      }
      if (e < c1) {
        in.init(wp, boff + e);
        ob.init(wp, boff + e - k);
        while (e < c1) {
          const uint2 d = lk[in.next() * 4 + ob.next()];
          const uint32_t nx = __funnelshift_l(y, x, 1) ^ d.x;
          y = x ^ d.y;
          x = nx;
          if (x <= thr) out[n_out++] = x;
          ++e;
        }
      }
    }
    p.cnt[(uint64_t)ki * p.n_items_ub + item] = (uint16_t)n_out;
  }
}

void launch_sketch(const SketchParams& p, cudaStream_t s, uint64_t* launches) {
  if (p.n_items_ub == 0) return;
  const uint32_t lut_words = (p.nk * 40 + 3) & ~3u;
  const size_t smem = (lut_words + (kSketchBlock / 32) * kStageWordsPerWarp) * sizeof(uint32_t);
  const uint32_t grid = (p.n_items_ub + kSketchBlock - 1) / kSketchBlock;
  sketch_kernel<<<grid, kSketchBlock, smem, s>>>(p);
  if (launches) ++*launches;
}

// ---------- items: split reads into chunks of <= SQ_CHUNK window ends ----------

__global__ void items_count_kernel(const uint32_t* __restrict__ len, uint32_t n, uint32_t* __restrict__ nit) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) nit[r] = items_of(len[r]);
}

__global__ void items_expand_kernel(const uint32_t* __restrict__ item_start, uint32_t n, uint32_t n_items_ub,
                                    uint32_t* __restrict__ item_read) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t b = item_start[r], e = min(item_start[r + 1], n_items_ub);
  for (uint32_t i = b; i < e; ++i) item_read[i] = r;
}

void launch_items(const uint32_t* len, uint32_t n_reads, uint32_t* nit, uint32_t* item_start, uint32_t* item_read,
                  uint32_t n_items_ub, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches) {
  if (n_reads == 0) return;
  const uint32_t grid = (n_reads + 255) / 256;
  items_count_kernel<<<grid, 256, 0, s>>>(len, n_reads, nit);
  if (launches) ++*launches;
  launch_exclusive_scan(nit, item_start, n_reads, scan_tmp, s, launches);
  items_expand_kernel<<<grid, 256, 0, s>>>(item_start, n_reads, n_items_ub, item_read);
  if (launches) ++*launches;
}

}  // namespace sq

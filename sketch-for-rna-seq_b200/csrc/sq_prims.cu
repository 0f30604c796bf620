// Device primitives used by the quant path: exclusive scan (u32) and a stable LSD radix sort of
// (key64, val32) pairs.  Hand-written for sm_100a; no CUB/Thrust.
#include "sq_common.cuh"

namespace sq {

// ------------------------------------------------------------------ scan
static constexpr int kScanThreads = 256;
static constexpr int kScanItems = 16;
static constexpr int kScanTile = kScanThreads * kScanItems;  // 2048

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  // exclusive scan of one value per thread over a 256-thread block
  __shared__ uint32_t wsum[kScanThreads / 32];
  __shared__ uint32_t wtot;
  const uint32_t incl = warp_incl_scan(v);
  if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = threadIdx.x < kScanThreads / 32 ? wsum[threadIdx.x] : 0;
    const uint32_t wi = warp_incl_scan(w);
    if (threadIdx.x < kScanThreads / 32) wsum[threadIdx.x] = wi - w;
    if (threadIdx.x == kScanThreads / 32 - 1) wtot = wi;
  }
  __syncthreads();
  const uint32_t r = wsum[threadIdx.x >> 5] + incl - v;
  *total = wtot;
  __syncthreads();
  return r;
}

// Single-pass scan with decoupled look-back: tiles take a ticket (so a tile only ever waits for tiles that have
// already started), publish their aggregate, and the first warp walks back over the predecessors' status words
// 32 at a time until it meets an inclusive prefix.  status: one 64-bit word per tile, bit 63 = inclusive prefix
// available, bit 62 = aggregate available, low 32 bits = value; zeroed (with the ticket) before every scan.
static constexpr unsigned long long kScanIncl = 1ull << 63, kScanAgg = 1ull << 62;

template <bool VEC>  // VEC: in and out are 16-byte aligned, whole tiles move as uint4
__global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(const uint32_t* __restrict__ in, uint32_t n,
                                                                     uint32_t* __restrict__ out,
                                                                     unsigned long long* status, uint32_t* ticket) {
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
  const bool whole = VEC && base + kScanItems <= n;
  if (whole) {
#pragma unroll
    for (int i = 0; i < kScanItems; i += 4) {
      const uint4 q = *reinterpret_cast<const uint4*>(in + base + i);
      v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) sum += v[i];
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      v[i] = base + i < n ? in[base + i] : 0;
      sum += v[i];
    }
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(sum, &tot);
  if (threadIdx.x < 32) {
    const uint32_t lane = threadIdx.x;
    uint32_t prefix = 0;
    if (tile == 0) {
      if (lane == 0) {
        *(volatile unsigned long long*)&status[0] = kScanIncl | tot;
      }
    } else {
      if (lane == 0) {
        *(volatile unsigned long long*)&status[tile] = kScanAgg | tot;
      }
      __threadfence();
      int64_t hi = (int64_t)tile - 1;  // nearest predecessor not yet accounted for
      for (;;) {
        const int64_t j = hi - lane;
        unsigned long long w = kScanIncl;  // tiles before 0 act as an inclusive prefix of 0
        if (j >= 0) {
          do { w = *(volatile unsigned long long*)&status[j]; } while ((w & (kScanIncl | kScanAgg)) == 0);
        }
        const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, (w & kScanIncl) != 0);
        const int first = incl_mask ? __ffs(incl_mask) - 1 : 32;  // closest inclusive prefix in the window
        uint32_t contrib = (int)lane <= first ? (uint32_t)w : 0u;
#pragma unroll
        for (int d = 16; d; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
        prefix += contrib;
        if (incl_mask) break;
        hi -= 32;
      }
      if (lane == 0) {
        *(volatile unsigned long long*)&status[tile] = kScanIncl | (unsigned long long)(prefix + tot);
      }
    }
    if (lane == 0) s_prefix = prefix;
  }
  __syncthreads();
  ex += s_prefix;
  if (whole) {
#pragma unroll
    for (int i = 0; i < kScanItems; i += 4) {
      uint4 q;
      q.x = ex; ex += v[i];
      q.y = ex; ex += v[i + 1];
      q.z = ex; ex += v[i + 2];
      q.w = ex; ex += v[i + 3];
      *reinterpret_cast<uint4*>(out + base + i) = q;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
      if (base + i < n) out[base + i] = ex;
      ex += v[i];
    }
  }
  if (base <= n - 1 && n - 1 < base + kScanItems) out[n] = ex;  // the thread holding the last element closes the scan
}

size_t scan_tmp_words(uint32_t n) { return 2 * ((size_t)(n / kScanTile) + 2) + 4; }

void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t s,
                           uint64_t* launches) {
  if (n == 0) {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
    return;
  }
  const uint32_t nb = (n + kScanTile - 1) / kScanTile;
  // tmp: [0] ticket (u32, padded to 8 bytes), then nb status words
  cudaMemsetAsync(tmp, 0, 8 + (size_t)nb * 8, s);
  if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0)
    scan_lookback_kernel<true><<<nb, kScanThreads, 0, s>>>(in, n, out, reinterpret_cast<unsigned long long*>(tmp + 2), tmp);
  else
    scan_lookback_kernel<false><<<nb, kScanThreads, 0, s>>>(in, n, out, reinterpret_cast<unsigned long long*>(tmp + 2), tmp);
  if (launches) *launches += 1;
}

// ------------------------------------------------------------------ radix sort
// 8-bit digits, least significant first.  Each pass builds per-tile digit histograms, scans them
// (digit-major) and scatters tile by tile.  Inside a tile keys are
// ranked stably (warp match + per-warp digit counters), staged in shared memory in digit order and written
// out in runs, so global writes are coalesced per digit.
static constexpr int kSortThreads = 256;
static constexpr int kSortItems = 8;
static constexpr int kSortTile = kSortThreads * kSortItems;  // 2048
static constexpr int kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n,
                                                                  int shift, uint32_t ntiles,
                                                                  uint32_t* __restrict__ hist) {
  // per-tile digit histogram of the CURRENT key order (tiles change content after every pass);
  // hist layout: [digit][tile]
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * kSortTile;
  uint64_t k[kSortItems];  // all loads of the thread in flight before the first shared-memory atomic
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint64_t idx = base + (uint64_t)i * kSortThreads + threadIdx.x;
    k[i] = idx < n ? keys[idx] : 0ull;
  }
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint64_t idx = base + (uint64_t)i * kSortThreads + threadIdx.x;
    if (idx < n) atomicAdd(&h[(k[i] >> shift) & 255], 1u);
  }
  __syncthreads();
  hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <bool VALS>  // keys-only sorts carry no payload registers or staging
__global__ void __launch_bounds__(kSortThreads, VALS ? 4 : 5) radix_scatter_kernel(const uint64_t* __restrict__ kin,
                                                                     const uint32_t* __restrict__ vin,
                                                                     uint64_t* __restrict__ kout,
                                                                     uint32_t* __restrict__ vout, uint64_t n,
                                                                     int shift, uint32_t ntiles,
                                                                     const uint32_t* __restrict__ offs) {
  // offs: exclusive scan of this pass's [digit][tile] histogram
  __shared__ uint32_t wcnt[kSortWarps][256];  // per-warp digit counters -> exclusive over warps
  __shared__ uint32_t dstart[256];            // tile-local start of each digit
  __shared__ uint32_t gbase[256];             // global start of this tile's run of each digit
  __shared__ uint64_t skey[kSortTile];
  __shared__ uint32_t sval[VALS ? kSortTile : 1];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const uint64_t tbase = (uint64_t)blockIdx.x * kSortTile;
  // element order inside the tile: warp-major, then round, then lane (this defines stability)
  uint64_t key[kSortItems];
  uint32_t val[VALS ? kSortItems : 1];
  uint32_t rank[kSortItems];
  const uint64_t wbase = tbase + (uint64_t)warp * (kSortItems * 32);
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint64_t idx = wbase + i * 32 + lane;
    const bool ok = idx < n;
    key[i] = ok ? kin[idx] : ~0ull;
    if (VALS) val[i] = ok ? vin[idx] : 0u;
    const uint32_t d = ok ? (uint32_t)(key[i] >> shift) & 255u : 256u;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
    const uint32_t before = __popc(peers & ((1u << lane) - 1));
    uint32_t old = 0;
    if (ok && before == 0) {  // lowest lane of the group owns the counter update
      old = wcnt[warp][d];
      wcnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xFFFFFFFFu, old, __ffs(peers) - 1);
    rank[i] = old + before;
    __syncwarp();
  }
  __syncthreads();
  {  // thread d: exclusive prefix over warps, tile totals
    const uint32_t d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = wcnt[w][d];
      wcnt[w][d] = run;
      run += c;
    }
    uint32_t tot;
    const uint32_t ex = block_excl_scan(run, &tot);
    dstart[d] = ex;
    gbase[d] = offs[(uint64_t)d * ntiles + blockIdx.x];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kSortItems; ++i) {
    const uint64_t idx = wbase + i * 32 + lane;
    if (idx < n) {
      const uint32_t d = (uint32_t)(key[i] >> shift) & 255u;
      const uint32_t pos = dstart[d] + wcnt[warp][d] + rank[i];
      skey[pos] = key[i];
      if (VALS) sval[pos] = val[i];
    }
  }
  __syncthreads();
  const uint32_t cnt = (uint32_t)min((uint64_t)kSortTile, n - tbase);
  for (uint32_t i = threadIdx.x; i < cnt; i += kSortThreads) {
    const uint64_t k = skey[i];
    const uint32_t d = (uint32_t)(k >> shift) & 255u;
    const uint64_t dst = (uint64_t)gbase[d] + (i - dstart[d]);
    kout[dst] = k;
    if (VALS) vout[dst] = sval[i];
  }
}

size_t radix_tmp_words(uint64_t n) {
  const uint64_t ntiles = (n + kSortTile - 1) / kSortTile;
  // one pass's histogram + its scan (+1) + scan scratch
  return (size_t)(256 * ntiles) + (size_t)(256 * ntiles + 2) + scan_tmp_words((uint32_t)(256 * ntiles)) + 16;
}

void launch_radix_sort(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint64_t n, int nbits,
                       uint32_t* tmp, uint64_t** keys_out, uint32_t** vals_out, cudaStream_t s,
                       uint64_t* launches, int bit_lo) {
  *keys_out = keys_a;
  *vals_out = vals_a;
  if (n == 0 || nbits <= 0) return;
  const int npass = (nbits + 7) / 8;
  const uint32_t ntiles = (uint32_t)((n + kSortTile - 1) / kSortTile);
  uint32_t* hist = tmp;
  uint32_t* offs = hist + (size_t)256 * ntiles;
  uint32_t* scan_tmp = offs + (size_t)256 * ntiles + 2;  // keeps the 8-byte alignment the scan's status words need
  uint64_t* kin = keys_a;
  uint64_t* kout = keys_b;
  uint32_t* vin = vals_a;
  uint32_t* vout = vals_b;
  for (int p = 0; p < npass; ++p) {
    radix_hist_kernel<<<ntiles, kSortThreads, 0, s>>>(kin, n, bit_lo + 8 * p, ntiles, hist);
    if (launches) ++*launches;
    launch_exclusive_scan(hist, offs, 256 * ntiles, scan_tmp, s, launches);
    if (vin && vout)
      radix_scatter_kernel<true><<<ntiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, bit_lo + 8 * p, ntiles, offs);
    else
      radix_scatter_kernel<false><<<ntiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, bit_lo + 8 * p, ntiles, offs);
    if (launches) ++*launches;
    uint64_t* tk = kin; kin = kout; kout = tk;
    uint32_t* tv = vin; vin = vout; vout = tv;
  }
  *keys_out = kin;
  *vals_out = vin;
}

}  // namespace sq

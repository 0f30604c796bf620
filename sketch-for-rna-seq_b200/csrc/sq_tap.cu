// Gather kernels behind sq_sketch() (debug tap) and sq_build_postings() (index construction: the inverted
// map of reference src/sketch.cpp:51-74 built from per-sequence sketches).
#include "sq_common.cuh"
#include "sq_kernels.cuh"
#include "sq_tap.cuh"

namespace sq {

__device__ __forceinline__ uint32_t items_of(uint32_t L) { return L == 0 ? 1u : (L + SQ_CHUNK - 1) / SQ_CHUNK; }

// counts[r*nk+ki] = selected k-mers of read r for k-index ki (sum over the read's items)
__global__ void tap_count_kernel(const uint32_t* __restrict__ item_start, uint32_t n_reads, uint32_t nk,
                                 const uint16_t* __restrict__ cnt, uint32_t n_items_ub, uint32_t* __restrict__ counts) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_reads * nk) return;
  const uint32_t r = (uint32_t)(i / nk), ki = (uint32_t)(i % nk);
  uint32_t s = 0;
  for (uint32_t it = item_start[r]; it < item_start[r + 1]; ++it) s += cnt[(uint64_t)ki * n_items_ub + it] & SQ_CNT_MASK;
  counts[i] = s;
}

// copy the selected hashes of (r, ki) to out[offs[r*nk+ki] ..]; when keys64 != NULL write
// (hash << tbits | seq_tid[r]) instead (input of the postings sort)
__global__ void tap_gather_kernel(const uint32_t* __restrict__ item_start, uint32_t n_reads, uint32_t nk,
                                  const uint16_t* __restrict__ cnt, uint32_t n_items_ub,
                                  const uint32_t* __restrict__ hsel, uint64_t hstride, const uint32_t* __restrict__ hoff,
                                  const uint32_t* __restrict__ offs, uint64_t cap, uint32_t* __restrict__ out,
                                  uint64_t* __restrict__ keys64, const uint32_t* __restrict__ seq_tid, uint32_t tbits) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= (uint64_t)n_reads * nk) return;
  const uint32_t r = (uint32_t)(i / nk), ki = (uint32_t)(i % nk);
  const uint32_t item0 = item_start[r], n_it = item_start[r + 1] - item0;
  uint64_t o = offs[i];
  for (uint32_t it = 0; it < n_it; ++it) {
    const uint32_t c = cnt[(uint64_t)ki * n_items_ub + item0 + it] & SQ_CNT_MASK;
    const uint32_t* src = hsel + (uint64_t)ki * hstride + hoff[(uint64_t)ki * n_items_ub + item0 + it];
    for (uint32_t j = 0; j < c; ++j, ++o) {
      if (o >= cap) continue;
      if (keys64) keys64[o] = ((uint64_t)src[j] << tbits) | seq_tid[r];
      else out[o] = src[j];
    }
  }
}

void launch_tap_count(const uint32_t* item_start, uint32_t n_reads, uint32_t nk, const uint16_t* cnt,
                      uint32_t n_items_ub, uint32_t* counts, cudaStream_t s, uint64_t* launches) {
  const uint64_t n = (uint64_t)n_reads * nk;
  if (!n) return;
  tap_count_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(item_start, n_reads, nk, cnt, n_items_ub, counts);
  if (launches) ++*launches;
}

void launch_tap_gather(const uint32_t* item_start, uint32_t n_reads, uint32_t nk, const uint16_t* cnt,
                       uint32_t n_items_ub, const uint32_t* hsel, uint64_t hstride, const uint32_t* hoff,
                       const uint32_t* offs, uint64_t cap, uint32_t* out, uint64_t* keys64, const uint32_t* seq_tid,
                       uint32_t tbits, cudaStream_t s, uint64_t* launches) {
  const uint64_t n = (uint64_t)n_reads * nk;
  if (!n) return;
  tap_gather_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(item_start, n_reads, nk, cnt, n_items_ub, hsel, hstride,
                                                                hoff, offs, cap, out, keys64, seq_tid, tbits);
  if (launches) ++*launches;
}

// sorted (hash<<tbits|tid) keys -> boundary flags
__global__ void post_flags_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t tbits,
                                  uint32_t* __restrict__ newpair, uint32_t* __restrict__ newkey) {
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint64_t k = keys[j];
  const bool np = j == 0 || k != keys[j - 1];
  const bool nk = j == 0 || (k >> tbits) != (keys[j - 1] >> tbits);
  newpair[j] = np;
  newkey[j] = nk;
}

__global__ void post_scatter_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t tbits,
                                    const uint32_t* __restrict__ newpair, const uint32_t* __restrict__ newkey,
                                    const uint32_t* __restrict__ ppos, const uint32_t* __restrict__ kpos,
                                    uint32_t* __restrict__ out_keys, unsigned long long* __restrict__ out_off,
                                    uint32_t* __restrict__ out_tid) {
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j > n) return;
  if (j == n) {
    out_off[kpos[n]] = ppos[n];
    return;
  }
  const uint64_t k = keys[j];
  if (newpair[j]) out_tid[ppos[j]] = (uint32_t)(k & ((1ull << tbits) - 1));
  if (newkey[j]) {
    out_keys[kpos[j]] = (uint32_t)(k >> tbits);
    out_off[kpos[j]] = ppos[j];
  }
}

void launch_post_flags(const uint64_t* keys, uint64_t n, uint32_t tbits, uint32_t* newpair, uint32_t* newkey,
                       cudaStream_t s, uint64_t* launches) {
  if (!n) return;
  post_flags_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(keys, n, tbits, newpair, newkey);
  if (launches) ++*launches;
}

void launch_post_scatter(const uint64_t* keys, uint64_t n, uint32_t tbits, const uint32_t* newpair,
                         const uint32_t* newkey, const uint32_t* ppos, const uint32_t* kpos, uint32_t* out_keys,
                         unsigned long long* out_off, uint32_t* out_tid, cudaStream_t s, uint64_t* launches) {
  post_scatter_kernel<<<(uint32_t)((n + 1 + 255) / 256), 256, 0, s>>>(keys, n, tbits, newpair, newkey, ppos, kpos,
                                                                      out_keys, out_off, out_tid);
  if (launches) ++*launches;
}

}  // namespace sq

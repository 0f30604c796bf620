// Read classes over the candidate store, the EM / assignment kernels and their exchange between GPUs.
//
// EM follows estimate_isoform_abundance_em (reference src/isoform_assignment.cpp:9-68) and
// assign_reads_to_isoforms (:70-97) on a flat CSR of (read -> candidates).  Every reduction has a fixed
// order (no floating-point atomics): the E-step denominator is a per-read sequential sum in candidate
// order; the per-transcript posterior sums run over a transcript-major copy of the pairs (stable radix
// sort), split into fixed-size segments that are reduced by one warp each and then added in segment order.
#include "sq_common.cuh"
#include "sq_kernels.cuh"

namespace sq {

static constexpr uint32_t kHashMul = 0x9E3779B1u;

// ------------------------------------------------------------------ index table
// key bitmap with interleaved ranks (see IndexTable): sector s = {keys below s*224, bits of keys s*224 .. s*224+223}
__global__ void bmap_set_kernel(const uint32_t* __restrict__ keys, uint64_t nkeys, uint32_t* __restrict__ words) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nkeys) return;
  const uint32_t key = keys[i], sec = key / SQ_BMAP_BITS, r = key - sec * SQ_BMAP_BITS;
  atomicOr(words + (size_t)sec * 8 + 1 + (r >> 5), 1u << (r & 31));
}
__global__ void bmap_count_kernel(const uint32_t* __restrict__ words, uint32_t n_sectors, uint32_t* __restrict__ cnt) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_sectors) return;
  uint32_t c = 0;
#pragma unroll
  for (int j = 1; j < 8; ++j) c += __popc(words[(size_t)s * 8 + j]);
  cnt[s] = c;
}
__global__ void bmap_rank_kernel(uint32_t* __restrict__ words, uint32_t n_sectors, const uint32_t* __restrict__ excl) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_sectors) words[(size_t)s * 8] = excl[s];
}

// distinct keys (device) -> bitmap + ranks; cnt/excl/scan_tmp are scratch of n_sectors+1 / scan_tmp_words(n_sectors)
void launch_bmap_build(const uint32_t* keys, uint64_t nkeys, uint4* bmap, uint32_t n_sectors, uint32_t* cnt,
                       uint32_t* excl, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches) {
  cudaMemsetAsync(bmap, 0, (size_t)n_sectors * 32, s);
  uint32_t* words = reinterpret_cast<uint32_t*>(bmap);
  if (nkeys) {
    bmap_set_kernel<<<(uint32_t)((nkeys + 255) / 256), 256, 0, s>>>(keys, nkeys, words);
    if (launches) ++*launches;
  }
  bmap_count_kernel<<<(n_sectors + 255) / 256, 256, 0, s>>>(words, n_sectors, cnt);
  launch_exclusive_scan(cnt, excl, n_sectors, scan_tmp, s, launches);
  bmap_rank_kernel<<<(n_sectors + 255) / 256, 256, 0, s>>>(words, n_sectors, excl);
  if (launches) *launches += 2;
}

// results leave the engine in the caller's transcript numbering: out[ext_of[i]] = in[i]
__global__ void permute_out_kernel(const double* __restrict__ pi, const double* __restrict__ numreads,
                                   const uint32_t* __restrict__ present, const uint32_t* __restrict__ ext_of,
                                   uint32_t T, double* __restrict__ pi_out, double* __restrict__ nr_out,
                                   uint8_t* __restrict__ present_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const uint32_t x = ext_of ? ext_of[i] : i;
  pi_out[x] = pi[i];
  nr_out[x] = numreads[i];
  present_out[x] = present[i] ? 1 : 0;
}

void launch_permute_out(const double* pi, const double* numreads, const uint32_t* present, const uint32_t* ext_of,
                        uint32_t T, double* pi_out, double* nr_out, uint8_t* present_out, cudaStream_t s,
                        uint64_t* launches) {
  permute_out_kernel<<<(T + 255) / 256, 256, 0, s>>>(pi, numreads, present, ext_of, T, pi_out, nr_out, present_out);
  if (launches) ++*launches;
}

// ------------------------------------------------------------------ candidate store and read classes
// The vote kernels write a read's candidates straight into the engine's store (one atomic cursor; rd[r] = {start, count}
// say where): nothing is staged, scanned or moved afterwards.  EM does not care which read is which: reads with
// the same candidate list (same transcripts, same scores) contribute identical terms, so they are collapsed into
// one class with a weight.  Behind every batch's vote a kernel folds a 128-bit fingerprint of each
// list -- two independent 64-bit non-cryptographic hashes over length, transcripts and scores -- and the read's
// class sort key (best candidate, a few hash bits, read index).  sq_finish sorts the keys; a read starts a class
// when its list differs from its predecessor's: by fingerprint (an accidental 128-bit match of two different
// lists would merge two EM terms and nothing else), or element-wise with option exact_classes (~10 random
// sectors per read, 2.8 ms at 20 M reads).  Ordering by best candidate keeps the classes of one gene adjacent,
// which makes the 1/den gathers of the transcript-major pass local.  Summing w identical terms becomes one
// multiplication by w: a re-association only.
__global__ void __launch_bounds__(256) read_keys_kernel(const uint2* __restrict__ rd, uint64_t r0, uint32_t n,
                                                        const uint2* __restrict__ cand, uint32_t T,
                                                        uint32_t hash_bits, uint64_t* __restrict__ rkey,
                                                        ulonglong2* __restrict__ rfp) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t r = r0 + i;
  const uint32_t c = rd[r].y, so = rd[r].x;
  ListHash lh;
  lh.init(c);
  uint32_t top = T;  // reads without candidates go last (one empty class)
  // blocks of 4 candidates: the four loads go out together (the kernel waits on latency, the fold is serial)
  for (uint32_t i0 = 0; i0 < c; i0 += 4) {
    uint32_t t[4];
    int32_t sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint2 pr = i0 + u < c ? cand[so + i0 + u] : make_uint2(0u, 0u);
      t[u] = pr.x;
      sc[u] = (int32_t)pr.y;
    }
    if (i0 == 0) top = t[0];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < c) lh.add(t[u], sc[u]);
  }
  rkey[r] = lh.key(top, hash_bits, r);
  rfp[r] = make_ulonglong2(lh.h, lh.g);
}

void launch_read_keys(const uint2* rd, uint64_t r0, uint64_t n, const uint2* cand,
                      uint32_t T, uint32_t hash_bits, uint64_t* rkey, void* rfp, cudaStream_t s, uint64_t* launches) {
  if (!n) return;
  read_keys_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(rd, r0, (uint32_t)n, cand, T, hash_bits, rkey,
                                                               static_cast<ulonglong2*>(rfp));
  if (launches) ++*launches;
}

// the store in read order (sq_get_candidates): off = exclusive scan of the counts
__global__ void rd_counts_kernel(const uint2* __restrict__ rd, uint64_t n_reads, uint32_t* __restrict__ cnt) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r < n_reads) cnt[r] = rd[r].y;
}

__global__ void csr_gather_kernel(const uint2* __restrict__ rd,
                                  const uint32_t* __restrict__ off, uint64_t n_reads,
                                  const uint2* __restrict__ cand, uint32_t* __restrict__ out_tid,
                                  int32_t* __restrict__ out_score) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t b = rd[r].x, n = rd[r].y, d = off[r];
  for (uint32_t j = 0; j < n; ++j) {
    const uint2 pr = cand[b + j];
    out_tid[d + j] = pr.x;
    out_score[d + j] = (int32_t)pr.y;
  }
}

void launch_csr_gather(const uint2* rd, uint32_t* cnt_tmp, uint32_t* off, uint64_t n_reads,
                       uint32_t* scan_tmp, const uint2* cand, uint32_t* out_tid, int32_t* out_score, cudaStream_t s,
                       uint64_t* launches) {
  if (n_reads) rd_counts_kernel<<<(uint32_t)((n_reads + 255) / 256), 256, 0, s>>>(rd, n_reads, cnt_tmp);
  launch_exclusive_scan(cnt_tmp, off, (uint32_t)n_reads, scan_tmp, s, launches);
  if (!n_reads) return;
  csr_gather_kernel<<<(uint32_t)((n_reads + 255) / 256), 256, 0, s>>>(rd, off, n_reads, cand, out_tid,
                                                                      out_score);
  if (launches) ++*launches;
}

// Four sorted positions per thread: the four fingerprint gathers (random 16-byte reads, the cost of this kernel) are
// in flight together, and so is the one extra (key, fingerprint) the warp's first lane needs for the position before
// its own; positions compare with their left neighbour in registers, the first one across lanes.
__global__ void __launch_bounds__(256) class_head_kernel(const uint64_t* __restrict__ keys, uint64_t n_reads,
                                                         const ulonglong2* __restrict__ fp, uint32_t* __restrict__ head) {
  const uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 4;
  uint64_t k[4], pk = 0;
  ulonglong2 f[4], pf = make_ulonglong2(0, 0);
  const bool first_lane = lane_id() == 0 && i0 > 0 && i0 < n_reads;
  if (first_lane) pk = keys[i0 - 1];
#pragma unroll
  for (int u = 0; u < 4; ++u) k[u] = i0 + u < n_reads ? keys[i0 + u] : 0ull;
  if (first_lane) pf = fp[(uint32_t)pk];
#pragma unroll
  for (int u = 0; u < 4; ++u) f[u] = i0 + u < n_reads ? fp[(uint32_t)k[u]] : make_ulonglong2(0, 0);
  const uint64_t sk = __shfl_up_sync(0xFFFFFFFFu, k[3], 1);
  const unsigned long long sx = __shfl_up_sync(0xFFFFFFFFu, f[3].x, 1), sy = __shfl_up_sync(0xFFFFFFFFu, f[3].y, 1);
  if (lane_id() != 0) { pk = sk; pf.x = sx; pf.y = sy; }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (i0 + u < n_reads)
      head[i0 + u] = (i0 + u == 0 || (k[u] >> 32) != (pk >> 32) || f[u].x != pf.x || f[u].y != pf.y) ? 1u : 0u;
    pk = k[u];
    pf = f[u];
  }
}

// exact variant (option exact_classes): equal key AND element-wise equal lists; ~10 random sectors per read
__global__ void class_head_exact_kernel(const uint64_t* __restrict__ keys, uint64_t n_reads,
                                        const uint2* __restrict__ rd, const uint2* __restrict__ cand,
                                        uint32_t* __restrict__ head) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  uint32_t h = 1;
  if (i > 0 && (keys[i] >> 32) == (keys[i - 1] >> 32)) {
    const uint32_t r = (uint32_t)keys[i], q = (uint32_t)keys[i - 1];
    const uint32_t b = rd[r].x, n = rd[r].y, bq = rd[q].x;
    if (rd[q].y == n) {
      h = 0;
      for (uint32_t j = 0; j < n; ++j) {
        const uint2 x = cand[b + j], y = cand[bq + j];
        if (x.x != y.x || x.y != y.y) { h = 1; break; }
      }
    }
  }
  head[i] = h;
}

// class c = run of sorted positions starting at a head: its list is the head's, its weight the run length
__global__ void class_fill_kernel(const uint32_t* __restrict__ head, const uint32_t* __restrict__ cid,
                                  const uint64_t* __restrict__ keys, uint64_t n_reads,
                                  const uint2* __restrict__ rd, uint32_t* __restrict__ class_read,
                                  uint32_t* __restrict__ class_pos, uint32_t* __restrict__ class_cnt) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  if (head[i]) {
    const uint32_t c = cid[i], r = (uint32_t)keys[i];
    class_read[c] = r;
    class_pos[c] = (uint32_t)i;
    class_cnt[c] = rd[r].y;
  }
  if (i == n_reads - 1) class_pos[cid[n_reads]] = (uint32_t)n_reads;  // cid[n_reads] = number of classes
}

__global__ void class_gather_kernel(const uint32_t* __restrict__ class_read, const uint32_t* __restrict__ class_pos,
                                    const uint32_t* __restrict__ class_off, uint32_t n_classes,
                                    const uint2* __restrict__ rd, const uint2* __restrict__ cand,
                                    uint32_t* __restrict__ out_tid,
                                    int32_t* __restrict__ out_score,
                                    uint32_t* __restrict__ out_pack,
                                    uint32_t* __restrict__ pack_bad, double* __restrict__ weight) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_classes) return;
  const uint32_t r = class_read[c];
  const uint2 loc = rd[r];  // one sector says where the class's list is and how long
  const uint32_t b = loc.x, n = loc.y, d = class_off[c];
  bool bad = false;
  for (uint32_t j = 0; j < n; ++j) {
    const uint2 pr = cand[b + j];
    const uint32_t t = pr.x;
    const int32_t sc = (int32_t)pr.y;
    out_tid[d + j] = t;
    out_score[d + j] = sc;
    // packed copy for the EM iterations: transcript in 24 bits, score in 8 (half the bytes per pair: the 20 x 2
    // passes over the pairs then run out of L2); a score above 255 or an id above 2^24 keeps the unpacked arrays
    out_pack[d + j] = (t & 0xFFFFFFu) | ((uint32_t)sc << 24);
    bad |= (uint32_t)sc > 255u || t > 0xFFFFFFu;
  }
  if (bad) *pack_bad = 1;
  weight[c] = (double)(class_pos[c + 1] - class_pos[c]);
}

// heads + class ids (cid = exclusive scan of head, n_reads+1 entries) + per class: (first read, position, count)
void launch_class_heads(const uint64_t* keys, uint64_t n_reads, const uint2* rd,
                        const void* fp, const uint2* cand, bool exact, uint32_t* head, uint32_t* cid,
                        uint32_t* scan_tmp, uint32_t* class_read, uint32_t* class_pos, uint32_t* class_cnt,
                        cudaStream_t s, uint64_t* launches) {
  if (!n_reads) return;
  const uint32_t grid = (uint32_t)((n_reads + 255) / 256);
  if (exact) class_head_exact_kernel<<<grid, 256, 0, s>>>(keys, n_reads, rd, cand, head);
  else class_head_kernel<<<(uint32_t)((n_reads + 1023) / 1024), 256, 0, s>>>(keys, n_reads, static_cast<const ulonglong2*>(fp), head);
  launch_exclusive_scan(head, cid, (uint32_t)n_reads, scan_tmp, s, launches);
  class_fill_kernel<<<grid, 256, 0, s>>>(head, cid, keys, n_reads, rd, class_read, class_pos, class_cnt);
  if (launches) *launches += 2;
}

void launch_class_gather(const uint32_t* class_read, const uint32_t* class_pos, const uint32_t* class_cnt,
                         uint32_t* class_off, uint32_t n_classes, uint32_t* scan_tmp, const uint2* rd, const uint2* cand, uint32_t* out_tid, int32_t* out_score,
                         uint32_t* out_pack, uint32_t* pack_bad, double* weight, cudaStream_t s, uint64_t* launches) {
  launch_exclusive_scan(class_cnt, class_off, n_classes, scan_tmp, s, launches);
  cudaMemsetAsync(pack_bad, 0, 4, s);
  if (!n_classes) return;
  class_gather_kernel<<<(n_classes + 255) / 256, 256, 0, s>>>(class_read, class_pos, class_off, n_classes, rd,
                                                              cand, out_tid, out_score, out_pack, pack_bad,
                                                              weight);
  if (launches) ++*launches;
}

// ------------------------------------------------------------------ transcript-major view
// packed: cand_tid holds transcript | score << 24; the score moves to the top 8 bits of the class word
__global__ void make_sort_keys_kernel(const uint32_t* __restrict__ read_off, uint64_t n_reads,
                                      const uint32_t* __restrict__ cand_tid, uint64_t* __restrict__ keys, bool packed) {
  const uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint32_t b = read_off[r], e = read_off[r + 1];
  for (uint32_t j = b; j < e; ++j) {
    const uint32_t t = cand_tid[j];
    keys[j] = packed ? ((uint64_t)((uint32_t)r | (t & 0xFF000000u)) << 32) | (t & 0xFFFFFFu) : ((uint64_t)r << 32) | t;
  }
}

// keys sorted by transcript (low 32 bits): toff[t] = first pair of transcript t, toff[T] = P
__global__ void seg_offsets_kernel(const uint64_t* __restrict__ keys, uint64_t P, uint32_t T,
                                   uint32_t* __restrict__ toff) {
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j > P) return;
  const int64_t prev = j == 0 ? -1 : (int64_t)(uint32_t)keys[j - 1];
  const int64_t cur = j == P ? (int64_t)T : (int64_t)(uint32_t)keys[j];
  for (int64_t t = prev + 1; t <= cur; ++t) toff[t] = (uint32_t)j;
}

// tm_read[j] = class of the j-th pair in transcript order (packed: with its score in the top 8 bits)
__global__ void split_keys_kernel(const uint64_t* __restrict__ keys, uint64_t P, uint32_t* __restrict__ tm_read) {
  const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (j < P) tm_read[j] = (uint32_t)(keys[j] >> 32);
}

__global__ void seg_count_kernel(const uint32_t* __restrict__ toff, uint32_t T, uint32_t seg, uint32_t* nseg) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) nseg[t] = (toff[t + 1] - toff[t] + seg - 1) / seg;
}

__global__ void seg_expand_kernel(const uint32_t* __restrict__ toff, const uint32_t* __restrict__ seg_off,
                                  uint32_t T, uint32_t seg, uint32_t* __restrict__ seg_tid,
                                  uint32_t* __restrict__ seg_begin) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t b = toff[t], e = toff[t + 1];
  uint32_t s = seg_off[t];
  for (uint32_t p = b; p < e; p += seg, ++s) {
    seg_tid[s] = t;
    seg_begin[s] = p;
  }
}

// ------------------------------------------------------------------ EM
__global__ void em_init_kernel(double* pi, uint32_t T, uint32_t* state) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) pi[t] = 1.0 / (double)T;  // isoform_assignment.cpp:17-20
  if (t == 0) { state[0] = 0; state[1] = 0; state[2] = 0; }  // [0]=converged flag, [1]=iterations executed, [2]=block ticket
}

// streaming read of data that is used once per pass: do not let it displace the gathered vectors (pi, 1/den) in L1
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int32_t ld_stream(const int32_t* p) {
  return (int32_t)ld_stream(reinterpret_cast<const uint32_t*>(p));
}

// per read (class): den = sum_j pi[t_j]*s_j in candidate order; E-step: inv = 1/den when den > 1e-10 (:36-45),
// else 0; assignment: tot = den (:80-85).  A block owns 256 consecutive rows: their pairs are one contiguous
// range, so the products are formed with coalesced loads into shared memory and each thread then adds up its
// own row in order (same sums as a plain per-thread loop, without the strided global reads).
static constexpr uint32_t kDenCap = 3072;

template <bool ASSIGN, bool PACKED>  // PACKED: cand_tid holds transcript | score << 24, cand_score is not read
__global__ void __launch_bounds__(256) em_den_kernel(const uint32_t* __restrict__ read_off, uint64_t n_reads,
                                                     const uint32_t* __restrict__ cand_tid,
                                                     const int32_t* __restrict__ cand_score,
                                                     const double* __restrict__ pi, const double* __restrict__ weight,
                                                     double* __restrict__ out, const uint32_t* __restrict__ state) {
  __shared__ double s_term[kDenCap];
  if (!ASSIGN && state[0]) return;
  const uint64_t c0 = (uint64_t)blockIdx.x * 256;
  const uint32_t nc = (uint32_t)min((uint64_t)256, n_reads - c0);
  const uint32_t b0 = read_off[c0], np = read_off[c0 + nc] - b0;
  const bool staged = np <= kDenCap;
  if (staged) {
    // four independent (id, score, pi) load chains per thread: the kernel waits on DRAM latency, not bandwidth
    for (uint32_t j = threadIdx.x; j < np; j += 1024) {
      uint32_t t[4];
      int32_t sc[4];
      double p[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t q = j + 256 * u;
        t[u] = q < np ? ld_stream(cand_tid + b0 + q) : 0u;
        if (PACKED) {
          sc[u] = (int32_t)(t[u] >> 24);
          t[u] &= 0xFFFFFFu;
        } else {
          sc[u] = q < np ? ld_stream(cand_score + b0 + q) : 0;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = j + 256 * u < np ? pi[t[u]] : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (j + 256 * u < np) s_term[j + 256 * u] = p[u] * (double)sc[u];
    }
    __syncthreads();
  }
  if (threadIdx.x >= nc) return;
  const uint64_t r = c0 + threadIdx.x;
  const uint32_t b = read_off[r], e = read_off[r + 1];
  double den = 0.0;
  if (staged) {
    for (uint32_t j = b; j < e; ++j) den += s_term[j - b0];
  } else {
    for (uint32_t j = b; j < e; ++j) {
      const uint32_t t = cand_tid[j];
      den += PACKED ? pi[t & 0xFFFFFFu] * (double)(t >> 24) : pi[t] * (double)cand_score[j];
    }
  }
  // a class of w identical reads adds w identical posteriors: fold w into the reciprocal (exact for w = 1)
  out[r] = ASSIGN ? den : (den > 1e-10 ? (1.0 / den) * weight[r] : 0.0);
}

// one group of G lanes per segment of <= seg pairs of one transcript: partial posterior sum (:46-49).  A
// transcript has a few dozen pairs on average, so 8-lane groups keep the lanes busy; G = 32 for deep data.
template <int G, bool PACKED>  // PACKED: tm_read holds class | score << 24, tm_score is not read
__global__ void em_partial_kernel(const uint32_t* __restrict__ seg_tid, const uint32_t* __restrict__ seg_begin,
                                  const uint32_t* __restrict__ toff, const uint32_t* __restrict__ n_seg_ptr, uint32_t seg,
                                  const uint32_t* __restrict__ tm_read, const uint32_t* __restrict__ tm_score,
                                  const double* __restrict__ inv_den, const double* __restrict__ pi,
                                  double* __restrict__ partial, const uint32_t* __restrict__ state) {
  if (state[0]) return;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / G, gl = threadIdx.x & (G - 1);
  const bool valid = w < *n_seg_ptr;  // the grid covers an upper bound (no host round trip for the exact count)
  double acc = 0.0;
  if (valid) {
    const uint32_t t = seg_tid[w];
    const uint32_t b = seg_begin[w], e = min(b + seg, toff[t + 1]);
    const double p = pi[t];
    uint32_t j = b + gl;
    for (; j + 3 * G < e; j += 4 * G) {  // four gathers in flight, added in the same order as the plain loop
      uint32_t sc[4], c[4];
      double iv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = ld_stream(tm_read + j + u * G);
        if (PACKED) { sc[u] = c[u] >> 24; c[u] &= 0xFFFFFFu; } else { sc[u] = ld_stream(tm_score + j + u * G); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) iv[u] = inv_den[c[u]];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += (p * (double)(int32_t)sc[u]) * iv[u];
    }
    for (; j < e; j += G) {
      uint32_t c = ld_stream(tm_read + j), sc;
      if (PACKED) { sc = c >> 24; c &= 0xFFFFFFu; } else { sc = ld_stream(tm_score + j); }
      acc += (p * (double)(int32_t)sc) * inv_den[c];
    }
  }
#pragma unroll
  for (int d = G / 2; d; d >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, d, G);
  if (valid && gl == 0) partial[w] = acc;
}

__global__ void seg_sum_kernel(const uint32_t* __restrict__ seg_off, uint32_t T,
                               const double* __restrict__ partial, double* __restrict__ out,
                               const uint32_t* __restrict__ state) {
  if (state && state[0]) return;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double s = 0.0;
  for (uint32_t i = seg_off[t]; i < seg_off[t + 1]; ++i) s += partial[i];
  out[t] = s;
}

// M-step (:54-60): new = ps + (double)(0.01f/(float)R) + (double)0.01f, evaluated left to right
__global__ void __launch_bounds__(256) em_update_kernel(const double* __restrict__ ps, double* __restrict__ pi,
                                                        uint32_t T, double add_a, double add_b,
                                                        double* __restrict__ block_change,
                                                        const uint32_t* __restrict__ state) {
  if (state[0]) return;
  __shared__ double sh[256];
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  double ch = 0.0;
  if (t < T) {
    const double np = (ps[t] + add_a) + add_b;
    ch = fabs(np - pi[t]);
    pi[t] = np;
  }
  sh[threadIdx.x] = ch;
  __syncthreads();
  for (int d = 128; d; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_change[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) em_converge_kernel(const double* __restrict__ block_change, uint32_t nb,
                                                          double tol, uint32_t* state, double* last_change) {
  if (state[0]) return;
  __shared__ double sh[256];
  double s = 0.0;
  for (uint32_t i = threadIdx.x; i < nb; i += 256) s += block_change[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int d = 128; d; d >>= 1) {
    if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    state[1] += 1;
    *last_change = sh[0];
    if (sh[0] < tol) state[0] = 1;  // :62-64
  }
}

// fixed-order sum of one double per thread over a 256-thread block (warp shuffles, then the 8 warp sums)
__device__ __forceinline__ double block_sum_256(double v, double* sh8) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31) == 0) sh8[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) s += sh8[w];
  return s;  // valid in thread 0
}

// single GPU: segment sums, M-step and the convergence test in one launch.  The last block to finish (ticket
// in state[2]) adds the per-block changes in a fixed order.
__global__ void __launch_bounds__(256) em_mstep_fused_kernel(const uint32_t* __restrict__ seg_off,
                                                             const double* __restrict__ partial,
                                                             double* __restrict__ ps, double* __restrict__ pi,
                                                             uint32_t T, double add_a, double add_b,
                                                             double* __restrict__ block_change, double tol,
                                                             uint32_t* state, double* last_change) {
  if (state[0]) return;
  __shared__ double sh[8], sh2[8];
  __shared__ uint32_t s_last;
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  double ch = 0.0;
  if (t < T) {
    double sum = 0.0;
    for (uint32_t i = seg_off[t]; i < seg_off[t + 1]; ++i) sum += partial[i];
    ps[t] = sum;
    const double np = (sum + add_a) + add_b;
    ch = fabs(np - pi[t]);
    pi[t] = np;
  }
  const double bsum = block_sum_256(ch, sh);
  if (threadIdx.x == 0) {
    block_change[blockIdx.x] = bsum;
    __threadfence();
    s_last = atomicAdd(&state[2], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += 256) s += *(volatile double*)&block_change[i];
  const double tot = block_sum_256(s, sh2);
  if (threadIdx.x == 0) {
    state[2] = 0;
    *last_change = tot;
    __threadfence();
    state[1] += 1;
    if (tot < tol) state[0] = 1;  // :62-64
  }
}

// ------------------------------------------------------------------ several GPUs: exchange fused into the M-step
// Reads are sharded, so every rank holds partial posterior sums of all T transcripts.  Instead of seg_sum ->
// ncclAllReduce -> update -> converge (a collective and three launches per iteration), the ranks exchange through
// buffers they have mapped from each other (CUDA IPC over NVLink / NVSwitch), fused into the two kernels that have
// to run anyway:
//   1. seg_sum_signal: a rank's per-transcript sums go to its exchange slot; the last block to finish raises the
//      rank's flag in every peer's memory.
//   2. em_mstep_peer: waits for everybody's flag, loads the N vectors over the links (eight loads in flight per
//      thread), adds them in rank order and does the M-step and the convergence test exactly like the one-GPU
//      kernel: the same bits, hence the same decision, on every rank, and no collective.
// Every rank reads all of every peer's vector, so the bytes grow with N while NCCL's in-switch reduction does not:
// measured on 2 MB vectors, this beats ncclAllReduce at N = 2 (EM 1.74 against 1.93 ms per 20 iterations) and N = 4
// (1.97 against 2.06 ms) and loses at N = 8 (2.55 against 2.33 ms; remote loads reached ~250 GB/s).  The engine
// therefore uses it up to four ranks and NCCL beyond.  Also measured and dropped (profiles/r02_notes.md): pushing the sums into every peer
// with remote stores (1.83 ms at N = 2), and a reduce-scatter + all-gather in two flag rounds (1.97 ms at N = 2:
// the second round costs more than the bytes it saves).  Slots and flags are double-buffered by iteration
// parity: nobody can be more than one iteration ahead of the slowest rank, because every M-step waits for
// everybody's flag; epochs only grow, so a stale flag never matches.
__device__ __forceinline__ void peer_flag_store(unsigned long long* f, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
}
// spin until the flag reaches `epoch`; a peer that died must not hang the GPU: after ~4 s the error word is set
__device__ __forceinline__ void peer_flag_wait(const unsigned long long* f, unsigned long long epoch, uint32_t* err) {
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v >= epoch) return;
    if (clock64() - t0 > 8000000000ll) { atomicOr(err, 1u); return; }
  }
}
__device__ __forceinline__ double ld_peer(const double* p) {  // written by another GPU during this kernel's life
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// a rank's exchange memory: [2 slots][T] doubles; flags: [2 slots][nranks] u64
struct PeerView {
  double* const* x;               // every rank's exchange memory (mine included)
  unsigned long long* const* f;   // every rank's flags
  uint32_t nranks, rank, slot, T;
  unsigned long long epoch;
  uint32_t* err;                  // [0] time-out flag, [1] ticket of seg_sum_signal
};

__global__ void __launch_bounds__(256) seg_sum_signal_kernel(const uint32_t* __restrict__ seg_off,
                                                           const double* __restrict__ partial,
                                                           const uint32_t* __restrict__ state, const PeerView P) {
  if (state[0]) return;  // converged (on every rank at the same iteration): nobody waits for this flag
  __shared__ uint32_t s_last;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < P.T) {
    double s = 0.0;
    for (uint32_t i = seg_off[t]; i < seg_off[t + 1]; ++i) s += partial[i];
    P.x[P.rank][(size_t)P.slot * P.T + t] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(P.err + 1, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();  // every block's sums are visible system-wide before the flag is
  if (threadIdx.x < P.nranks) peer_flag_store(P.f[threadIdx.x] + (size_t)P.slot * P.nranks + P.rank, P.epoch);
  if (threadIdx.x == 0) P.err[1] = 0;
}

__global__ void __launch_bounds__(256) em_mstep_peer_kernel(const PeerView P, double* ps, double* pi, double add_a,
                                                            double add_b, double* __restrict__ block_change, double tol,
                                                            uint32_t* state, double* last_change) {
  if (state[0]) return;
  __shared__ double sh[8], sh2[8];
  __shared__ uint32_t s_last;
  if (threadIdx.x < P.nranks)
    peer_flag_wait(P.f[P.rank] + (size_t)P.slot * P.nranks + threadIdx.x, P.epoch, P.err);
  __syncthreads();
  const uint32_t t = blockIdx.x * 256 + threadIdx.x;
  double ch = 0.0;
  if (t < P.T) {
    double sum = 0.0;
    for (uint32_t r0 = 0; r0 < P.nranks; r0 += 8) {  // eight remote loads in flight, added in rank order (everywhere)
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = r0 + u < P.nranks ? ld_peer(P.x[r0 + u] + (size_t)P.slot * P.T + t) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < P.nranks) sum += v[u];
    }
    ps[t] = sum;
    const double np = (sum + add_a) + add_b;
    ch = fabs(np - pi[t]);
    pi[t] = np;
  }
  const double bsum = block_sum_256(ch, sh);
  if (threadIdx.x == 0) {
    block_change[blockIdx.x] = bsum;
    __threadfence();
    s_last = atomicAdd(&state[2], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double s = 0.0;
  for (uint32_t i = threadIdx.x; i < gridDim.x; i += 256) s += *(volatile double*)&block_change[i];
  const double tot = block_sum_256(s, sh2);
  if (threadIdx.x == 0) {
    state[2] = 0;
    *last_change = tot;
    __threadfence();
    state[1] += 1;
    if (tot < tol) state[0] = 1;  // :62-64
  }
}

// ------------------------------------------------------------------ assignment (:70-97)
template <int G, bool PACKED>
__global__ void as_partial_kernel(const uint32_t* __restrict__ seg_tid, const uint32_t* __restrict__ seg_begin,
                                  const uint32_t* __restrict__ toff, const uint32_t* __restrict__ n_seg_ptr, uint32_t seg,
                                  const uint32_t* __restrict__ tm_read, const uint32_t* __restrict__ tm_score,
                                  const double* __restrict__ tot, const double* __restrict__ weight,
                                  const double* __restrict__ pi, double* __restrict__ partial,
                                  uint32_t* __restrict__ present_u32) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) / G, gl = threadIdx.x & (G - 1);
  const bool valid = w < *n_seg_ptr;
  uint32_t t = 0;
  double acc = 0.0;
  bool any = false;
  if (valid) {
    t = seg_tid[w];
    const uint32_t b = seg_begin[w], e = min(b + seg, toff[t + 1]);
    const double p = pi[t];
    uint32_t j = b + gl;
    for (; j + 3 * G < e; j += 4 * G) {  // four gather chains in flight, terms added in the plain loop's order
      uint32_t c[4], sc[4];
      double tt[4], wt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = ld_stream(tm_read + j + u * G);
        if (PACKED) { sc[u] = c[u] >> 24; c[u] &= 0xFFFFFFu; } else { sc[u] = ld_stream(tm_score + j + u * G); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { tt[u] = tot[c[u]]; wt[u] = weight[c[u]]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (tt[u] > 0.0) {
          acc += ((p * (double)(int32_t)sc[u]) / tt[u]) * wt[u];
          any = true;
        }
    }
    for (; j < e; j += G) {
      uint32_t c = tm_read[j], sc;
      if (PACKED) { sc = c >> 24; c &= 0xFFFFFFu; } else { sc = tm_score[j]; }
      const double tt = tot[c];
      if (tt > 0.0) {
        acc += ((p * (double)(int32_t)sc) / tt) * weight[c];  // :90 divides per term; w identical reads
        any = true;
      }
    }
  }
#pragma unroll
  for (int d = G / 2; d; d >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, d, G);
  const uint32_t anyw = __ballot_sync(0xFFFFFFFFu, any);
  const uint32_t gmask = (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << (lane_id() & ~(uint32_t)(G - 1));
  if (valid && gl == 0) {
    partial[w] = acc;
    if (anyw & gmask) present_u32[t] = 1;  // benign race: all writers store 1
  }
}

// ------------------------------------------------------------------ host-side launch helpers
void launch_make_sort_keys(const uint32_t* read_off, uint64_t n_reads, const uint32_t* cand_tid, bool packed,
                           uint64_t* keys, cudaStream_t s, uint64_t* launches) {
  if (!n_reads) return;
  make_sort_keys_kernel<<<(uint32_t)((n_reads + 255) / 256), 256, 0, s>>>(read_off, n_reads, cand_tid, keys, packed);
  if (launches) ++*launches;
}

void launch_tmajor(const uint64_t* keys, uint64_t P, uint32_t T, uint32_t seg, uint32_t* toff, uint32_t* tm_read,
                   uint32_t* nseg, uint32_t* seg_off, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches) {
  seg_offsets_kernel<<<(uint32_t)((P + 1 + 255) / 256), 256, 0, s>>>(keys, P, T, toff);
  if (P) split_keys_kernel<<<(uint32_t)((P + 255) / 256), 256, 0, s>>>(keys, P, tm_read);
  seg_count_kernel<<<(T + 255) / 256, 256, 0, s>>>(toff, T, seg, nseg);
  if (launches) *launches += 3;
  launch_exclusive_scan(nseg, seg_off, T, scan_tmp, s, launches);
}

void launch_seg_expand(const uint32_t* toff, const uint32_t* seg_off, uint32_t T, uint32_t seg, uint32_t* seg_tid,
                       uint32_t* seg_begin, cudaStream_t s, uint64_t* launches) {
  seg_expand_kernel<<<(T + 255) / 256, 256, 0, s>>>(toff, seg_off, T, seg, seg_tid, seg_begin);
  if (launches) ++*launches;
}

void launch_em_init(double* pi, uint32_t T, uint32_t* state, cudaStream_t s, uint64_t* launches) {
  em_init_kernel<<<(T + 255) / 256, 256, 0, s>>>(pi, T, state);
  if (launches) ++*launches;
}

// lanes per segment: a transcript-major run averages n_pairs / n_seg pairs
static inline bool narrow_groups(const EmView& v) { return v.n_pairs / (v.T ? v.T : 1) < 96; }  // ~pairs per transcript

void launch_em_estep(const EmView& v, cudaStream_t s, uint64_t* launches, bool with_sum) {
  if (v.n_reads) {
    if (v.packed)
      em_den_kernel<false, true><<<(uint32_t)((v.n_reads + 255) / 256), 256, 0, s>>>(
          v.read_off, v.n_reads, v.cand_tid, v.cand_score, v.pi, v.weight, v.read_tmp, v.state);
    else
      em_den_kernel<false, false><<<(uint32_t)((v.n_reads + 255) / 256), 256, 0, s>>>(
          v.read_off, v.n_reads, v.cand_tid, v.cand_score, v.pi, v.weight, v.read_tmp, v.state);
    if (launches) ++*launches;
  }
  if (v.n_seg) {
#define SQ_EM_PARTIAL(G, PK)                                                                                   \
  em_partial_kernel<G, PK><<<(uint32_t)(((uint64_t)v.n_seg * G + 255) / 256), 256, 0, s>>>(                    \
      v.seg_tid, v.seg_begin, v.toff, v.seg_off + v.T, v.seg, v.tm_read, v.tm_score, v.read_tmp, v.pi, v.partial, v.state)
    if (narrow_groups(v)) {
      if (v.packed) SQ_EM_PARTIAL(8, true); else SQ_EM_PARTIAL(8, false);
    } else {
      if (v.packed) SQ_EM_PARTIAL(32, true); else SQ_EM_PARTIAL(32, false);
    }
#undef SQ_EM_PARTIAL
    if (launches) ++*launches;
  }
  if (with_sum) {
    seg_sum_kernel<<<(v.T + 255) / 256, 256, 0, s>>>(v.seg_off, v.T, v.partial, v.ps, v.state);
    if (launches) ++*launches;
  }
}

void launch_em_mstep(const EmView& v, double add_a, double add_b, double tol, cudaStream_t s, uint64_t* launches) {
  const uint32_t nb = (v.T + 255) / 256;
  em_update_kernel<<<nb, 256, 0, s>>>(v.ps, v.pi, v.T, add_a, add_b, v.block_change, v.state);
  em_converge_kernel<<<1, 256, 0, s>>>(v.block_change, nb, tol, v.state, v.last_change);
  if (launches) *launches += 2;
}

// segment sums + M-step + convergence test, no collective in between (one GPU)
void launch_em_mstep_fused(const EmView& v, double add_a, double add_b, double tol, cudaStream_t s, uint64_t* launches) {
  em_mstep_fused_kernel<<<(v.T + 255) / 256, 256, 0, s>>>(v.seg_off, v.partial, v.ps, v.pi, v.T, add_a, add_b,
                                                         v.block_change, tol, v.state, v.last_change);
  if (launches) ++*launches;
}

// the exchange + M-step of one iteration (see above)
void launch_em_mstep_peer(const EmView& v, double* const* peer_x, unsigned long long* const* peer_flags, uint32_t rank,
                          uint32_t nranks, unsigned long long epoch, double add_a, double add_b, double tol, uint32_t* err,
                          cudaStream_t s, uint64_t* launches) {
  PeerView P;
  P.x = peer_x;
  P.f = peer_flags;
  P.nranks = nranks;
  P.rank = rank;
  P.slot = (uint32_t)(epoch & 1);
  P.T = v.T;
  P.epoch = epoch;
  P.err = err;
  seg_sum_signal_kernel<<<(v.T + 255) / 256, 256, 0, s>>>(v.seg_off, v.partial, v.state, P);
  em_mstep_peer_kernel<<<(v.T + 255) / 256, 256, 0, s>>>(P, v.ps, v.pi, add_a, add_b, v.block_change, tol, v.state,
                                                        v.last_change);
  if (launches) *launches += 2;
}

void launch_assign(const EmView& v, double* numreads, uint32_t* present_u32, cudaStream_t s, uint64_t* launches) {
  cudaMemsetAsync(present_u32, 0, sizeof(uint32_t) * v.T, s);
  if (v.n_reads) {
    if (v.packed)
      em_den_kernel<true, true><<<(uint32_t)((v.n_reads + 255) / 256), 256, 0, s>>>(
          v.read_off, v.n_reads, v.cand_tid, v.cand_score, v.pi, nullptr, v.read_tmp, nullptr);
    else
      em_den_kernel<true, false><<<(uint32_t)((v.n_reads + 255) / 256), 256, 0, s>>>(
          v.read_off, v.n_reads, v.cand_tid, v.cand_score, v.pi, nullptr, v.read_tmp, nullptr);
    if (launches) ++*launches;
  }
  if (v.n_seg) {
#define SQ_AS_PARTIAL(G, PK)                                                                                   \
  as_partial_kernel<G, PK><<<(uint32_t)(((uint64_t)v.n_seg * G + 255) / 256), 256, 0, s>>>(                    \
      v.seg_tid, v.seg_begin, v.toff, v.seg_off + v.T, v.seg, v.tm_read, v.tm_score, v.read_tmp, v.weight, v.pi,   \
      v.partial, present_u32)
    if (narrow_groups(v)) {
      if (v.packed) SQ_AS_PARTIAL(8, true); else SQ_AS_PARTIAL(8, false);
    } else {
      if (v.packed) SQ_AS_PARTIAL(32, true); else SQ_AS_PARTIAL(32, false);
    }
#undef SQ_AS_PARTIAL
    if (launches) ++*launches;
  }
  seg_sum_kernel<<<(v.T + 255) / 256, 256, 0, s>>>(v.seg_off, v.T, v.partial, numreads, nullptr);
  if (launches) ++*launches;
}


}  // namespace sq

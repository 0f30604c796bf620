// Launcher prototypes of sq_em.cu / sq_vote.cu helpers used by the engine.
#pragma once
#include "sq_common.cuh"

namespace sq {

struct EmView {
  // read-major CSR
  const uint32_t* read_off;
  const uint32_t* cand_tid;   // packed: transcript | score << 24 (cand_score unused)
  const int32_t* cand_score;
  bool packed;                // both copies of the pairs carry the score in the top 8 bits of their 32-bit word
  uint64_t n_reads;           // rows of the CSR = equivalence classes of reads
  const double* weight;       // reads per class
  // transcript-major copy, split into segments of <= seg pairs
  const uint32_t* toff;
  const uint32_t* tm_read;    // packed: class | score << 24 (tm_score unused)
  const uint32_t* tm_score;
  const uint32_t* seg_off;    // per transcript: first segment (T+1 entries)
  const uint32_t* seg_tid;
  const uint32_t* seg_begin;
  uint32_t n_seg;             // upper bound (grids); the exact count is seg_off[T] on the device
  uint32_t seg;
  uint64_t n_pairs;           // rows of the transcript-major copy
  uint32_t T;
  // state
  double* pi;
  double* ps;
  double* read_tmp;      // per read: 1/den (E-step) or tot (assignment)
  double* partial;       // per segment
  double* block_change;  // per 256-transcript block
  double* last_change;
  uint32_t* state;       // [0] converged, [1] iterations executed, [2] block ticket of the fused M-step
};

void launch_fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t s);
size_t vote_smem_bytes(uint32_t nk);

void launch_bmap_build(const uint32_t* keys, uint64_t nkeys, uint4* bmap, uint32_t n_sectors, uint32_t* cnt,
                       uint32_t* excl, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches);
void launch_permute_out(const double* pi, const double* numreads, const uint32_t* present, const uint32_t* ext_of,
                        uint32_t T, double* pi_out, double* nr_out, uint8_t* present_out, cudaStream_t s,
                        uint64_t* launches);

// candidate store: per-read class keys + fingerprints behind a batch's vote; the store in read order (taps)
void launch_read_keys(const uint2* rd, uint64_t r0, uint64_t n,
                      const uint2* cand, uint32_t T, uint32_t hash_bits, uint64_t* rkey, void* rfp, cudaStream_t s,
                      uint64_t* launches);
void launch_csr_gather(const uint2* rd, uint32_t* cnt_tmp, uint32_t* off, uint64_t n_reads,
                       uint32_t* scan_tmp, const uint2* cand, uint32_t* out_tid, int32_t* out_score, cudaStream_t s,
                       uint64_t* launches);
void launch_class_heads(const uint64_t* keys, uint64_t n_reads, const uint2* rd,
                        const void* fp, const uint2* cand, bool exact, uint32_t* head, uint32_t* cid,
                        uint32_t* scan_tmp, uint32_t* class_read, uint32_t* class_pos, uint32_t* class_cnt,
                        cudaStream_t s, uint64_t* launches);
void launch_class_gather(const uint32_t* class_read, const uint32_t* class_pos, const uint32_t* class_cnt,
                         uint32_t* class_off, uint32_t n_classes, uint32_t* scan_tmp, const uint2* rd, const uint2* cand, uint32_t* out_tid, int32_t* out_score,
                         uint32_t* out_pack, uint32_t* pack_bad, double* weight, cudaStream_t s, uint64_t* launches);
void launch_make_sort_keys(const uint32_t* read_off, uint64_t n_reads, const uint32_t* cand_tid, bool packed,
                           uint64_t* keys, cudaStream_t s, uint64_t* launches);
void launch_tmajor(const uint64_t* keys, uint64_t P, uint32_t T, uint32_t seg, uint32_t* toff, uint32_t* tm_read,
                   uint32_t* nseg, uint32_t* seg_off, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches);
void launch_seg_expand(const uint32_t* toff, const uint32_t* seg_off, uint32_t T, uint32_t seg, uint32_t* seg_tid,
                       uint32_t* seg_begin, cudaStream_t s, uint64_t* launches);
void launch_em_init(double* pi, uint32_t T, uint32_t* state, cudaStream_t s, uint64_t* launches);
void launch_em_estep(const EmView& v, cudaStream_t s, uint64_t* launches, bool with_sum);
void launch_em_mstep_fused(const EmView& v, double add_a, double add_b, double tol, cudaStream_t s, uint64_t* launches);
void launch_em_mstep(const EmView& v, double add_a, double add_b, double tol, cudaStream_t s, uint64_t* launches);
void launch_em_mstep_peer(const EmView& v, double* const* peer_x, unsigned long long* const* peer_flags, uint32_t rank,
                          uint32_t nranks, unsigned long long epoch, double add_a, double add_b, double tol, uint32_t* err,
                          cudaStream_t s, uint64_t* launches);
void launch_assign(const EmView& v, double* numreads, uint32_t* present_u32, cudaStream_t s, uint64_t* launches);

}  // namespace sq

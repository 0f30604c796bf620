// Launchers of sq_tap.cu.
#pragma once
#include "sq_common.cuh"

namespace sq {
void launch_tap_count(const uint32_t* item_start, uint32_t n_reads, uint32_t nk, const uint16_t* cnt,
                      uint32_t n_items_ub, uint32_t* counts, cudaStream_t s, uint64_t* launches);
void launch_tap_gather(const uint32_t* item_start, uint32_t n_reads, uint32_t nk, const uint16_t* cnt,
                       uint32_t n_items_ub, const uint32_t* hsel, uint64_t hstride, const uint32_t* hoff,
                       const uint32_t* offs, uint64_t cap, uint32_t* out, uint64_t* keys64, const uint32_t* seq_tid,
                       uint32_t tbits, cudaStream_t s, uint64_t* launches);
void launch_post_flags(const uint64_t* keys, uint64_t n, uint32_t tbits, uint32_t* newpair, uint32_t* newkey,
                       cudaStream_t s, uint64_t* launches);
void launch_post_scatter(const uint64_t* keys, uint64_t n, uint32_t tbits, const uint32_t* newpair,
                         const uint32_t* newkey, const uint32_t* ppos, const uint32_t* kpos, uint32_t* out_keys,
                         unsigned long long* out_off, uint32_t* out_tid, cudaStream_t s, uint64_t* launches);
}  // namespace sq

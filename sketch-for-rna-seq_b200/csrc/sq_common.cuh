// Shared declarations of the sm_100a quant path: launcher prototypes and small device helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SQ_MAXK 8             // k values per index (SQ_MAX_K_COUNT)
#define SQ_CHUNK 256          // window-end positions handled by one sketch thread
#define SQ_EMPTY 0xFFFFFFFFu  // empty marker in hash tables / bucket offsets
#define SQ_LAST 0x80000000u   // flag on the last transcript id of a posting list
#define SQ_LIST_HDR 8          // header words of a posting list: length, then one or two (base, 64-bit mask) id ranges
#define SQ_NOMASK 0xFFFFFFFFu  // first base of a list whose transcripts need more than two 64-id ranges

namespace sq {

// ---------- ntHash2 forward hash, 33-bit lane only ----------
// The reference keeps (uint32_t)get_forward_hash() (src/sketch.cpp:33).  ntHash2's split rotation never
// mixes bits 63..33 with bits 32..0, so those 32 bits depend only on the 33-bit low lane.  A lane value s
// is held as two words: x = s[31:0], y = s[32:1]; rotating the lane left by one is then
//   x' = funnelshift_l(y, x, 1), y' = x.
__host__ __device__ inline uint64_t seed33(uint32_t code) {
  // low 33 bits of ntHash2's SEED_A/C/G/T (0x3c8bfbb395c60474, 0x3193c18562a02b4c, 0x20323ed082572324,
  // 0x295549f54be24456)
  return code == 0 ? 0x195c60474ULL : code == 1 ? 0x162a02b4cULL : code == 2 ? 0x082572324ULL : 0x14be24456ULL;
}
__host__ __device__ inline uint64_t rol33(uint64_t s, uint32_t n) {
  n %= 33;
  if (n == 0) return s;
  return ((s << n) | (s >> (33 - n))) & 0x1FFFFFFFFULL;
}

// per-k lookup table in the two-word form (384 bytes, 128-byte aligned in shared memory):
//   e[0..15]  = seed[in] ^ rol33^k(seed[out])       index in*4+out   (one rolling step)
//   e[16..31] = rol33(seed[a]) ^ seed[b]            index a + 4*b    (two bases at once while the first
//                                                                      window is being filled; a comes first)
//   e[32..35] = seed[in]                                              (one base while filling)
struct KLut {
  uint2 e[48];
};

struct SketchParams {
  const uint32_t* packed;      // 2-bit bases, 16 per word
  uint64_t n_words;            // words readable at `packed` (padded to a multiple of 4)
  const uint32_t* base_off;    // per read: first base; (base_off - bias) is relative to packed[0]
  uint32_t bias;
  const uint32_t* len;         // per read
  const uint32_t* item_read;   // per item: read index
  const uint32_t* item_start;  // per read: first item (exclusive scan of items per read), n_reads+1 entries
  uint32_t n_reads;
  uint32_t n_items_ub;         // launch bound; the true count is item_start[n_reads]
  uint32_t nk;
  uint32_t threshold;
  uint32_t ks[SQ_MAXK];
  uint32_t kmax;
  // output: the selected hashes of the batch, DENSE per k-index: an item's hashes are hsel[k*hstride + hoff .. + cnt);
  // the regions of a warp's 32 items are contiguous and in item order, warps reserve theirs with one atomic
  uint32_t* hsel;
  uint64_t hstride;            // words per k-index (capacity: every k-mer of the batch)
  uint32_t* hoff;              // [nk][n_items_ub]
  uint16_t* cnt;               // [nk][n_items_ub] hashes of the item; SQ_CNT_RAW set: not de-duplicated
  uint32_t* cursor;            // [nk] words used so far (zeroed by the caller)
  uint32_t cap;                // staging entries per lane in shared memory (a lane that selects more re-rolls)
  uint32_t stage_words;        // packed words a warp can stage in shared memory (a multiple of 4; a warp whose span is
                               // longer reads from global memory)
  uint32_t dedup;              // 1: an item's equal hashes are stored once (the sketch is a set, include/sketch.h:15)
  unsigned long long* stats;   // optional: += selected hashes of the launch before de-duplication (one atomic per block)
  KLut lut[SQ_MAXK];
};
#define SQ_CNT_RAW 0x8000u     // flag in cnt: the item overflowed its staging area, duplicates were not removed
#define SQ_CNT_MASK 0x7FFFu

// Index of one k in HBM (replicated per GPU), sized to stay in L2 at human scale (random 32-byte gathers run at
// 230 G/s from <= 64 MB and fall to 45-60 G/s from DRAM, profiles/micro/gather_bench.cu):
//   bmap     the key set as a bitmap over [0, max key] with interleaved ranks: sectors of 32 B = {number of keys
//            in all earlier sectors, 224 bits}.  One sector answers "is h a key" exactly (no chains, no
//            divergence, whatever the load) and gives the key's rank.  Keys are FracMinHash values <= threshold:
//            31 MB at scale 0.05.
//   desc     32-bit list descriptor of every key, in key order (the rank indexes it): 6.2 M keys -> 25 MB
//            bit 31 = 0  the list itself: transcript `base` (low tbits bits) and, above it, a mask of the following
//                        31 - tbits ids (bit i <=> base+1+i is in the list) -- the isoforms of a gene have neighbouring
//                        ids, nine lists in ten fit and need no further access
//            bit 31 = 1  id of the list in lhdr
//            (SQ_EMPTY is what the lookup reports for a hash that is not a key)
//   lhdr     one 16-byte header per DISTINCT posting list (lists with equal content are stored once, so equal
//            descriptors <=> equal lists): {base1 | two<<31, mask1_lo, mask1_hi, posting offset}: bit i of mask1 <=>
//            transcript base1+i is in the list; SQ_NOMASK in the first word when two 64-id ranges do not cover
//            the list.  1.2 M lists -> 20 MB, touched by one hit in ten
//   postings per list (32-byte aligned): 8 header words {length, base1 | two<<31, mask1_lo, mask1_hi, base2,
//            mask2_lo, mask2_hi, 0}, then the transcript ids ascending, the last one flagged with SQ_LAST
// Transcript ids in here are the engine's INTERNAL ids (see sq_load_index: transcripts that share posting
// lists are renumbered next to each other, so that a list is a window base + bit mask whatever order the
// caller's ids came in).
#define SQ_BMAP_BITS 224u   // key values per bitmap sector
struct IndexTable {
  const uint4* bmap;        // 2*n_sectors uint4
  const uint32_t* desc;     // one per key, ascending key order
  const uint4* lhdr;        // n_lists
  const uint32_t* postings;
  uint32_t n_sectors;       // covers keys < n_sectors * SQ_BMAP_BITS
  uint32_t present;         // 0: k-index has no map (sparse_chaining.cpp:51-53)
  uint32_t tbits;           // bits of a transcript id inside an inline descriptor
};

#ifdef __CUDACC__
// ------------------------------------------------------------------ list fingerprint (EM read classes, see sq_em.cu)
// Two independent 64-bit hashes over (length, transcripts, scores) of a candidate list, folded step by step.
struct ListHash {
  uint64_t h, g;
  __device__ __forceinline__ void init(uint32_t n) {
    h = 0xcbf29ce484222325ull ^ n;
    g = 0x9E3779B97F4A7C15ull + n;
  }
  __device__ __forceinline__ void add(uint32_t tid, int32_t score) {
    const uint64_t x = ((uint64_t)tid << 32) | (uint32_t)score;
    h ^= x;
    h *= 0x100000001b3ull;
    h ^= h >> 31;
    g = (g ^ (x * 0xC2B2AE3D27D4EB4Full)) * 0xD6E8FEB86659FD93ull;
    g ^= g >> 29;
  }
  // sort key in the high word (best candidate, then a few hash bits), read index in the low word
  __device__ __forceinline__ uint64_t key(uint64_t top, uint32_t hash_bits, uint64_t r) const {
    return (((top << hash_bits) | ((h ^ (h >> 32)) & ((1ull << hash_bits) - 1))) << 32) | r;
  }
};
#endif

struct VoteParams {
  const uint32_t* base_off;
  uint32_t bias;
  const uint32_t* len;
  const uint32_t* item_start;
  uint32_t n_reads;
  uint32_t n_items_ub;
  uint32_t nk;
  double fraction;
  const uint32_t* hsel;              // the sketch kernel's output (see SketchParams)
  uint32_t* pay;                     // list descriptor of every selected hash (lookup kernel), same layout as hsel
  uint64_t hstride;
  const uint32_t* hoff;
  const uint16_t* cnt;
  const uint32_t* hcursor;           // [nk] selected hashes of the batch per k-index
  uint32_t count_bits;               // planes of the bit-sliced counters (5: up to 31 hashes per read and k, 7: 127)
  uint32_t force_tier;               // tests: 0 = automatic, 1 = every read to the window kernel, 2 = every read to the general kernel
  IndexTable tab[SQ_MAXK];
  // output: the free tail of the engine's candidate store (no staging: what the vote writes is final)
  uint2* stage;                      // store + stage_base: a pair = {transcript, score} in one 8-byte word
  uint64_t stage_cap;                // free pairs behind stage_base
  uint32_t stage_base;               // pairs of the earlier batches
  unsigned long long* stage_cursor;  // device counter: pairs of this batch
  uint2* read_loc;                   // per read of the batch: {start of its list in the store (absolute), candidates}
  // per read of the batch: class sort key and 128-bit list fingerprint, folded by whichever kernel emits the list
  uint64_t* rkey;
  void* rfp;                         // ulonglong2
  uint64_t read_base;                // index of the batch's first read among all pushed reads (low word of the key)
  uint32_t key_T, key_hash_bits;
  uint32_t* mid_list;                // reads the bit-sliced kernel hands to the warp-per-read window kernel
  uint32_t* mid_count;
  uint32_t* slow_list;               // reads the window kernel hands to the general warp-per-read kernel
  uint32_t* slow_count;
  uint32_t* ovf_list;                // reads that overflowed the shared-memory tables
  uint32_t* ovf_count;
  uint32_t* flags;                   // bit1: large-table overflow
  unsigned long long* work;          // [0] queries, [1] hits (lookup kernel), [2] postings voted (vote kernels)
  // large-table scratch (one region per worker block of the overflow kernel)
  uint32_t* big_keys; uint32_t* big_cnt; uint32_t* big_list; uint32_t* big_set; unsigned long long* big_cand;
  uint32_t big_cap_log2;     // table slots per worker
  uint32_t big_set_log2;     // dedup-set slots per worker
  uint32_t n_workers;
};

// ---------- launchers (definitions in the .cu files) ----------
struct KList {
  uint32_t nk;
  uint32_t k[SQ_MAXK];
};
// stats (optional): [0] += k-mers of the batch, [1] += bases of the batch
void launch_items(const uint32_t* len, uint32_t n_reads, uint32_t* nit, uint32_t* item_start, uint32_t* item_read,
                  uint32_t n_items_ub, uint32_t* scan_tmp, cudaStream_t s, uint64_t* launches, const KList& ks,
                  unsigned long long* stats);
void launch_sketch(const SketchParams& p, cudaStream_t s, uint64_t* launches);
cudaError_t sketch_configure();  // per device
uint32_t sketch_stage_words_max();
// list descriptor (see IndexTable) of every selected hash of the batch, one k-index at a time
struct VoteDeviceCfg;
void launch_lookup(const VoteParams& p, const VoteDeviceCfg& cfg, uint32_t ki, cudaStream_t s, uint64_t* launches);
void launch_fixed_layout(uint32_t* len, uint32_t* base_off, uint32_t n_reads, uint32_t read_len, uint32_t* item_start,
                         uint32_t* item_read, const KList& ks, unsigned long long* stats, cudaStream_t s,
                         uint64_t* launches);
void launch_derive_offsets(const uint32_t* len, uint32_t n_reads, uint32_t* tmp, uint32_t* base_off, uint32_t* scan_tmp,
                           cudaStream_t s, uint64_t* launches);
// ev_a/ev_b (optional): recorded right before/after the main short-read kernel only.  tail/fork (optional): the
// small follow-up kernels (reads the first kernel handed on) run on `tail` after `fork`, so that they overlap
// whatever the caller enqueues next on `s`.  Returns the stream the last kernel went to.
struct VoteDeviceCfg {  // per engine, i.e. per device: filled by vote_configure() with the engine's device current
  int sm_count = 0;
  int vote_grid = 0;  // persistent grid of the warp-per-read kernel
  int long_grid = 0;  // persistent grid of the warp-per-read window kernel
  int lookup_grid = 0;
  uint32_t nk = 0;
};
cudaError_t vote_configure(uint32_t nk, VoteDeviceCfg* cfg);
cudaStream_t launch_vote(const VoteParams& p, const VoteDeviceCfg& cfg, cudaStream_t s, uint64_t* launches,
                         cudaEvent_t ev_a = nullptr, cudaEvent_t ev_b = nullptr, cudaStream_t tail = nullptr,
                         cudaEvent_t fork = nullptr);

// exclusive scan of n u32 values; out[n] receives the total (out has n+1 entries); tmp holds >= n/2048+2 u32
void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp, cudaStream_t s,
                           uint64_t* launches);
size_t scan_tmp_words(uint32_t n);

// stable LSD radix sort of (key64, val32) pairs (vals may be NULL) on bits [bit_lo, bit_lo+nbits); buffers are
// ping-ponged, the result pointers are returned through *keys_out / *vals_out (one of the two buffers each)
void launch_radix_sort(uint64_t* keys_a, uint64_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, uint64_t n,
                       int nbits, uint32_t* hist_tmp, uint64_t** keys_out, uint32_t** vals_out, cudaStream_t s,
                       uint64_t* launches, int bit_lo = 0);
size_t radix_tmp_words(uint64_t n);

}  // namespace sq

// warp helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if ((int)lane_id() >= d) v += t;
  }
  return v;
}

// C ABI of the quant hot path (include/sketchquant.h): engine state, batch pipeline, EM driver, NCCL glue.
// There is deliberately no CPU implementation behind these entry points: without a CUDA device every call
// fails with SQ_ERR_NO_DEVICE.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sketchquant.h"
#include "sq_common.cuh"
#include "sq_kernels.cuh"
#include "sq_tap.cuh"

using namespace sq;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as() const { return static_cast<T*>(p); }
  // grow-only; contents are NOT preserved
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Slot {
  DevBuf packed, base_off, len;  // only used for host pushes
  DevBuf nit, item_start, item_read, cnt, hsel, pay, hoff, ovf_list, slow_list, mid_list, scan_tmp;
  cudaEvent_t copied = nullptr, voted = nullptr, fork = nullptr;
  bool pending = false;     // vote enqueued, its exact candidate count not yet seen by the host
  uint32_t n_reads = 0;
  uint64_t read_base = 0;
  VoteParams vp;
  int id = 0;
  void release() {
    DevBuf* all[] = {&packed, &base_off, &len, &nit, &item_start, &item_read, &cnt, &hsel, &pay, &hoff,
                     &ovf_list, &slow_list, &mid_list, &scan_tmp};
    for (DevBuf* b : all) b->release();
  }
};

struct KTab {
  DevBuf bmap, desc, lhdr, postings;
  uint32_t n_sectors = 0;
  bool present = false;
  uint64_t nkeys = 0, npost = 0, npost_stored = 0, nlists = 0;
};

struct StageEvent { cudaEvent_t a, b; int stage; };

// NCCL entry points resolved at run time (single-GPU use needs no NCCL at all)
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string* err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { *err = std::string("cannot load libnccl: ") + dlerror(); return false; }
    GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
    CommInitRank = reinterpret_cast<decltype(CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
    AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { *err = "libnccl lacks required symbols"; return false; }
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

struct sq_engine {
  int device = 0;
  uint32_t nk = 0;
  uint32_t ks[SQ_MAXK] = {0};
  uint32_t kmax = 0, kmin = 0;
  uint32_t threshold = 0;
  double fraction = 0.9;
  uint64_t T = 0;
  KLut lut[SQ_MAXK];
  VoteDeviceCfg vote_cfg;  // kernel attributes and grids of THIS engine's device
  cudaStream_t stream = nullptr, own_stream = nullptr, copy_stream = nullptr, tail_stream = nullptr;
  bool profiling = false;
  std::string err;
  KTab tab[SQ_MAXK];
  // internal transcript numbering (decided by the first sq_load_index): ext_of[internal] = caller's id
  std::vector<uint32_t> ext_of, int_of;
  bool perm_ready = false, perm_identity = true;
  DevBuf d_ext_of;
  // options
  uint64_t batch_bases = 1ull << 28;
  uint32_t cand_per_read = 16;
  uint32_t n_workers = 64;
  uint32_t max_read_len = 1u << 20;
  uint32_t em_seg = 1024;
  uint32_t sub_batch_reads = 1u << 20;  // sq_push_reads_fixed: reads per internal batch
  bool exact_classes = false;  // compare candidate lists element-wise instead of by 128-bit fingerprint
  uint32_t vote_tier = 0;      // tests: force every read through one vote tier (see VoteParams::force_tier)
  // batch slots
  Slot slot[2];
  int next_slot = 0;
  // device counters + pinned mirror
  unsigned long long* d_totals = nullptr;        // stats: [0] sketch hashes, [1..3] probes/hits/postings, [4] k-mers, [5] bases
  unsigned long long* d_slot_ctr = nullptr;      // per slot: [2*i] staging cursor, [2*i+1] = overflow reads (low u32) | slow-path reads (high u32)
  uint32_t* d_flags = nullptr;
  uint32_t* d_fail = nullptr;
  uint32_t* d_hcur = nullptr;                    // selected-hash cursors: [8*slot + k], slot 2 = tap
  unsigned long long* h_mirror = nullptr;        // pinned copy of d_slot_ctr after each vote
  uint64_t P = 0;                                 // candidate pairs of all finalized batches (exact)
  uint64_t ovf_total = 0, slow_total = 0, mid_total = 0;
  // large-table scratch
  DevBuf big_keys, big_cnt, big_list, big_set, big_cand;
  uint32_t big_cap_log2 = 0, big_set_log2 = 0;
  bool big_ready = false;
  // candidate store over all pushed reads: the vote kernels append a read's list at an atomic cursor (read r:
  // cand[rd[r].x .. + rd[r].y)), so the lists of a batch are contiguous but not in read order
  uint2* cand = nullptr;  // a pair = {transcript, score}: one 8-byte word, one sector where two arrays would touch two
  uint2* rd = nullptr;  // per read: {start of its list in cand, candidates}
  uint64_t* rkey = nullptr;              // per read: class sort key and 128-bit list fingerprint (see sq_em.cu), written
  void* rfp = nullptr;                   // behind every batch's vote
  bool keys_valid = true;                // false: the store was filled by sq_set_candidates
  uint32_t class_hash_bits = 14;
  uint64_t cand_cap = 0, read_cap = 0;
  uint64_t n_reads = 0, n_bases = 0, n_batches = 0;  // n_reads: all enqueued batches
  // EM scratch
  DevBuf keys_a, keys_b, vals_a, vals_b, sort_tmp, toff, tm_read, nseg, seg_off, seg_tid, seg_begin, pi, ps,
      read_tmp, partial, block_change, misc, numreads, present, scan_tmp, em_off, em_cnt, em_tid, em_score, em_pack,
      cls_head, cls_id, cls_read, cls_pos, cls_weight, out_pi, out_nr, out_present;
  uint64_t n_classes_last = 0, n_cpairs_last = 0;
  int em_iterations = 0;
  // sq_sketch / sq_build_postings scratch
  Slot tap;
  DevBuf tap_counts, tap_offs, tap_out, tap_tid, bp_newpair, bp_newkey, bp_ppos, bp_kpos, bp_keys, bp_off, bp_post;
  uint64_t bp_nkeys = 0, bp_npost = 0;
  int bp_kidx = -1;
  const void* bp_sig = nullptr;
  // profiling
  std::vector<StageEvent> events;
  float ms[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // sketch, vote, (unused), sort, em, assign, items, vote main kernel, lookup
  uint32_t n_stage[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t launches = 0;
  // NCCL
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // EM exchange over peer memory (see sq_em.cu): this rank's exchange memory ([2 slots][T] sums) and
  // flags ([2][nranks]), the peers' mappings of theirs (CUDA IPC), the same as device arrays of pointers; epoch
  // counts iterations over the engine's life (slot = epoch & 1, flags only ever grow)
  bool peer_ok = false, peer_used = false;
  int peer_wanted = 1;  // option peer_exchange: 0 never, 1 where it was measured faster (up to four ranks), 2 whenever possible
  double* xbuf = nullptr;
  unsigned long long* xflags = nullptr;
  std::vector<void*> peer_mapped;  // what cudaIpcOpenMemHandle returned (closed in sq_destroy)
  DevBuf d_peer_ps, d_peer_flags, d_peer_err;
  unsigned long long epoch = 0;
};

namespace {

int fail(sq_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define SQ_CUDA(e, call)                                                                          \
  do {                                                                                            \
    cudaError_t _err = (call);                                                                    \
    if (_err != cudaSuccess)                                                                      \
      return fail((e), SQ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_err), __FILE__, __LINE__); \
  } while (0)

#define SQ_TRY(call)              \
  do {                            \
    int _rc = (call);             \
    if (_rc != SQ_OK) return _rc; \
  } while (0)

struct StageScope {
  sq_engine* e;
  int stage;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  StageScope(sq_engine* e_, int s, cudaStream_t on = nullptr) : e(e_), stage(s), st(on ? on : e_->stream) {
    if (e->profiling) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~StageScope() {
    if (e->profiling) {
      cudaEventRecord(b, st);
      e->events.push_back({a, b, stage});
    }
  }
};

void resolve_events(sq_engine* e) {
  for (auto& ev : e->events) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess) { e->ms[ev.stage] += ms; e->n_stage[ev.stage]++; }
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  e->events.clear();
}

uint32_t log2_ceil(uint64_t v) {
  uint32_t l = 0;
  while ((1ull << l) < v) ++l;
  return l;
}

void make_lut(uint32_t k, KLut* out) {
  memset(out, 0, sizeof(*out));
  auto two = [](uint64_t v) { return make_uint2((uint32_t)v, (uint32_t)(v >> 1)); };
  for (uint32_t in = 0; in < 4; ++in) {
    for (uint32_t o = 0; o < 4; ++o) out->e[in * 4 + o] = two(seed33(in) ^ rol33(seed33(o), k));
    for (uint32_t b = 0; b < 4; ++b) out->e[16 + in + 4 * b] = two(rol33(seed33(in), 1) ^ seed33(b));
    out->e[32 + in] = two(seed33(in));
  }
}

// expected upper end of the selected hashes of one item (SQ_CHUNK window ends at most): twice the mean + 8
uint32_t item_hash_bound(const sq_engine* e, uint64_t n_bases, uint32_t n_reads) {
  const double scale = ((double)e->threshold + 1.0) / 4294967296.0;
  const uint64_t mean_len = n_reads ? n_bases / n_reads : 0;
  const double windows = mean_len <= 200 ? (double)mean_len + 32.0 : (double)SQ_CHUNK;
  const double b = windows * scale * 2.0 + 8.0;
  return (uint32_t)std::min<double>(b, (double)SQ_CHUNK);
}
// Shared memory of the sketch kernel, sized for the batch: `cap` entries per lane for the selected hashes of an item
// (mean + 4 sigma of a binomial count, plus the 16 entries of head room the kernel's unrolled block wants; a lane
// that selects more rolls a second time, straight to global memory, and its read is voted by the window kernel) and `stage_words` packed words per warp for the reads
// themselves (a warp whose 32 items span more reads from global memory).  Both are upper ends, not limits.
void sketch_smem_plan(const sq_engine* e, uint64_t n_bases, uint32_t n_reads, uint32_t* cap, uint32_t* stage_words) {
  const double scale = ((double)e->threshold + 1.0) / 4294967296.0;
  const uint64_t mean_len = n_reads ? (n_bases + n_reads - 1) / n_reads : 0;
  const double item_bases = (double)std::min<uint64_t>(mean_len + 8, SQ_CHUNK);
  const double windows = mean_len + 8 <= SQ_CHUNK ? std::max(item_bases - (double)e->kmin + 1.0, 1.0) : (double)SQ_CHUNK;
  const double m = windows * scale;
  // the unrolled 16-step block of the kernel runs only while 16 more entries fit: that much head room on top
  const uint32_t bound = (uint32_t)std::min<double>(m + 4.0 * std::sqrt(m) + 1.0, (double)SQ_CHUNK);
  *cap = std::min<uint32_t>(SQ_CHUNK, ((bound + 1) & ~1u) + 16);
  // a lane touches its item's bases plus kmax - 1 before them; 32 lanes, one word of slack each, 16-byte rounding
  const uint64_t lane_bases = std::min<uint64_t>(mean_len + 3, SQ_CHUNK) + (mean_len > SQ_CHUNK ? e->kmax : 0);
  const uint64_t w = 32 * ((lane_bases + 15) / 16 + 1) + 8;
  *stage_words = (uint32_t)std::min<uint64_t>(sketch_stage_words_max(), (w + 3) & ~3ull);
}

int check_flags(sq_engine* e) {
  uint32_t flags = 0;
  SQ_CUDA(e, cudaMemcpyAsync(&flags, e->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, e->stream));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  if (flags & 2u) return fail(e, SQ_ERR_CAPACITY, "large-table overflow: a read is longer than option max_read_len (%u)", e->max_read_len);
  if (flags & 4u) return fail(e, SQ_ERR_CAPACITY, "candidate store overflow (internal growth bound violated)");
  return SQ_OK;
}

// make sure the store can take `reads` more reads (beyond read_base) and `pairs` more pairs (beyond P); contents
// are preserved.  Growth is geometric; a steady-state pass (sq_reset_reads between passes) never reallocates.
int ensure_store(sq_engine* e, uint64_t read_base, uint64_t reads, uint64_t pairs) {
  // called with every earlier batch finalized, i.e. with no kernel in flight that touches the store
  cudaStream_t cs = e->tail_stream;
  const uint64_t need_reads = read_base + reads + 1;
  if (need_reads > e->read_cap) {
    const uint64_t cap = std::max<uint64_t>(need_reads + need_reads / 2, 1 << 16);
    uint2* ps = nullptr;
    uint64_t* pk = nullptr;
    void* pf = nullptr;
    SQ_CUDA(e, cudaMalloc(&ps, cap * sizeof(uint2)));
    SQ_CUDA(e, cudaMalloc(&pk, cap * sizeof(uint64_t)));
    SQ_CUDA(e, cudaMalloc(&pf, cap * 16));
    if (e->rd && read_base) {
      SQ_CUDA(e, cudaMemcpyAsync(ps, e->rd, read_base * sizeof(uint2), cudaMemcpyDeviceToDevice, cs));
      SQ_CUDA(e, cudaMemcpyAsync(pk, e->rkey, read_base * sizeof(uint64_t), cudaMemcpyDeviceToDevice, cs));
      SQ_CUDA(e, cudaMemcpyAsync(pf, e->rfp, read_base * 16, cudaMemcpyDeviceToDevice, cs));
    }
    SQ_CUDA(e, cudaStreamSynchronize(cs));
    if (e->rd) SQ_CUDA(e, cudaFree(e->rd));
    if (e->rkey) SQ_CUDA(e, cudaFree(e->rkey));
    if (e->rfp) SQ_CUDA(e, cudaFree(e->rfp));
    e->rd = ps;
    e->rkey = pk;
    e->rfp = pf;
    e->read_cap = cap;
  }
  const uint64_t need_pairs = e->P + pairs;
  if (need_pairs >= 0xFFFFFFF0ull) return fail(e, SQ_ERR_CAPACITY, "more than 2^32 candidate pairs on one engine");
  if (need_pairs > e->cand_cap) {
    uint64_t cap = std::max<uint64_t>(need_pairs + need_pairs / 2, 1 << 20);
    if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
    uint2* t = nullptr;
    SQ_CUDA(e, cudaMalloc(&t, cap * sizeof(uint2)));
    if (e->cand && e->P) SQ_CUDA(e, cudaMemcpyAsync(t, e->cand, e->P * sizeof(uint2), cudaMemcpyDeviceToDevice, cs));
    SQ_CUDA(e, cudaStreamSynchronize(cs));
    if (e->cand) SQ_CUDA(e, cudaFree(e->cand));
    e->cand = t;
    e->cand_cap = cap;
  }
  return SQ_OK;
}

// point a batch's vote at the free tail of the store (called with the exact P of all earlier batches)
void aim_vote_at_store(sq_engine* e, Slot& s) {
  s.vp.stage = e->cand + e->P;
  s.vp.stage_cap = e->cand_cap - e->P;
  s.vp.stage_base = (uint32_t)e->P;
  s.vp.read_loc = e->rd + s.read_base;
  s.vp.rkey = e->rkey;
  s.vp.rfp = e->rfp;
  s.vp.read_base = s.read_base;
  s.vp.key_T = (uint32_t)e->T;
  s.vp.key_hash_bits = e->class_hash_bits;
}

int ensure_big_scratch(sq_engine* e) {
  if (e->big_ready) return SQ_OK;
  e->big_cap_log2 = std::max<uint32_t>(10, log2_ceil(e->T + e->T / 3 + 2));
  e->big_set_log2 = std::max<uint32_t>(12, log2_ceil((uint64_t)e->max_read_len + e->max_read_len / 3 + 2));
  const size_t tab = (size_t)1 << e->big_cap_log2, set = (size_t)1 << e->big_set_log2, w = e->n_workers;
  SQ_CUDA(e, e->big_keys.ensure(w * tab * 4));
  SQ_CUDA(e, e->big_cnt.ensure(w * tab * e->nk * 4));
  SQ_CUDA(e, e->big_list.ensure(w * tab * 4));
  SQ_CUDA(e, e->big_set.ensure(w * set * 4));
  SQ_CUDA(e, e->big_cand.ensure(w * tab * 8));
  launch_fill_u32(e->big_keys.as<uint32_t>(), w * tab, SQ_EMPTY, e->stream);
  launch_fill_u32(e->big_set.as<uint32_t>(), w * set, SQ_EMPTY, e->stream);
  SQ_CUDA(e, cudaMemsetAsync(e->big_cnt.p, 0, w * tab * e->nk * 4, e->stream));
  e->launches += 2;
  e->big_ready = true;
  return SQ_OK;
}

int enqueue_vote(sq_engine* e, Slot& s) {
  unsigned long long* ctr = e->d_slot_ctr + 4 * s.id;
  SQ_CUDA(e, cudaMemsetAsync(ctr, 0, 32, e->stream));
  // The vote's follow-up kernels (the few reads its first kernel hands on) go to the tail stream: nothing on
  // the engine stream needs them until the host has seen `voted`, and the next batch's sketch overlaps them.
  cudaStream_t last = e->stream;
  {
    StageScope st(e, 1);
    cudaEvent_t a = nullptr, b = nullptr;
    if (e->profiling) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
    }
    last = launch_vote(s.vp, e->vote_cfg, e->stream, &e->launches, a, b, e->tail_stream, s.fork);
    if (a) e->events.push_back({a, b, 7});
  }
  SQ_CUDA(e, cudaMemcpyAsync(e->h_mirror + 4 * s.id, ctr, 32, cudaMemcpyDeviceToHost, last));
  SQ_CUDA(e, cudaGetLastError());
  SQ_CUDA(e, cudaEventRecord(s.voted, last));
  return SQ_OK;
}

// The vote of a batch writes into the free tail of the store and reports the exact number of candidate pairs.
// If the tail was too short the store grows and the vote is simply re-run (the batch's descriptors are still in
// the slot).
int finalize_slot(sq_engine* e, Slot& s) {
  if (!s.pending) return SQ_OK;
  SQ_CUDA(e, cudaEventSynchronize(s.voted));
  uint64_t needed = e->h_mirror[4 * s.id];
  // Re-run the vote chain when the store's tail was too short, or when a read needs the large-table scratch for
  // the first time.  The work counters were complete after the first pass (every read was voted, only not
  // stored), so the re-runs do not count.
  for (;;) {
    const bool need_big = (e->h_mirror[4 * s.id + 1] & 0xFFFFFFFFull) != 0 && s.vp.n_workers == 0;
    if (needed <= s.vp.stage_cap && !need_big) break;
    if (needed > s.vp.stage_cap) {
      SQ_TRY(ensure_store(e, s.read_base, s.n_reads, needed + needed / 8));
      aim_vote_at_store(e, s);
    }
    if (need_big) {
      SQ_TRY(ensure_big_scratch(e));
      s.vp.big_keys = e->big_keys.as<uint32_t>();
      s.vp.big_cnt = e->big_cnt.as<uint32_t>();
      s.vp.big_list = e->big_list.as<uint32_t>();
      s.vp.big_set = e->big_set.as<uint32_t>();
      s.vp.big_cand = e->big_cand.as<unsigned long long>();
      s.vp.big_cap_log2 = e->big_cap_log2;
      s.vp.big_set_log2 = e->big_set_log2;
      s.vp.n_workers = e->n_workers;
    }
    s.vp.work = nullptr;
    SQ_TRY(enqueue_vote(e, s));
    SQ_CUDA(e, cudaEventSynchronize(s.voted));
    needed = e->h_mirror[4 * s.id];
  }
  const uint64_t ovf = e->h_mirror[4 * s.id + 1] & 0xFFFFFFFFull;
  // nothing is left to do for the batch: its lists, class keys and fingerprints were written by the vote kernels
  s.pending = false;
  e->P += needed;
  e->ovf_total += ovf;
  e->slow_total += e->h_mirror[4 * s.id + 1] >> 32;
  e->mid_total += e->h_mirror[4 * s.id + 2] & 0xFFFFFFFFull;
  return SQ_OK;
}

// next slot, free of its previous batch
int acquire_slot(sq_engine* e, Slot** out) {
  Slot& s = e->slot[e->next_slot];
  s.id = e->next_slot;
  e->next_slot ^= 1;
  if (!s.copied) {
    SQ_CUDA(e, cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
    SQ_CUDA(e, cudaEventCreateWithFlags(&s.voted, cudaEventDisableTiming));
    SQ_CUDA(e, cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
  }
  SQ_TRY(finalize_slot(e, s));
  *out = &s;
  return SQ_OK;
}

// enqueue sketch + lookup + vote for one batch whose inputs are in device memory
int run_batch(sq_engine* e, Slot& s, const uint32_t* d_packed, uint64_t n_words, const uint32_t* d_boff, uint32_t bias,
              const uint32_t* d_len, uint32_t n_reads, uint64_t n_bases, cudaEvent_t inputs_ready,
              uint32_t* derive_boff = nullptr, uint32_t fixed_len = 0) {
  if (n_reads == 0) return SQ_OK;
  const uint64_t items_ub64 = (uint64_t)n_reads + n_bases / SQ_CHUNK + 1;
  if (items_ub64 >= 0xFFFFFFFFull || n_bases >= 0xFFFFFFFFull) return fail(e, SQ_ERR_ARG, "batch too large");
  const uint32_t items_ub = (uint32_t)items_ub64;
  const uint64_t hstride = ((n_bases + 3) & ~3ull) + 4;  // capacity per k: every k-mer of the batch (+ lookup padding)
  SQ_CUDA(e, s.nit.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.item_start.ensure(((size_t)n_reads + 1) * 4));
  SQ_CUDA(e, s.item_read.ensure((size_t)items_ub * 4));
  SQ_CUDA(e, s.cnt.ensure((size_t)items_ub * e->nk * 2));
  SQ_CUDA(e, s.hsel.ensure((size_t)hstride * e->nk * 4));
  SQ_CUDA(e, s.pay.ensure((size_t)hstride * e->nk * 4));
  SQ_CUDA(e, s.hoff.ensure((size_t)items_ub * e->nk * 4));
  SQ_CUDA(e, s.ovf_list.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.slow_list.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.mid_list.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.scan_tmp.ensure(scan_tmp_words(std::max(n_reads, items_ub)) * 4));
  // The batch before this one (other slot) must be finalized: the host waits for its vote and learns its exact
  // pair count, which fixes where in the store this batch's vote writes.  Host batches: now (its vote has been
  // overlapping our copies).  Device batches: after our sketch is enqueued (below), so the sketch runs while the
  // previous vote's tail kernels finish.
  if (inputs_ready) {
    SQ_TRY(finalize_slot(e, e->slot[s.id ^ 1]));
    SQ_CUDA(e, cudaStreamWaitEvent(e->stream, inputs_ready, 0));
  }
  KList kl;
  kl.nk = e->nk;
  for (uint32_t i = 0; i < e->nk; ++i) kl.k[i] = e->ks[i];
  const bool one_item_each = derive_boff && fixed_len && fixed_len <= SQ_CHUNK;
  if (derive_boff && fixed_len) {  // equal lengths: offsets and lengths are written on the GPU, nothing was copied
    StageScope st(e, 6);
    launch_fixed_layout(const_cast<uint32_t*>(d_len), derive_boff, n_reads, fixed_len,
                        one_item_each ? s.item_start.as<uint32_t>() : nullptr,
                        one_item_each ? s.item_read.as<uint32_t>() : nullptr, kl, e->d_totals + 4, e->stream, &e->launches);
  } else if (derive_boff) {  // offsets not supplied: reads are packed back to back on 4-base boundaries
    StageScope st(e, 6);
    launch_derive_offsets(d_len, n_reads, s.nit.as<uint32_t>(), derive_boff, s.scan_tmp.as<uint32_t>(), e->stream,
                          &e->launches);
  }

  if (!one_item_each) {
    StageScope st(e, 6);
    launch_items(d_len, n_reads, s.nit.as<uint32_t>(), s.item_start.as<uint32_t>(), s.item_read.as<uint32_t>(),
                 items_ub, s.scan_tmp.as<uint32_t>(), e->stream, &e->launches, kl, e->d_totals + 4);
    // cnt needs no clearing: the sketch kernel writes the count of every item that exists, for every k
  }
  {
    SketchParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.packed = d_packed;
    sp.n_words = n_words;
    sp.base_off = d_boff;
    sp.bias = bias;
    sp.len = d_len;
    sp.item_read = s.item_read.as<uint32_t>();
    sp.item_start = s.item_start.as<uint32_t>();
    sp.n_reads = n_reads;
    sp.n_items_ub = items_ub;
    sp.nk = e->nk;
    sp.threshold = e->threshold;
    for (uint32_t i = 0; i < e->nk; ++i) { sp.ks[i] = e->ks[i]; sp.lut[i] = e->lut[i]; }
    sp.kmax = e->kmax;
    sp.hsel = s.hsel.as<uint32_t>();
    sp.hstride = hstride;
    sp.hoff = s.hoff.as<uint32_t>();
    sp.cnt = s.cnt.as<uint16_t>();
    sp.cursor = e->d_hcur + 8 * s.id;
    sketch_smem_plan(e, n_bases, n_reads, &sp.cap, &sp.stage_words);
    sp.dedup = 1;
    sp.stats = e->d_totals;  // [0] += sketch hashes
    SQ_CUDA(e, cudaMemsetAsync(sp.cursor, 0, 32, e->stream));
    {
      StageScope st(e, 0);  // the sketch kernel alone
      launch_sketch(sp, e->stream, &e->launches);
    }
  }
  if (!inputs_ready) SQ_TRY(finalize_slot(e, e->slot[s.id ^ 1]));
  {
    VoteParams& vp = s.vp;
    memset(&vp, 0, sizeof(vp));
    vp.base_off = d_boff;
    vp.bias = bias;
    vp.len = d_len;
    vp.item_start = s.item_start.as<uint32_t>();
    vp.n_reads = n_reads;
    vp.n_items_ub = items_ub;
    vp.nk = e->nk;
    vp.fraction = e->fraction;
    vp.hsel = s.hsel.as<uint32_t>();
    vp.pay = s.pay.as<uint32_t>();
    vp.hstride = hstride;
    vp.hoff = s.hoff.as<uint32_t>();
    vp.cnt = s.cnt.as<uint16_t>();
    vp.hcursor = e->d_hcur + 8 * s.id;
    vp.count_bits = item_hash_bound(e, n_bases, n_reads) <= 31 ? 5 : 7;
    vp.force_tier = e->vote_tier;
    const uint32_t tbits = std::max<uint32_t>(1, log2_ceil(e->T));
    for (uint32_t i = 0; i < e->nk; ++i) {
      vp.tab[i].bmap = e->tab[i].bmap.as<uint4>();
      vp.tab[i].desc = e->tab[i].desc.as<uint32_t>();
      vp.tab[i].lhdr = e->tab[i].lhdr.as<uint4>();
      vp.tab[i].postings = e->tab[i].postings.as<uint32_t>();
      vp.tab[i].n_sectors = e->tab[i].n_sectors;
      vp.tab[i].present = e->tab[i].present ? 1u : 0u;
      vp.tab[i].tbits = tbits;
    }
    // the vote writes into the free tail of the store (the batch before this one is finalized: P is exact); room
    // for the expected pairs is made here, a batch that needs more is re-run by finalize_slot
    s.n_reads = n_reads;
    s.read_base = e->n_reads;
    uint64_t expect = (uint64_t)e->cand_per_read * n_reads;
    if (e->n_reads && e->P) expect = std::min<uint64_t>(expect, 2 * (e->P / e->n_reads + 1) * n_reads);  // twice the rate so far
    SQ_TRY(ensure_store(e, s.read_base, n_reads, std::max<uint64_t>(expect, 4096)));
    aim_vote_at_store(e, s);
    vp.stage_cursor = e->d_slot_ctr + 4 * s.id;
    vp.ovf_list = s.ovf_list.as<uint32_t>();
    vp.ovf_count = reinterpret_cast<uint32_t*>(e->d_slot_ctr + 4 * s.id + 1);
    vp.slow_list = s.slow_list.as<uint32_t>();
    vp.slow_count = vp.ovf_count + 1;
    vp.mid_list = s.mid_list.as<uint32_t>();
    vp.mid_count = vp.ovf_count + 2;
    vp.flags = e->d_flags;
    vp.big_keys = e->big_keys.as<uint32_t>();
    vp.big_cnt = e->big_cnt.as<uint32_t>();
    vp.big_list = e->big_list.as<uint32_t>();
    vp.big_set = e->big_set.as<uint32_t>();
    vp.big_cand = e->big_cand.as<unsigned long long>();
    vp.big_cap_log2 = e->big_cap_log2;
    vp.big_set_log2 = e->big_set_log2;
    vp.n_workers = e->big_ready ? e->n_workers : 0;  // the large-table scratch is made when a read first needs it
    vp.work = e->d_totals + 1;
    {
      StageScope st(e, 8);  // seed lookup, one k-index at a time (each table L2-resident during its pass)
      for (uint32_t i = 0; i < e->nk; ++i) launch_lookup(vp, e->vote_cfg, i, e->stream, &e->launches);
    }
    SQ_TRY(enqueue_vote(e, s));
    s.pending = true;
  }
  e->n_reads += n_reads;
  e->n_bases += n_bases;
  e->n_batches += 1;
  return SQ_OK;
}

}  // namespace

extern "C" {

const char* sq_version(void) { return "sketchquant-b200 0.1 (sm_100a)"; }

int sq_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return SQ_ERR_NO_DEVICE; }
  return n;
}

const char* sq_last_error(const sq_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

uint32_t sq_threshold_from_fraction(double fraction) {
  const uint32_t H = 0xFFFFFFFFu;
  return static_cast<uint32_t>(H * fraction);
}

int sq_create(sq_engine** out, int device, uint32_t nk, const uint32_t* ks, uint32_t threshold, double chain_fraction,
              uint64_t n_transcripts) {
  if (!out) return fail(nullptr, SQ_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (nk == 0 || nk > SQ_MAXK || !ks) return fail(nullptr, SQ_ERR_ARG, "nk must be in 1..%d", SQ_MAXK);
  for (uint32_t i = 0; i < nk; ++i)
    if (ks[i] == 0 || ks[i] > 4096) return fail(nullptr, SQ_ERR_ARG, "k[%u]=%u out of range 1..4096", i, ks[i]);
  if (n_transcripts == 0 || n_transcripts >= 0x7FFFFFFFull) return fail(nullptr, SQ_ERR_ARG, "n_transcripts out of range");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, SQ_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, SQ_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  sq_engine* e = new sq_engine();
  e->device = device;
  e->nk = nk;
  e->kmin = 0xFFFFFFFFu;
  for (uint32_t i = 0; i < nk; ++i) {
    e->ks[i] = ks[i];
    e->kmax = std::max(e->kmax, ks[i]);
    e->kmin = std::min(e->kmin, ks[i]);
    make_lut(ks[i], &e->lut[i]);
  }
  e->threshold = threshold;
  e->fraction = chain_fraction;
  e->T = n_transcripts;
  {
    const uint32_t top_bits = std::max<uint32_t>(1, log2_ceil(n_transcripts + 1));
    if (top_bits > 24) { delete e; return fail(nullptr, SQ_ERR_CAPACITY, "more than 2^24 transcripts are not supported"); }
    e->class_hash_bits = 32 - top_bits;  // class sort key = (best candidate, hash bits) in 32 bits: 4 radix passes
  }
  auto bail = [&](cudaError_t ce, const char* what) {
    fail(nullptr, SQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ce));
    delete e;
    return SQ_ERR_CUDA;
  };
  cudaError_t ce;
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return bail(ce, "cudaSetDevice");
  if ((ce = vote_configure(nk, &e->vote_cfg)) != cudaSuccess) return bail(ce, "vote_configure");
  if ((ce = sketch_configure()) != cudaSuccess) return bail(ce, "sketch_configure");
  if ((ce = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "cudaStreamCreate");
  if ((ce = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "cudaStreamCreate");
  if ((ce = cudaStreamCreateWithFlags(&e->tail_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(ce, "cudaStreamCreate");
  e->stream = e->own_stream;
  void* ctr = nullptr;
  if ((ce = cudaMalloc(&ctr, 256)) != cudaSuccess) return bail(ce, "cudaMalloc");
  if ((ce = cudaMemset(ctr, 0, 256)) != cudaSuccess) return bail(ce, "cudaMemset");
  e->d_totals = static_cast<unsigned long long*>(ctr);                 // byte 0
  e->d_slot_ctr = e->d_totals + 8;                                     // bytes 64..127 (2 slots x 32 B)
  e->d_flags = reinterpret_cast<uint32_t*>(e->d_totals + 16);          // byte 128
  e->d_fail = e->d_flags + 1;                                          // byte 132
  e->d_hcur = reinterpret_cast<uint32_t*>(e->d_totals + 20);           // bytes 160..255 (3 x 8 counters)
  if ((ce = cudaHostAlloc(reinterpret_cast<void**>(&e->h_mirror), 128, cudaHostAllocMapped)) != cudaSuccess) return bail(ce, "cudaHostAlloc");
  memset(e->h_mirror, 0, 128);
  *out = e;
  return SQ_OK;
}

void sq_destroy(sq_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  resolve_events(e);
  for (void* m : e->peer_mapped) cudaIpcCloseMemHandle(m);
  if (e->xbuf) cudaFree(e->xbuf);
  if (e->xflags) cudaFree(e->xflags);
  e->d_peer_ps.release();
  e->d_peer_flags.release();
  e->d_peer_err.release();
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (auto& s : e->slot) {
    s.release();
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.voted) cudaEventDestroy(s.voted);
    if (s.fork) cudaEventDestroy(s.fork);
  }
  for (auto& t : e->tab) { t.bmap.release(); t.desc.release(); t.lhdr.release(); t.postings.release(); }
  e->d_ext_of.release();
  e->tap.release();
  {
    DevBuf* tb[] = {&e->tap_counts, &e->tap_offs, &e->tap_out, &e->tap_tid, &e->bp_newpair, &e->bp_newkey,
                    &e->bp_ppos, &e->bp_kpos, &e->bp_keys, &e->bp_off, &e->bp_post};
    for (DevBuf* b : tb) b->release();
  }
  DevBuf* all[] = {&e->big_keys, &e->big_cnt, &e->big_list, &e->big_set, &e->big_cand, &e->keys_a, &e->keys_b,
                   &e->vals_a, &e->vals_b, &e->sort_tmp, &e->toff, &e->tm_read, &e->nseg, &e->seg_off, &e->seg_tid,
                   &e->seg_begin, &e->pi, &e->ps, &e->read_tmp, &e->partial, &e->block_change, &e->misc,
                   &e->numreads, &e->present, &e->scan_tmp, &e->em_off, &e->em_cnt, &e->em_tid, &e->em_score, &e->em_pack,
                   &e->cls_head, &e->cls_id, &e->cls_read, &e->cls_pos, &e->cls_weight, &e->out_pi,
                   &e->out_nr, &e->out_present};
  for (DevBuf* b : all) b->release();
  if (e->cand) cudaFree(e->cand);
  if (e->rd) cudaFree(e->rd);
  if (e->rkey) cudaFree(e->rkey);
  if (e->rfp) cudaFree(e->rfp);
  if (e->d_totals) cudaFree(e->d_totals);
  if (e->h_mirror) cudaFreeHost(e->h_mirror);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->tail_stream) cudaStreamDestroy(e->tail_stream);
  delete e;
}

int sq_set_stream(sq_engine* e, void* cuda_stream) {
  if (!e) return SQ_ERR_ARG;
  SQ_CUDA(e, cudaSetDevice(e->device));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return SQ_OK;
}

int sq_set_profiling(sq_engine* e, int enabled) {
  if (!e) return SQ_ERR_ARG;
  e->profiling = enabled != 0;
  return SQ_OK;
}

int sq_set_option(sq_engine* e, const char* name, int64_t value) {
  if (!e || !name) return SQ_ERR_ARG;
  const std::string n(name);
  if (n == "exact_classes") { e->exact_classes = value != 0; return SQ_OK; }
  if (n == "vote_tier") { e->vote_tier = (uint32_t)value; return SQ_OK; }
  if (n == "peer_exchange") { e->peer_wanted = value < 0 ? 0 : (value > 2 ? 2 : (int)value); return SQ_OK; }
  if (e->n_batches) return fail(e, SQ_ERR_STATE, "options must be set before the first push");
  if (value <= 0) return fail(e, SQ_ERR_ARG, "option %s needs a positive value", name);
  if (n == "batch_bases") e->batch_bases = std::min<uint64_t>((uint64_t)value, 0xF0000000ull);
  else if (n == "cand_per_read") e->cand_per_read = (uint32_t)value;
  else if (n == "overflow_workers") { e->n_workers = (uint32_t)value; e->big_ready = false; }
  else if (n == "max_read_len") { e->max_read_len = (uint32_t)value; e->big_ready = false; }
  else if (n == "em_segment") e->em_seg = (uint32_t)value;
  else if (n == "sub_batch_reads") e->sub_batch_reads = (uint32_t)std::min<uint64_t>((uint64_t)value, 1u << 30);
  else return fail(e, SQ_ERR_ARG, "unknown option %s", name);
  return SQ_OK;
}

int sq_load_index(sq_engine* e, uint32_t kidx, uint64_t nkeys, const uint32_t* keys, const uint64_t* post_off,
                  const uint32_t* post_tid) {
  if (!e) return SQ_ERR_ARG;
  if (kidx >= e->nk) return fail(e, SQ_ERR_ARG, "kidx %u out of range", kidx);
  if (nkeys && (!keys || !post_off || !post_tid)) return fail(e, SQ_ERR_ARG, "NULL index array");
  SQ_CUDA(e, cudaSetDevice(e->device));
  const uint64_t npost = nkeys ? post_off[nkeys] : 0;
  if (npost >= 0xFFFFFFF0ull) return fail(e, SQ_ERR_CAPACITY, "more than 2^32 postings for one k");
  if (nkeys >= 0xFFFFFFF0ull) return fail(e, SQ_ERR_CAPACITY, "more than 2^32 keys for one k");
  for (uint64_t i = 0; i < npost; ++i)
    if (post_tid[i] >= e->T) return fail(e, SQ_ERR_ARG, "posting %llu names transcript %u >= T", (unsigned long long)i, post_tid[i]);
  KTab& t = e->tab[kidx];
  const uint32_t T = (uint32_t)e->T;
  const bool trace = getenv("SQ_TRACE") != nullptr;
  auto wall = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_lap = wall();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const double n = wall();
    fprintf(stderr, "[sq trace] load_index k-index %u: %-22s %.3f s\n", kidx, what, n - t_lap);
    t_lap = n;
  };
  // Posting lists with identical content are stored once (all k-mers of an exon shared by the same isoforms
  // have the same list): the table maps a key to the id of its DISTINCT list, so equal ids mean equal lists and
  // a read whose hits share a list merges it once with a weight.  Host threads: (1) sort the lists that are not
  // ascending (the reference's file order is arbitrary) and hash every list's content; (2) each thread owns the
  // lists whose hash falls in its partition: open-addressing table on the 64-bit hash, verified by comparing the
  // content; (3) transcripts are renumbered (first index loaded into this engine only); (4) headers and id lists
  // are laid out in the internal numbering.
  unsigned nth = std::thread::hardware_concurrency();
  if (const char* ev = getenv("SQ_HOST_THREADS")) nth = (unsigned)atoi(ev);
  nth = std::max(1u, std::min(nth, 32u));
  if (nkeys < 200000) nth = 1;
  auto run = [&](auto&& fn) {
    if (nth == 1) { fn(0u); return; }
    std::vector<std::thread> th;
    for (unsigned x = 0; x < nth; ++x) th.emplace_back(fn, x);
    for (auto& x : th) x.join();
  };
  // (1) ascending copies where needed, content hashes
  std::vector<uint64_t> lh(nkeys);
  std::vector<uint32_t> sorted_tid;
  std::vector<uint8_t> unsorted_any(nth, 0);
  run([&](unsigned x) {
    const uint64_t i0 = nkeys * x / nth, i1 = nkeys * (x + 1) / nth;
    for (uint64_t i = i0; i < i1 && !unsorted_any[x]; ++i)
      if (!std::is_sorted(post_tid + post_off[i], post_tid + post_off[i + 1])) unsorted_any[x] = 1;
  });
  const uint32_t* pt = post_tid;
  if (std::any_of(unsorted_any.begin(), unsorted_any.end(), [](uint8_t v) { return v != 0; })) {
    sorted_tid.assign(post_tid, post_tid + npost);
    run([&](unsigned x) {
      const uint64_t i0 = nkeys * x / nth, i1 = nkeys * (x + 1) / nth;
      for (uint64_t i = i0; i < i1; ++i) std::sort(sorted_tid.begin() + post_off[i], sorted_tid.begin() + post_off[i + 1]);
    });
    pt = sorted_tid.data();
  }
  run([&](unsigned x) {
    const uint64_t i0 = nkeys * x / nth, i1 = nkeys * (x + 1) / nth;
    for (uint64_t i = i0; i < i1; ++i) {
      const uint64_t b0 = post_off[i], b1 = post_off[i + 1];
      uint64_t h = 0xcbf29ce484222325ull ^ (b1 - b0);
      for (uint64_t j = b0; j < b1; ++j) { h ^= pt[j]; h *= 0x100000001b3ull; h ^= h >> 29; }
      lh[i] = h ? h : 1;
    }
  });
  lap("sort check + hashes");
  // (2) per-partition de-duplication: rep = first key of each distinct list, support = keys that share it
  std::vector<std::vector<uint32_t>> rep(nth), support(nth);
  std::vector<uint32_t> loc(nkeys, SQ_EMPTY);  // distinct-list number inside the partition
  auto part_of = [&](uint64_t h) { return (unsigned)(((h >> 40) * nth) >> 24); };
  // the keys of every partition, in key order (a counting sort): a thread then walks its own keys only instead of
  // filtering all of them
  std::vector<uint64_t> pstart(nth + 1, 0);
  std::vector<uint32_t> pkeys;
  {
    std::vector<uint8_t> part(nkeys);
    run([&](unsigned x) {
      const uint64_t i0 = nkeys * x / nth, i1 = nkeys * (x + 1) / nth;
      for (uint64_t i = i0; i < i1; ++i) part[i] = post_off[i + 1] > post_off[i] ? (uint8_t)part_of(lh[i]) : (uint8_t)0xFF;
    });
    for (uint64_t i = 0; i < nkeys; ++i)
      if (part[i] != 0xFF) ++pstart[part[i] + 1];
    for (unsigned x = 0; x < nth; ++x) pstart[x + 1] += pstart[x];
    pkeys.resize(pstart[nth]);
    std::vector<uint64_t> cur(pstart.begin(), pstart.end() - 1);
    for (uint64_t i = 0; i < nkeys; ++i)
      if (part[i] != 0xFF) pkeys[cur[part[i]]++] = (uint32_t)i;
  }
  run([&](unsigned x) {
    const uint64_t mine = pstart[x + 1] - pstart[x];
    uint64_t cap = 16;
    while (cap < mine * 2 + 2) cap <<= 1;
    std::vector<uint64_t> hkey(cap, 0);
    std::vector<uint32_t> hval(cap, SQ_EMPTY);
    for (uint64_t pi = pstart[x]; pi < pstart[x + 1]; ++pi) {
      const uint64_t i = pkeys[pi];
      const uint64_t b0 = post_off[i], b1 = post_off[i + 1], h = lh[i];
      uint64_t slot = (h * 0x9E3779B97F4A7C15ull) & (cap - 1);
      for (;;) {
        if (hval[slot] == SQ_EMPTY) {
          hkey[slot] = h;
          hval[slot] = (uint32_t)rep[x].size();
          loc[i] = hval[slot];
          rep[x].push_back((uint32_t)i);
          support[x].push_back(1);
          break;
        }
        if (hkey[slot] == h) {  // same hash: verify content
          const uint64_t q = rep[x][hval[slot]];
          const uint64_t q0 = post_off[q], q1 = post_off[q + 1];
          if (q1 - q0 == b1 - b0 && std::equal(pt + b0, pt + b1, pt + q0)) {
            loc[i] = hval[slot];
            ++support[x][hval[slot]];
            break;
          }
        }
        slot = (slot + 1) & (cap - 1);
      }
    }
  });
  std::vector<uint64_t> lbase(nth + 1, 0);
  for (unsigned x = 0; x < nth; ++x) lbase[x + 1] = lbase[x] + rep[x].size();
  const uint64_t n_lists = lbase[nth];
  std::vector<uint32_t> lid(nkeys, SQ_EMPTY), lrep(n_lists), lsup(n_lists);
  run([&](unsigned x) {
    std::copy(rep[x].begin(), rep[x].end(), lrep.begin() + lbase[x]);
    std::copy(support[x].begin(), support[x].end(), lsup.begin() + lbase[x]);
    const uint64_t i0 = nkeys * x / nth, i1 = nkeys * (x + 1) / nth;
    for (uint64_t i = i0; i < i1; ++i)
      if (loc[i] != SQ_EMPTY) lid[i] = (uint32_t)(lbase[part_of(lh[i])] + loc[i]);
  });
  lap("list de-duplication");
  // (3) Internal transcript numbering.  The vote kernels want the transcripts that share posting lists (the
  // isoforms of a gene) to have neighbouring ids, whatever order the caller's ids came in: a reference-written
  // index stores its transcripts in unordered_map order (src/data_io.cpp:185-196).  Union-find over the lists
  // that at least two keys share (a list with a single key may be a 32-bit hash collision joining two unrelated
  // genes), components ordered by their smallest external id, members ascending.  Decided once per engine.
  if (!e->perm_ready) {
    std::vector<uint32_t> parent(T);
    for (uint32_t i = 0; i < T; ++i) parent[i] = i;
    auto find = [&](uint32_t v) {
      while (parent[v] != v) { parent[v] = parent[parent[v]]; v = parent[v]; }
      return v;
    };
    for (uint64_t l = 0; l < n_lists; ++l) {
      if (lsup[l] < 2) continue;
      const uint64_t b0 = post_off[lrep[l]], b1 = post_off[lrep[l] + 1];
      uint32_t r0 = find(pt[b0]);
      for (uint64_t j = b0 + 1; j < b1; ++j) {
        const uint32_t r1 = find(pt[j]);
        if (r1 != r0) { if (r1 < r0) { parent[r0] = r1; r0 = r1; } else parent[r1] = r0; }  // root = smallest id
      }
    }
    std::vector<uint64_t> order(T);
    for (uint32_t i = 0; i < T; ++i) order[i] = ((uint64_t)find(i) << 32) | i;
    std::sort(order.begin(), order.end());
    e->ext_of.resize(T);
    e->int_of.resize(T);
    e->perm_identity = true;
    for (uint32_t i = 0; i < T; ++i) {
      e->ext_of[i] = (uint32_t)order[i];
      e->int_of[(uint32_t)order[i]] = i;
      if (e->ext_of[i] != i) e->perm_identity = false;
    }
    SQ_CUDA(e, e->d_ext_of.ensure((size_t)T * 4));
    SQ_CUDA(e, cudaMemcpy(e->d_ext_of.p, e->ext_of.data(), (size_t)T * 4, cudaMemcpyHostToDevice));
    e->perm_ready = true;
  }
  lap("renumbering");
  // (4) per distinct list: 8 header words + the internal ids ascending (last one flagged), 32-byte aligned; and
  // the 16-byte entry of the header table
  std::vector<uint32_t> loff(n_lists + 1, 0);
  {
    uint64_t acc = 0;
    for (uint64_t l = 0; l < n_lists; ++l) {
      loff[l] = (uint32_t)acc;
      const uint64_t len = post_off[lrep[l] + 1] - post_off[lrep[l]];
      acc += SQ_LIST_HDR + ((len + 7) & ~7ull);
      if (acc >= 0xFFFFFFF0ull) return fail(e, SQ_ERR_CAPACITY, "posting store exceeds 2^32 words for one k");
    }
    loff[n_lists] = (uint32_t)acc;
  }
  const uint64_t ndp = loff[n_lists];
  std::vector<uint32_t> dpost(ndp, 0);
  std::vector<uint4> hdr(n_lists);
  // descriptor of each list (see IndexTable): the list itself when it is a base and up to 31 - tbits following ids
  std::vector<uint32_t> ldesc(n_lists);
  const uint32_t tbits = std::max<uint32_t>(1, log2_ceil(T));
  const uint32_t wbits = 31 - tbits;
  run([&](unsigned x) {
    const uint64_t l0 = n_lists * x / nth, l1 = n_lists * (x + 1) / nth;
    std::vector<uint32_t> ids;
    for (uint64_t l = l0; l < l1; ++l) {
      const uint64_t b0 = post_off[lrep[l]], b1 = post_off[lrep[l] + 1];
      ids.resize(b1 - b0);
      for (uint64_t j = b0; j < b1; ++j) ids[j - b0] = e->int_of[pt[j]];
      std::sort(ids.begin(), ids.end());
      // a list that names a transcript twice votes twice for it (the reference increments per posting,
      // src/sparse_chaining.cpp:64-69): masks cannot say that, every kernel walks such a list
      const bool dup = std::adjacent_find(ids.begin(), ids.end()) != ids.end();
      const uint32_t tmin = ids[0];
      uint64_t mask1 = 0, mask2 = 0;
      size_t j = 0;
      for (; j < ids.size() && ids[j] - tmin < 64; ++j) mask1 |= 1ull << (ids[j] - tmin);
      const uint32_t base2 = j < ids.size() ? ids[j] : 0;
      for (; j < ids.size() && ids[j] - base2 < 64; ++j) mask2 |= 1ull << (ids[j] - base2);
      uint32_t* d = dpost.data() + loff[l];
      d[0] = (uint32_t)ids.size();
      if (j < ids.size() || dup) {  // three or more ranges (or a repeated id): walked by the general kernels
        d[1] = SQ_NOMASK;
      } else {
        d[1] = tmin | (mask2 ? 0x80000000u : 0u);
        d[2] = (uint32_t)mask1;
        d[3] = (uint32_t)(mask1 >> 32);
        d[4] = base2;
        d[5] = (uint32_t)mask2;
        d[6] = (uint32_t)(mask2 >> 32);
      }
      for (size_t q = 0; q < ids.size(); ++q) d[SQ_LIST_HDR + q] = ids[q] | (q + 1 == ids.size() ? SQ_LAST : 0u);
      hdr[l] = make_uint4(d[1], d[2], d[3], loff[l]);
      const bool fits = d[1] != SQ_NOMASK && !mask2 && (mask1 >> 1) < (1ull << wbits);
      ldesc[l] = fits ? (uint32_t)((mask1 >> 1) << tbits) | tmin : 0x80000000u | (uint32_t)l;
    }
  });
  if (n_lists >= 0x7FFFFFFFull) return fail(e, SQ_ERR_CAPACITY, "more than 2^31 distinct posting lists for one k");
  lap("headers + id lists");
  // keys with a list, in ascending order, each once: the reference's loader does mapping[kmer] = vec
  // (src/data_io.cpp:297), so of a key that a hand-made index repeats the LAST occurrence counts
  std::vector<uint32_t> k2, d2;
  k2.reserve(nkeys);
  d2.reserve(nkeys);
  {
    bool ascending = true;
    for (uint64_t i = 0; i < nkeys; ++i) {
      if (lid[i] == SQ_EMPTY) continue;
      if (!k2.empty() && keys[i] <= k2.back()) ascending = false;
      k2.push_back(keys[i]);
      d2.push_back(ldesc[lid[i]]);
    }
    if (!ascending) {
      std::vector<uint64_t> ord(k2.size());
      for (uint64_t i = 0; i < ord.size(); ++i) ord[i] = ((uint64_t)k2[i] << 32) | i;
      std::sort(ord.begin(), ord.end());
      std::vector<uint32_t> ks, ds;
      ks.reserve(ord.size());
      ds.reserve(ord.size());
      for (uint64_t i = 0; i < ord.size(); ++i) {
        if (i + 1 < ord.size() && (ord[i + 1] >> 32) == (ord[i] >> 32)) continue;  // a later occurrence follows
        ks.push_back((uint32_t)(ord[i] >> 32));
        ds.push_back(d2[(uint32_t)ord[i]]);
      }
      k2.swap(ks);
      d2.swap(ds);
    }
  }
  lap("key order");
  const uint64_t n2 = k2.size();
  const uint32_t n_sectors = n2 ? k2.back() / SQ_BMAP_BITS + 1 : 1;
  SQ_CUDA(e, t.bmap.ensure((size_t)n_sectors * 32));
  SQ_CUDA(e, t.desc.ensure((n2 + 1) * 4));
  SQ_CUDA(e, t.lhdr.ensure((n_lists + 1) * 16));
  SQ_CUDA(e, t.postings.ensure((ndp + 8) * 4));
  t.n_sectors = n_sectors;
  t.nkeys = n2;
  t.npost = npost;
  t.npost_stored = ndp;
  t.nlists = n_lists;
  DevBuf dkeys, dcnt, dexcl, dtmp;
  SQ_CUDA(e, dkeys.ensure((n2 + 1) * 4));
  SQ_CUDA(e, dcnt.ensure(((size_t)n_sectors + 1) * 4));
  SQ_CUDA(e, dexcl.ensure(((size_t)n_sectors + 2) * 4));
  SQ_CUDA(e, dtmp.ensure(scan_tmp_words(n_sectors) * 4));
  if (n2) {
    SQ_CUDA(e, cudaMemcpyAsync(dkeys.p, k2.data(), n2 * 4, cudaMemcpyHostToDevice, e->stream));
    SQ_CUDA(e, cudaMemcpyAsync(t.desc.p, d2.data(), n2 * 4, cudaMemcpyHostToDevice, e->stream));
    if (ndp) SQ_CUDA(e, cudaMemcpyAsync(t.postings.p, dpost.data(), ndp * 4, cudaMemcpyHostToDevice, e->stream));
    if (n_lists) SQ_CUDA(e, cudaMemcpyAsync(t.lhdr.p, hdr.data(), n_lists * 16, cudaMemcpyHostToDevice, e->stream));
  }
  launch_bmap_build(dkeys.as<uint32_t>(), n2, t.bmap.as<uint4>(), n_sectors, dcnt.as<uint32_t>(), dexcl.as<uint32_t>(),
                    dtmp.as<uint32_t>(), e->stream, &e->launches);
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  SQ_CUDA(e, cudaGetLastError());
  dkeys.release();
  dcnt.release();
  dexcl.release();
  dtmp.release();
  lap("upload + bitmap");
  t.present = true;
  return SQ_OK;
}

int sq_push_reads_device(sq_engine* e, const uint32_t* d_packed_words, uint64_t n_words, const uint32_t* d_base_off,
                         const uint32_t* d_len, uint32_t n_reads, uint64_t n_bases_hint) {
  if (!e) return SQ_ERR_ARG;
  if (n_reads == 0) return SQ_OK;
  if (!d_packed_words || !d_base_off || !d_len) return fail(e, SQ_ERR_ARG, "NULL device pointer");
  if ((reinterpret_cast<uintptr_t>(d_packed_words) & 15) != 0) return fail(e, SQ_ERR_ARG, "packed words must be 16-byte aligned");
  SQ_CUDA(e, cudaSetDevice(e->device));
  const uint64_t n_bases = n_bases_hint ? n_bases_hint : n_words * 16;
  Slot* s = nullptr;
  SQ_TRY(acquire_slot(e, &s));
  return run_batch(e, *s, d_packed_words, n_words, d_base_off, 0, d_len, n_reads, n_bases, nullptr);
}

int sq_push_reads(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, const uint32_t* base_off,
                  const uint32_t* len, uint32_t n_reads) {
  if (!e) return SQ_ERR_ARG;
  if (n_reads == 0) return SQ_OK;
  if (!packed_words || !len) return fail(e, SQ_ERR_ARG, "NULL host pointer");
  SQ_CUDA(e, cudaSetDevice(e->device));
  if (!base_off) {
    // compact form: only the lengths travel; read r starts at the next multiple of 4 bases after read r-1
    if (n_words * 16 > e->batch_bases + 64) return fail(e, SQ_ERR_ARG, "batch without base_off exceeds option batch_bases (%llu bases): push smaller batches", (unsigned long long)e->batch_bases);
    Slot* s = nullptr;
    SQ_TRY(acquire_slot(e, &s));
    SQ_CUDA(e, s->packed.ensure(((n_words + 3) & ~3ull) * 4 + 64));
    SQ_CUDA(e, s->base_off.ensure(((size_t)n_reads + 1) * 4));
    SQ_CUDA(e, s->len.ensure((size_t)n_reads * 4));
    SQ_CUDA(e, cudaMemcpyAsync(s->packed.p, packed_words, n_words * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaMemcpyAsync(s->len.p, len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaEventRecord(s->copied, e->copy_stream));
    SQ_TRY(run_batch(e, *s, s->packed.as<uint32_t>(), (n_words + 3) & ~3ull, s->base_off.as<uint32_t>(), 0,
                     s->len.as<uint32_t>(), n_reads, n_words * 16, s->copied, s->base_off.as<uint32_t>()));
    SQ_CUDA(e, cudaEventSynchronize(s->copied));
    return SQ_OK;
  }
  uint32_t r0 = 0;
  while (r0 < n_reads) {
    // sub-batch [r0, r1): bases from word-aligned start of read r0, at most batch_bases
    const uint64_t w0 = base_off[r0] >> 4;
    const uint64_t limit = w0 * 16 + e->batch_bases;
    uint32_t lo = r0 + 1, hi = n_reads;  // largest r1 with end(r1-1) <= limit (reads are laid out in order)
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo + 1) / 2;
      if ((uint64_t)base_off[mid - 1] + len[mid - 1] <= limit) lo = mid; else hi = mid - 1;
    }
    const uint32_t r1 = lo;
    const uint64_t end_base = (uint64_t)base_off[r1 - 1] + len[r1 - 1];
    if (end_base < w0 * 16) return fail(e, SQ_ERR_ARG, "base_off must be non-decreasing");
    const uint64_t w1 = std::min<uint64_t>((end_base + 15) >> 4, n_words);
    if (((end_base + 15) >> 4) > n_words) return fail(e, SQ_ERR_ARG, "read %u extends beyond n_words", r1 - 1);
    const uint64_t nw = w1 - w0, nb = end_base - w0 * 16;
    const uint32_t nr = r1 - r0;
    Slot* s = nullptr;
    SQ_TRY(acquire_slot(e, &s));
    SQ_CUDA(e, s->packed.ensure(((nw + 3) & ~3ull) * 4 + 64));
    SQ_CUDA(e, s->base_off.ensure((size_t)nr * 4));
    SQ_CUDA(e, s->len.ensure((size_t)nr * 4));
    SQ_CUDA(e, cudaMemcpyAsync(s->packed.p, packed_words + w0, nw * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaMemcpyAsync(s->base_off.p, base_off + r0, (size_t)nr * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaMemcpyAsync(s->len.p, len + r0, (size_t)nr * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaEventRecord(s->copied, e->copy_stream));
    SQ_TRY(run_batch(e, *s, s->packed.as<uint32_t>(), (nw + 3) & ~3ull, s->base_off.as<uint32_t>(), (uint32_t)(w0 * 16),
                     s->len.as<uint32_t>(), nr, nb, s->copied));
    // the caller may reuse its buffers when we return: wait for the copies (not for the kernels)
    SQ_CUDA(e, cudaEventSynchronize(s->copied));
    r0 = r1;
  }
  return SQ_OK;
}

int sq_push_reads_fixed(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, uint32_t read_len,
                        uint32_t n_reads) {
  if (!e) return SQ_ERR_ARG;
  if (n_reads == 0) return SQ_OK;
  if (!packed_words) return fail(e, SQ_ERR_ARG, "NULL host pointer");
  const uint64_t stride = ((uint64_t)read_len + 3) & ~3ull;
  const uint64_t bases = stride * n_reads;
  if (stride == 0 || bases >= 0xFFFFFFFFull) return fail(e, SQ_ERR_ARG, "batch too large");
  if ((bases + 15) / 16 > n_words) return fail(e, SQ_ERR_ARG, "%u reads of %u bases need %llu words, got %llu", n_reads, read_len,
                                                (unsigned long long)((bases + 15) / 16), (unsigned long long)n_words);
  if (bases > e->batch_bases + 64) return fail(e, SQ_ERR_ARG, "batch exceeds option batch_bases (%llu bases): push smaller batches", (unsigned long long)e->batch_bases);
  SQ_CUDA(e, cudaSetDevice(e->device));
  // Sub-batches of about sub_batch_reads reads (a multiple of 4 reads is a whole number of packed words): the
  // kernels of one overlap the copy of the next, so less work is left when the last copy lands.
  const uint32_t sub = std::max<uint32_t>(4, e->sub_batch_reads & ~3u);
  const uint32_t parts = n_reads <= sub + sub / 2 ? 1u : (n_reads + sub - 1) / sub;
  const uint32_t per = ((n_reads + parts - 1) / parts + 3) & ~3u;
  for (uint32_t r0 = 0; r0 < n_reads; r0 += per) {
    const uint32_t nr = std::min(per, n_reads - r0);
    const uint64_t w0 = stride * r0 / 16, nb = stride * nr, nw = (nb + 15) / 16;
    Slot* s = nullptr;
    SQ_TRY(acquire_slot(e, &s));
    SQ_CUDA(e, s->packed.ensure(((nw + 3) & ~3ull) * 4 + 64));
    SQ_CUDA(e, s->base_off.ensure(((size_t)nr + 1) * 4));
    SQ_CUDA(e, s->len.ensure((size_t)nr * 4));
    SQ_CUDA(e, cudaMemcpyAsync(s->packed.p, packed_words + w0, nw * 4, cudaMemcpyHostToDevice, e->copy_stream));
    SQ_CUDA(e, cudaEventRecord(s->copied, e->copy_stream));
    SQ_TRY(run_batch(e, *s, s->packed.as<uint32_t>(), (nw + 3) & ~3ull, s->base_off.as<uint32_t>(), 0,
                     s->len.as<uint32_t>(), nr, nb, s->copied, s->base_off.as<uint32_t>(), read_len));
    // the caller may reuse its buffer when we return: wait for the last copy (not for the kernels); the copies
    // in between are queued back to back
    if (r0 + per >= n_reads) SQ_CUDA(e, cudaEventSynchronize(s->copied));
  }
  return SQ_OK;
}

void* sq_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void sq_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int sq_sync(sq_engine* e) {
  if (!e) return SQ_ERR_ARG;
  SQ_CUDA(e, cudaSetDevice(e->device));
  SQ_CUDA(e, cudaStreamSynchronize(e->copy_stream));
  // at most one batch is waiting for its candidate count; finalize in age order anyway
  SQ_TRY(finalize_slot(e, e->slot[e->next_slot]));
  SQ_TRY(finalize_slot(e, e->slot[e->next_slot ^ 1]));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  SQ_CUDA(e, cudaStreamSynchronize(e->tail_stream));
  resolve_events(e);
  return check_flags(e);
}

int sq_reset_reads(sq_engine* e) {
  if (!e) return SQ_ERR_ARG;
  SQ_TRY(sq_sync(e));
  SQ_CUDA(e, cudaMemsetAsync(e->d_totals, 0, 256, e->stream));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  memset(e->h_mirror, 0, 128);
  e->P = 0;
  e->ovf_total = 0;
  e->slow_total = 0;
  e->mid_total = 0;
  e->keys_valid = true;
  e->n_reads = e->n_bases = e->n_batches = 0;
  for (int i = 0; i < 10; ++i) { e->ms[i] = 0; e->n_stage[i] = 0; }
  return SQ_OK;
}

int sq_num_pairs(sq_engine* e, uint64_t* n_reads, uint64_t* n_pairs) {
  if (!e) return SQ_ERR_ARG;
  SQ_TRY(sq_sync(e));
  if (n_reads) *n_reads = e->n_reads;
  if (n_pairs) *n_pairs = e->P;
  return SQ_OK;
}

int sq_get_candidates(sq_engine* e, uint64_t* read_off, uint32_t* tid, int32_t* score) {
  if (!e) return SQ_ERR_ARG;
  SQ_TRY(sq_sync(e));
  const uint64_t R = e->n_reads, P = e->P;
  // the store keeps the lists in the order the vote kernels finished them: gather them in read order
  std::vector<uint32_t> tmp(R + 1, 0);
  SQ_CUDA(e, e->em_off.ensure((R + 2) * 4));
  SQ_CUDA(e, e->em_tid.ensure((P + 1) * 4));
  SQ_CUDA(e, e->em_score.ensure((P + 1) * 4));
  SQ_CUDA(e, e->scan_tmp.ensure(scan_tmp_words((uint32_t)(R + 1)) * 4));
  if (R) {
    SQ_CUDA(e, e->em_cnt.ensure((R + 2) * 4));
    launch_csr_gather(e->rd, e->em_cnt.as<uint32_t>(), e->em_off.as<uint32_t>(), R, e->scan_tmp.as<uint32_t>(), e->cand,
                      e->em_tid.as<uint32_t>(), e->em_score.as<int32_t>(), e->stream, &e->launches);
    SQ_CUDA(e, cudaMemcpyAsync(tmp.data(), e->em_off.p, R * 4, cudaMemcpyDeviceToHost, e->stream));
    SQ_CUDA(e, cudaStreamSynchronize(e->stream));
    tmp[R] = (uint32_t)P;
  }
  if (read_off)
    for (uint64_t i = 0; i <= R; ++i) read_off[i] = tmp[i];
  if (P && tid) SQ_CUDA(e, cudaMemcpy(tid, e->em_tid.p, P * 4, cudaMemcpyDeviceToHost));
  if (P && score) SQ_CUDA(e, cudaMemcpy(score, e->em_score.p, P * 4, cudaMemcpyDeviceToHost));
  if (P && tid && e->perm_ready && !e->perm_identity) {
    // the store holds internal ids, ordered (score desc, internal id asc): hand out the caller's ids, equal
    // scores in ascending order of those
    std::vector<int32_t> sc_own;
    const int32_t* sc = score;
    if (!sc) {
      sc_own.resize(P);
      SQ_CUDA(e, cudaMemcpy(sc_own.data(), e->em_score.p, P * 4, cudaMemcpyDeviceToHost));
      sc = sc_own.data();
    }
    for (uint64_t i = 0; i < P; ++i) tid[i] = e->ext_of[tid[i]];
    for (uint64_t r = 0; r < R; ++r)
      for (uint32_t a = tmp[r]; a < tmp[r + 1];) {
        uint32_t b = a + 1;
        while (b < tmp[r + 1] && sc[b] == sc[a]) ++b;
        if (b - a > 1) std::sort(tid + a, tid + b);
        a = b;
      }
  }
  return SQ_OK;
}

int sq_set_candidates(sq_engine* e, uint64_t n_reads, const uint64_t* read_off, const uint32_t* tid,
                      const int32_t* score) {
  if (!e) return SQ_ERR_ARG;
  if (n_reads && !read_off) return fail(e, SQ_ERR_ARG, "NULL read_off");
  SQ_TRY(sq_reset_reads(e));
  const uint64_t P = n_reads ? read_off[n_reads] : 0;
  if (P && (!tid || !score)) return fail(e, SQ_ERR_ARG, "NULL candidate array");
  for (uint64_t i = 0; i < P; ++i)
    if (tid[i] >= e->T) return fail(e, SQ_ERR_ARG, "candidate %llu names transcript %u >= T", (unsigned long long)i, tid[i]);
  if (!e->perm_ready) {  // no index yet: the numbering is the caller's from here on
    const uint32_t T = (uint32_t)e->T;
    e->ext_of.resize(T);
    e->int_of.resize(T);
    for (uint32_t i = 0; i < T; ++i) e->ext_of[i] = e->int_of[i] = i;
    SQ_CUDA(e, e->d_ext_of.ensure((size_t)T * 4));
    SQ_CUDA(e, cudaMemcpy(e->d_ext_of.p, e->ext_of.data(), (size_t)T * 4, cudaMemcpyHostToDevice));
    e->perm_identity = true;
    e->perm_ready = true;
  }
  SQ_TRY(ensure_store(e, 0, n_reads, P));
  std::vector<uint32_t> start32(n_reads + 1, 0), cnt32(n_reads + 1, 0);
  for (uint64_t i = 0; i < n_reads; ++i) {
    if (read_off[i + 1] < read_off[i]) return fail(e, SQ_ERR_ARG, "read_off must be non-decreasing");
    start32[i] = (uint32_t)read_off[i];
    cnt32[i] = (uint32_t)(read_off[i + 1] - read_off[i]);
  }
  if (n_reads) {
    std::vector<uint2> loc(n_reads);
    for (uint64_t i = 0; i < n_reads; ++i) loc[i] = make_uint2(start32[i], cnt32[i]);
    SQ_CUDA(e, cudaMemcpy(e->rd, loc.data(), n_reads * sizeof(uint2), cudaMemcpyHostToDevice));
  }
  if (P) {
    std::vector<uint2> pairs(P);
    for (uint64_t i = 0; i < P; ++i) pairs[i] = make_uint2(e->perm_identity ? tid[i] : e->int_of[tid[i]], (uint32_t)score[i]);
    SQ_CUDA(e, cudaMemcpy(e->cand, pairs.data(), P * sizeof(uint2), cudaMemcpyHostToDevice));
  }
  e->P = P;
  e->n_reads = n_reads;
  e->keys_valid = false;  // no vote ran: sq_finish computes the class keys itself
  return SQ_OK;
}

static int allreduce(sq_engine* e, void* buf, size_t n, ncclDataType_t dt) {
  if (!e->comm) return SQ_OK;
  ncclResult_t r = g_nccl.AllReduce(buf, buf, n, dt, ncclSum, e->comm, e->stream);
  if (r != ncclSuccess) return fail(e, SQ_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  return SQ_OK;
}

int sq_finish(sq_engine* e, uint64_t R_total, int em_iters, double em_tol, double* pi, double* numreads,
              uint8_t* present, int* iters_done) {
  if (!e) return SQ_ERR_ARG;
  if (!pi || !numreads || !present) return fail(e, SQ_ERR_ARG, "NULL output array");
  SQ_TRY(sq_sync(e));
  const uint64_t R = e->n_reads, P = e->P;
  const uint32_t T = (uint32_t)e->T;
  cudaStream_t st = e->stream;
  SQ_CUDA(e, e->misc.ensure(256));
  uint32_t* state = e->misc.as<uint32_t>();                                // [0..1]
  double* last_change = reinterpret_cast<double*>(e->misc.as<char>() + 16);
  unsigned long long* d_R = reinterpret_cast<unsigned long long*>(e->misc.as<char>() + 32);

  uint64_t R_all = R_total;
  if (R_all == 0) {
    R_all = R;
    if (e->comm) {
      unsigned long long v = R;
      SQ_CUDA(e, cudaMemcpyAsync(d_R, &v, 8, cudaMemcpyHostToDevice, st));
      SQ_TRY(allreduce(e, d_R, 1, ncclUint64));
      SQ_CUDA(e, cudaMemcpyAsync(&v, d_R, 8, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaStreamSynchronize(st));
      R_all = v;
    }
  }

  // ---- transcript-major copy of the pairs: stable radix sort on the transcript id ----
  uint64_t* keys = nullptr;
  uint32_t* vals = nullptr;
  uint32_t n_seg = 0, n_classes = 0, n_cpairs = 0, pack_bad = 0;
  bool packed = false;
  uint32_t* d_pack_bad = reinterpret_cast<uint32_t*>(e->misc.as<char>() + 128);
  {
    StageScope sc(e, 3);
    SQ_CUDA(e, e->toff.ensure(((size_t)T + 1) * 4));
    SQ_CUDA(e, e->nseg.ensure((size_t)T * 4));
    SQ_CUDA(e, e->seg_off.ensure(((size_t)T + 1) * 4));
    SQ_CUDA(e, e->scan_tmp.ensure(scan_tmp_words(T) * 4));
    SQ_CUDA(e, e->keys_a.ensure((P + 1) * 8));
    SQ_CUDA(e, e->keys_b.ensure((P + 1) * 8));
    SQ_CUDA(e, e->vals_a.ensure((P + 1) * 4));
    SQ_CUDA(e, e->vals_b.ensure((P + 1) * 4));
    SQ_CUDA(e, e->tm_read.ensure((P + 1) * 4));
    SQ_CUDA(e, e->sort_tmp.ensure(radix_tmp_words(P) * 4));
    // equivalence classes of reads (identical candidate lists), ordered by best candidate (see sq_em.cu)
    SQ_CUDA(e, e->em_off.ensure((R + 2) * 4));
    SQ_CUDA(e, e->em_cnt.ensure((R + 2) * 4));
    SQ_CUDA(e, e->em_tid.ensure((P + 1) * 4));
    SQ_CUDA(e, e->em_score.ensure((P + 1) * 4));
    SQ_CUDA(e, e->em_pack.ensure((P + 1) * 4));
    SQ_CUDA(e, e->cls_head.ensure((R + 2) * 4));
    SQ_CUDA(e, e->cls_id.ensure((R + 2) * 4));
    SQ_CUDA(e, e->cls_read.ensure((R + 2) * 4));
    SQ_CUDA(e, e->cls_pos.ensure((R + 2) * 4));
    SQ_CUDA(e, e->cls_weight.ensure((R + 2) * 8));
    if (R) {
      SQ_CUDA(e, e->keys_a.ensure((std::max(P, R) + 1) * 8));
      SQ_CUDA(e, e->keys_b.ensure((std::max(P, R) + 1) * 8));
      SQ_CUDA(e, e->vals_a.ensure((std::max(P, R) + 1) * 4));
      SQ_CUDA(e, e->vals_b.ensure((std::max(P, R) + 1) * 4));
      SQ_CUDA(e, e->sort_tmp.ensure(radix_tmp_words(std::max(P, R)) * 4));
      SQ_CUDA(e, e->scan_tmp.ensure(scan_tmp_words((uint32_t)std::max<uint64_t>(T, R + 1)) * 4));
      const uint32_t top_bits = std::max<uint32_t>(1, log2_ceil((uint64_t)T + 1));
      const uint32_t hash_bits = e->class_hash_bits;
      const void* fp = e->rfp;
      // keys and fingerprints were produced batch by batch behind the votes (sq_set_candidates: here)
      if (!e->keys_valid)
        launch_read_keys(e->rd, 0, R, e->cand, T, hash_bits, e->rkey, e->rfp, st,
                         &e->launches);
      SQ_CUDA(e, cudaMemcpyAsync(e->keys_a.p, e->rkey, R * 8, cudaMemcpyDeviceToDevice, st));  // the sort works on a copy
      uint64_t* skeys = nullptr;
      uint32_t* dummy = nullptr;
      launch_radix_sort(e->keys_a.as<uint64_t>(), e->keys_b.as<uint64_t>(), nullptr, nullptr, R,
                        (int)(hash_bits + top_bits), e->sort_tmp.as<uint32_t>(), &skeys, &dummy, st, &e->launches, 32);
      launch_class_heads(skeys, R, e->rd, fp, e->cand, e->exact_classes,
                         e->cls_head.as<uint32_t>(),
                         e->cls_id.as<uint32_t>(), e->scan_tmp.as<uint32_t>(), e->cls_read.as<uint32_t>(),
                         e->cls_pos.as<uint32_t>(), e->em_cnt.as<uint32_t>(), st, &e->launches);
      SQ_CUDA(e, cudaMemcpyAsync(&n_classes, e->cls_id.as<uint32_t>() + R, 4, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaStreamSynchronize(st));
      launch_class_gather(e->cls_read.as<uint32_t>(), e->cls_pos.as<uint32_t>(), e->em_cnt.as<uint32_t>(),
                          e->em_off.as<uint32_t>(), n_classes, e->scan_tmp.as<uint32_t>(), e->rd, e->cand,
                          e->em_tid.as<uint32_t>(), e->em_score.as<int32_t>(), e->em_pack.as<uint32_t>(),
                          d_pack_bad, e->cls_weight.as<double>(), st, &e->launches);
      SQ_CUDA(e, cudaMemcpyAsync(&n_cpairs, e->em_off.as<uint32_t>() + n_classes, 4, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaMemcpyAsync(&pack_bad, d_pack_bad, 4, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaStreamSynchronize(st));
    } else {
      SQ_CUDA(e, cudaMemsetAsync(e->em_off.p, 0, 4, st));
    }
    // transcript and score of a pair in one word (24 + 8 bits) when every score, id and class index fits: the two
    // copies of the pairs the EM iterations stream are then half the size (and the sort below carries no values)
    packed = n_cpairs && !pack_bad && T <= (1u << 24) && n_classes <= (1u << 24);
    if (n_cpairs) {
      launch_make_sort_keys(e->em_off.as<uint32_t>(), n_classes, packed ? e->em_pack.as<uint32_t>() : e->em_tid.as<uint32_t>(),
                            packed, e->keys_a.as<uint64_t>(), st, &e->launches);
      if (!packed)
        SQ_CUDA(e, cudaMemcpyAsync(e->vals_a.p, e->em_score.p, (size_t)n_cpairs * 4, cudaMemcpyDeviceToDevice, st));
    }
    const int nbits = (int)std::max<uint32_t>(1, log2_ceil(T));
    launch_radix_sort(e->keys_a.as<uint64_t>(), e->keys_b.as<uint64_t>(), packed ? nullptr : e->vals_a.as<uint32_t>(),
                      packed ? nullptr : e->vals_b.as<uint32_t>(), n_cpairs, nbits, e->sort_tmp.as<uint32_t>(), &keys, &vals,
                      st, &e->launches);
    launch_tmajor(keys, n_cpairs, T, e->em_seg, e->toff.as<uint32_t>(), e->tm_read.as<uint32_t>(),
                  e->nseg.as<uint32_t>(), e->seg_off.as<uint32_t>(), e->scan_tmp.as<uint32_t>(), st, &e->launches);
    // segments: at most one per em_seg pairs plus one per transcript; the exact count stays on the device
    // (seg_off[T]), the kernels that walk the segments read it there
    n_seg = (uint32_t)(n_cpairs / e->em_seg + T + 1);
    SQ_CUDA(e, e->seg_tid.ensure(((size_t)n_seg + 1) * 4));
    SQ_CUDA(e, e->seg_begin.ensure(((size_t)n_seg + 1) * 4));
    SQ_CUDA(e, e->partial.ensure(((size_t)n_seg + 1) * 8));
    launch_seg_expand(e->toff.as<uint32_t>(), e->seg_off.as<uint32_t>(), T, e->em_seg, e->seg_tid.as<uint32_t>(),
                      e->seg_begin.as<uint32_t>(), st, &e->launches);
  }

  SQ_CUDA(e, e->pi.ensure((size_t)T * 8));
  SQ_CUDA(e, e->ps.ensure((size_t)T * 8));
  SQ_CUDA(e, e->numreads.ensure((size_t)T * 8));
  SQ_CUDA(e, e->present.ensure((size_t)T * 4));
  SQ_CUDA(e, e->read_tmp.ensure((R + 1) * 8));
  e->n_classes_last = n_classes;
  e->n_cpairs_last = n_cpairs;
  SQ_CUDA(e, e->block_change.ensure(((size_t)(T + 255) / 256 + 1) * 8));

  EmView v;
  v.read_off = e->em_off.as<uint32_t>();
  v.cand_tid = packed ? e->em_pack.as<uint32_t>() : e->em_tid.as<uint32_t>();
  v.cand_score = e->em_score.as<int32_t>();
  v.packed = packed;
  v.n_reads = n_classes;
  v.weight = e->cls_weight.as<double>();
  v.toff = e->toff.as<uint32_t>();
  v.tm_read = e->tm_read.as<uint32_t>();
  v.tm_score = vals;
  v.seg_off = e->seg_off.as<uint32_t>();
  v.seg_tid = e->seg_tid.as<uint32_t>();
  v.seg_begin = e->seg_begin.as<uint32_t>();
  v.n_seg = n_seg;
  v.seg = e->em_seg;
  v.n_pairs = n_cpairs;
  v.T = T;
  // measured: the one-shot exchange beats ncclAllReduce for two and four ranks and loses for eight (every rank reads
  // every peer's whole vector); option peer_exchange = 2 forces it for any number of ranks
  const bool peer = e->comm && e->peer_ok && (e->peer_wanted == 2 || (e->peer_wanted == 1 && e->nranks <= 4));
  e->peer_used = peer;
  v.pi = e->pi.as<double>();
  v.ps = e->ps.as<double>();
  v.read_tmp = e->read_tmp.as<double>();
  v.partial = e->partial.as<double>();
  v.block_change = e->block_change.as<double>();
  v.last_change = last_change;
  v.state = state;

  // M-step constants (isoform_assignment.cpp:54-57): float pseudocount, float division by R
  const float pseudocount = 0.01f;
  const double add_a = (double)(pseudocount / (float)R_all);
  const double add_b = (double)pseudocount;
  {
    StageScope sc(e, 4);
    launch_em_init(v.pi, T, state, st, &e->launches);
    // The exchange kernels spin on their peers' flags with a time-out (a dead peer must not hang the GPU).  Ranks
    // may arrive here seconds apart (uneven shards, slower hosts): one small all-reduce lines them up first -- NCCL
    // waits as long as it takes -- and from there on the iterations run in lock-step.
    if (peer) {
      SQ_CUDA(e, cudaMemsetAsync(d_R, 0, 8, st));
      SQ_TRY(allreduce(e, d_R, 1, ncclUint64));
    }
    for (int it = 0; it < em_iters; ++it) {
      if (peer) {
        // this rank's sums go to its exchange slot; the M-step kernel adds all ranks' slots over peer memory
        launch_em_estep(v, st, &e->launches, false);
        launch_em_mstep_peer(v, e->d_peer_ps.as<double*>(), e->d_peer_flags.as<unsigned long long*>(), (uint32_t)e->rank,
                             (uint32_t)e->nranks, ++e->epoch, add_a, add_b, em_tol, e->d_peer_err.as<uint32_t>(), st,
                             &e->launches);
      } else if (e->comm) {
        launch_em_estep(v, st, &e->launches, true);
        SQ_TRY(allreduce(e, v.ps, T, ncclDouble));
        launch_em_mstep(v, add_a, add_b, em_tol, st, &e->launches);
      } else {
        launch_em_estep(v, st, &e->launches, false);
        launch_em_mstep_fused(v, add_a, add_b, em_tol, st, &e->launches);
      }
    }
  }
  {
    StageScope sc(e, 5);
    launch_assign(v, e->numreads.as<double>(), e->present.as<uint32_t>(), st, &e->launches);
    SQ_TRY(allreduce(e, e->numreads.p, T, ncclDouble));
    SQ_TRY(allreduce(e, e->present.p, T, ncclUint32));
  }
  // results leave in the caller's transcript numbering
  SQ_CUDA(e, e->out_pi.ensure((size_t)T * 8));
  SQ_CUDA(e, e->out_nr.ensure((size_t)T * 8));
  SQ_CUDA(e, e->out_present.ensure((size_t)T));
  launch_permute_out(v.pi, e->numreads.as<double>(), e->present.as<uint32_t>(),
                     e->perm_ready && !e->perm_identity ? e->d_ext_of.as<uint32_t>() : nullptr, T, e->out_pi.as<double>(),
                     e->out_nr.as<double>(), e->out_present.as<uint8_t>(), st, &e->launches);
  uint32_t st_host[2] = {0, 0};
  SQ_CUDA(e, cudaMemcpyAsync(pi, e->out_pi.p, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaMemcpyAsync(numreads, e->out_nr.p, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaMemcpyAsync(present, e->out_present.p, (size_t)T, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaMemcpyAsync(st_host, state, 8, cudaMemcpyDeviceToHost, st));
  uint32_t peer_err = 0;
  if (e->peer_ok) SQ_CUDA(e, cudaMemcpyAsync(&peer_err, e->d_peer_err.p, 4, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaStreamSynchronize(st));
  SQ_CUDA(e, cudaGetLastError());
  if (peer_err) return fail(e, SQ_ERR_NCCL, "a peer rank did not deliver its sums within the exchange time-out");
  e->em_iterations = (int)st_host[1];
  if (iters_done) *iters_done = e->em_iterations;
  resolve_events(e);
  return SQ_OK;
}

int sq_get_stats(sq_engine* e, sq_stats* out) {
  if (!e || !out) return SQ_ERR_ARG;
  SQ_TRY(sq_sync(e));
  unsigned long long tot[6] = {0, 0, 0, 0, 0, 0};
  SQ_CUDA(e, cudaMemcpy(tot, e->d_totals, sizeof(tot), cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof(*out));
  out->reads = e->n_reads;
  out->bases = tot[5];
  out->kmers = tot[4];
  out->sketch_hashes = tot[0];
  out->pairs = e->P;
  out->overflow_reads = e->ovf_total;
  out->batches = e->n_batches;
  out->em_iterations = e->em_iterations;
  out->peer_exchange = e->peer_used ? 1 : 0;
  out->ms_sketch = e->ms[0]; out->ms_vote = e->ms[1]; out->ms_compact = e->ms[2];
  out->ms_sort = e->ms[3]; out->ms_em = e->ms[4]; out->ms_assign = e->ms[5];
  out->launches = e->launches;
  out->queries = tot[1]; out->hits = tot[2]; out->postings = tot[3];
  out->ms_items = e->ms[6];
  out->ms_vote_main = e->ms[7];
  out->ms_lookup = e->ms[8];
  out->sketch_launches = e->n_stage[0]; out->vote_launches = e->n_stage[1];
  out->slow_reads = e->slow_total;
  out->em_classes = e->n_classes_last;
  out->em_class_pairs = e->n_cpairs_last;
  out->mid_reads = e->mid_total;
  return SQ_OK;
}

int sq_nccl_unique_id(uint8_t id[SQ_NCCL_ID_BYTES]) {
  std::string err;
  if (!g_nccl.load(&err)) return fail(nullptr, SQ_ERR_NCCL, "%s", err.c_str());
  ncclUniqueId uid;
  static_assert(sizeof(uid) == SQ_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclResult_t r = g_nccl.GetUniqueId(&uid);
  if (r != ncclSuccess) return fail(nullptr, SQ_ERR_NCCL, "ncclGetUniqueId failed (%d)", (int)r);
  memcpy(id, &uid, SQ_NCCL_ID_BYTES);
  return SQ_OK;
}

// Map every rank's exchange buffers into every other rank (CUDA IPC handles travel through the communicator).  Not
// being able to (ranks that are threads of one process cannot open each other's handles, no peer access, an old
// libnccl) is not an error: all ranks agree, by a min all-reduce, to keep the per-iteration ncclAllReduce then.
static int setup_peer_exchange(sq_engine* e) {
  const int N = e->nranks;
  e->peer_ok = false;
  if (N < 2 || N > 32 || !g_nccl.AllGather) return SQ_OK;
  struct Card { cudaIpcMemHandle_t ps, flags; int ok; int pad; };
  const size_t T = (size_t)e->T;
  Card mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = 1;
  const size_t x_bytes = 2 * T * sizeof(double), f_bytes = 2 * (size_t)N * 8;
  if (cudaMalloc(&e->xbuf, x_bytes) != cudaSuccess || cudaMalloc(&e->xflags, f_bytes) != cudaSuccess) mine.ok = 0;
  if (mine.ok) {
    cudaMemset(e->xbuf, 0, x_bytes);
    cudaMemset(e->xflags, 0, f_bytes);
    if (cudaIpcGetMemHandle(&mine.ps, e->xbuf) != cudaSuccess || cudaIpcGetMemHandle(&mine.flags, e->xflags) != cudaSuccess) mine.ok = 0;
  }
  cudaGetLastError();
  DevBuf d_cards;
  SQ_CUDA(e, d_cards.ensure(sizeof(Card) * (size_t)(N + 1)));
  Card* d_all = d_cards.as<Card>();
  SQ_CUDA(e, cudaMemcpyAsync(d_all + N, &mine, sizeof(Card), cudaMemcpyHostToDevice, e->stream));
  ncclResult_t r = g_nccl.AllGather(d_all + N, d_all, sizeof(Card), ncclChar, e->comm, e->stream);
  if (r != ncclSuccess) { d_cards.release(); return fail(e, SQ_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); }
  std::vector<Card> all((size_t)N);
  SQ_CUDA(e, cudaMemcpyAsync(all.data(), d_all, sizeof(Card) * (size_t)N, cudaMemcpyDeviceToHost, e->stream));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  int ok = mine.ok;
  for (int i = 0; i < N; ++i) ok &= all[i].ok;
  std::vector<double*> ps((size_t)N, nullptr);
  std::vector<unsigned long long*> fl((size_t)N, nullptr);
  for (int i = 0; i < N && ok; ++i) {
    if (i == e->rank) { ps[i] = e->xbuf; fl[i] = e->xflags; continue; }
    void *a = nullptr, *b = nullptr;
    if (cudaIpcOpenMemHandle(&a, all[i].ps, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    e->peer_mapped.push_back(a);
    if (cudaIpcOpenMemHandle(&b, all[i].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
    e->peer_mapped.push_back(b);
    ps[i] = static_cast<double*>(a);
    fl[i] = static_cast<unsigned long long*>(b);
  }
  cudaGetLastError();
  // everybody or nobody
  int* d_ok = reinterpret_cast<int*>(d_all);
  SQ_CUDA(e, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, e->stream));
  r = g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, e->comm, e->stream);
  if (r != ncclSuccess) { d_cards.release(); return fail(e, SQ_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error"); }
  SQ_CUDA(e, cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  SQ_CUDA(e, cudaStreamSynchronize(e->stream));
  d_cards.release();
  if (!ok) {
    for (void* m : e->peer_mapped) cudaIpcCloseMemHandle(m);
    e->peer_mapped.clear();
    cudaGetLastError();
    return SQ_OK;
  }
  SQ_CUDA(e, e->d_peer_ps.ensure(sizeof(void*) * (size_t)N));
  SQ_CUDA(e, e->d_peer_flags.ensure(sizeof(void*) * (size_t)N));
  SQ_CUDA(e, e->d_peer_err.ensure(16));
  SQ_CUDA(e, cudaMemcpy(e->d_peer_ps.p, ps.data(), sizeof(void*) * (size_t)N, cudaMemcpyHostToDevice));
  SQ_CUDA(e, cudaMemcpy(e->d_peer_flags.p, fl.data(), sizeof(void*) * (size_t)N, cudaMemcpyHostToDevice));
  SQ_CUDA(e, cudaMemset(e->d_peer_err.p, 0, 16));
  e->peer_ok = true;
  return SQ_OK;
}

int sq_comm_init(sq_engine* e, int nranks, int rank, const uint8_t id[SQ_NCCL_ID_BYTES]) {
  if (!e) return SQ_ERR_ARG;
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(e, SQ_ERR_ARG, "bad rank %d of %d", rank, nranks);
  if (nranks == 1) { e->nranks = 1; e->rank = 0; return SQ_OK; }
  std::string err;
  if (!g_nccl.load(&err)) return fail(e, SQ_ERR_NCCL, "%s", err.c_str());
  SQ_CUDA(e, cudaSetDevice(e->device));
  ncclUniqueId uid;
  memcpy(&uid, id, SQ_NCCL_ID_BYTES);
  ncclResult_t r = g_nccl.CommInitRank(&e->comm, nranks, uid, rank);
  if (r != ncclSuccess) return fail(e, SQ_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  e->nranks = nranks;
  e->rank = rank;
  return setup_peer_exchange(e);
}

}  // extern "C"

// ---- sketch tap / postings build: run the sketch kernel on a host batch held in the engine's tap slot ----
namespace {

// uploads the batch and runs items + sketch for k-indices [k0, k0+nk); leaves counts per (read, k) in
// e->tap_counts and their exclusive scan in e->tap_offs (total -> *total)
int tap_sketch(sq_engine* e, uint32_t k0, uint32_t nk, const uint32_t* packed_words, uint64_t n_words,
               const uint32_t* base_off, const uint32_t* len, uint32_t n_reads, uint32_t* bias_out,
               uint32_t* items_ub_out, uint64_t* stride_out, uint64_t* total) {
  Slot& s = e->tap;
  cudaStream_t st = e->stream;
  const uint64_t w0 = base_off[0] >> 4;
  uint64_t end_base = 0;
  for (uint32_t r = 0; r < n_reads; ++r) end_base = std::max<uint64_t>(end_base, (uint64_t)base_off[r] + len[r]);
  if (((end_base + 15) >> 4) > n_words) return fail(e, SQ_ERR_ARG, "a sequence extends beyond n_words");
  if (end_base < w0 * 16) return fail(e, SQ_ERR_ARG, "base_off must be non-decreasing");
  const uint64_t nw = ((end_base + 15) >> 4) - w0, nb = end_base - w0 * 16;
  const uint64_t items_ub64 = (uint64_t)n_reads + nb / SQ_CHUNK + 1;
  if (items_ub64 >= 0xFFFFFFFFull || nb >= 0xFFFFFFFFull) return fail(e, SQ_ERR_ARG, "batch too large");
  const uint32_t items_ub = (uint32_t)items_ub64;
  const uint64_t stride = ((nb + 3) & ~3ull) + 4;
  SQ_CUDA(e, s.packed.ensure(((nw + 3) & ~3ull) * 4 + 64));
  SQ_CUDA(e, s.base_off.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.len.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.nit.ensure((size_t)n_reads * 4));
  SQ_CUDA(e, s.item_start.ensure(((size_t)n_reads + 1) * 4));
  SQ_CUDA(e, s.item_read.ensure((size_t)items_ub * 4));
  SQ_CUDA(e, s.cnt.ensure((size_t)items_ub * nk * 2));
  SQ_CUDA(e, s.hsel.ensure((size_t)stride * nk * 4));
  SQ_CUDA(e, s.hoff.ensure((size_t)items_ub * nk * 4));
  SQ_CUDA(e, s.scan_tmp.ensure(scan_tmp_words(std::max<uint32_t>(items_ub, n_reads * nk)) * 4));
  SQ_CUDA(e, e->tap_counts.ensure((size_t)n_reads * nk * 4));
  SQ_CUDA(e, e->tap_offs.ensure(((size_t)n_reads * nk + 1) * 4));
  SQ_CUDA(e, cudaMemcpyAsync(s.packed.p, packed_words + w0, nw * 4, cudaMemcpyHostToDevice, st));
  SQ_CUDA(e, cudaMemcpyAsync(s.base_off.p, base_off, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
  SQ_CUDA(e, cudaMemcpyAsync(s.len.p, len, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
  KList kl;
  kl.nk = 0;
  launch_items(s.len.as<uint32_t>(), n_reads, s.nit.as<uint32_t>(), s.item_start.as<uint32_t>(),
               s.item_read.as<uint32_t>(), items_ub, s.scan_tmp.as<uint32_t>(), st, &e->launches, kl, nullptr);
  SQ_CUDA(e, cudaMemsetAsync(s.cnt.p, 0, (size_t)items_ub * nk * 2, st));
  SketchParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.packed = s.packed.as<uint32_t>();
  sp.n_words = (nw + 3) & ~3ull;
  sp.base_off = s.base_off.as<uint32_t>();
  sp.bias = (uint32_t)(w0 * 16);
  sp.len = s.len.as<uint32_t>();
  sp.item_read = s.item_read.as<uint32_t>();
  sp.item_start = s.item_start.as<uint32_t>();
  sp.n_reads = n_reads;
  sp.n_items_ub = items_ub;
  sp.nk = nk;
  sp.threshold = e->threshold;
  sp.kmax = 0;
  for (uint32_t i = 0; i < nk; ++i) {
    sp.ks[i] = e->ks[k0 + i];
    sp.lut[i] = e->lut[k0 + i];
    sp.kmax = std::max(sp.kmax, sp.ks[i]);
  }
  sp.hsel = s.hsel.as<uint32_t>();
  sp.hstride = stride;
  sp.hoff = s.hoff.as<uint32_t>();
  sp.cnt = s.cnt.as<uint16_t>();
  sp.cursor = e->d_hcur + 16;
  sketch_smem_plan(e, nb, n_reads, &sp.cap, &sp.stage_words);
  sp.dedup = 0;  // the tap hands out the multiset, the postings build removes repeats after its sort
  sp.stats = nullptr;
  SQ_CUDA(e, cudaMemsetAsync(sp.cursor, 0, 32, st));
  launch_sketch(sp, st, &e->launches);
  launch_tap_count(s.item_start.as<uint32_t>(), n_reads, nk, s.cnt.as<uint16_t>(), items_ub,
                   e->tap_counts.as<uint32_t>(), st, &e->launches);
  launch_exclusive_scan(e->tap_counts.as<uint32_t>(), e->tap_offs.as<uint32_t>(), n_reads * nk,
                        s.scan_tmp.as<uint32_t>(), st, &e->launches);
  uint32_t tot32 = 0;
  SQ_CUDA(e, cudaMemcpyAsync(&tot32, e->tap_offs.as<uint32_t>() + (size_t)n_reads * nk, 4, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaStreamSynchronize(st));
  SQ_CUDA(e, cudaGetLastError());
  *bias_out = sp.bias;
  *items_ub_out = items_ub;
  *stride_out = stride;
  *total = tot32;
  return SQ_OK;
}

}  // namespace

extern "C" {

int sq_sketch(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, const uint32_t* base_off,
              const uint32_t* len, uint32_t n_reads, uint32_t* counts, uint32_t* hashes, uint64_t cap,
              uint64_t* total) {
  if (!e) return SQ_ERR_ARG;
  if (total) *total = 0;
  if (n_reads == 0) return SQ_OK;
  if (!packed_words || !base_off || !len || !counts) return fail(e, SQ_ERR_ARG, "NULL pointer");
  SQ_CUDA(e, cudaSetDevice(e->device));
  uint32_t bias = 0, items_ub = 0;
  uint64_t stride = 0, tot = 0;
  SQ_TRY(tap_sketch(e, 0, e->nk, packed_words, n_words, base_off, len, n_reads, &bias, &items_ub, &stride, &tot));
  Slot& s = e->tap;
  cudaStream_t st = e->stream;
  const uint64_t ncopy = std::min<uint64_t>(tot, cap);
  SQ_CUDA(e, e->tap_out.ensure((ncopy + 1) * 4));
  launch_tap_gather(s.item_start.as<uint32_t>(), n_reads, e->nk, s.cnt.as<uint16_t>(), items_ub, s.hsel.as<uint32_t>(),
                    stride, s.hoff.as<uint32_t>(), e->tap_offs.as<uint32_t>(), ncopy, e->tap_out.as<uint32_t>(), nullptr,
                    nullptr, 0, st, &e->launches);
  SQ_CUDA(e, cudaMemcpyAsync(counts, e->tap_counts.p, (size_t)n_reads * e->nk * 4, cudaMemcpyDeviceToHost, st));
  if (ncopy && hashes) SQ_CUDA(e, cudaMemcpyAsync(hashes, e->tap_out.p, ncopy * 4, cudaMemcpyDeviceToHost, st));
  SQ_CUDA(e, cudaStreamSynchronize(st));
  SQ_CUDA(e, cudaGetLastError());
  if (total) *total = tot;
  return SQ_OK;
}

int sq_build_postings(sq_engine* e, uint32_t kidx, const uint32_t* packed_words, uint64_t n_words,
                      const uint32_t* base_off, const uint32_t* len, const uint32_t* seq_tid, uint32_t n_seqs,
                      uint64_t* nkeys, uint64_t* npost, uint32_t* keys, uint64_t* post_off, uint32_t* post_tid) {
  if (!e) return SQ_ERR_ARG;
  if (kidx >= e->nk) return fail(e, SQ_ERR_ARG, "kidx %u out of range", kidx);
  if (!nkeys || !npost) return fail(e, SQ_ERR_ARG, "nkeys/npost must not be NULL");
  SQ_CUDA(e, cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  if (!keys) {  // first call: compute and cache on the device
    e->bp_kidx = -1;
    e->bp_nkeys = e->bp_npost = 0;
    if (n_seqs) {
      if (!packed_words || !base_off || !len || !seq_tid) return fail(e, SQ_ERR_ARG, "NULL pointer");
      for (uint32_t i = 0; i < n_seqs; ++i)
        if (seq_tid[i] >= e->T) return fail(e, SQ_ERR_ARG, "seq_tid[%u]=%u >= T", i, seq_tid[i]);
      uint32_t bias = 0, items_ub = 0;
      uint64_t stride = 0, tot = 0;
      SQ_TRY(tap_sketch(e, kidx, 1, packed_words, n_words, base_off, len, n_seqs, &bias, &items_ub, &stride, &tot));
      Slot& s = e->tap;
      const uint32_t tbits = std::max<uint32_t>(1, log2_ceil(e->T));
      SQ_CUDA(e, e->tap_tid.ensure((size_t)n_seqs * 4));
      SQ_CUDA(e, cudaMemcpyAsync(e->tap_tid.p, seq_tid, (size_t)n_seqs * 4, cudaMemcpyHostToDevice, st));
      SQ_CUDA(e, e->keys_a.ensure((tot + 1) * 8));
      SQ_CUDA(e, e->keys_b.ensure((tot + 1) * 8));
      SQ_CUDA(e, e->sort_tmp.ensure(radix_tmp_words(tot) * 4));
      launch_tap_gather(s.item_start.as<uint32_t>(), n_seqs, 1, s.cnt.as<uint16_t>(), items_ub, s.hsel.as<uint32_t>(),
                        stride, s.hoff.as<uint32_t>(), e->tap_offs.as<uint32_t>(), tot, nullptr, e->keys_a.as<uint64_t>(),
                        e->tap_tid.as<uint32_t>(), tbits, st, &e->launches);
      uint64_t* sorted = nullptr;
      uint32_t* dummy = nullptr;
      launch_radix_sort(e->keys_a.as<uint64_t>(), e->keys_b.as<uint64_t>(), nullptr, nullptr, tot, 32 + (int)tbits,
                        e->sort_tmp.as<uint32_t>(), &sorted, &dummy, st, &e->launches);
      SQ_CUDA(e, e->bp_newpair.ensure((tot + 1) * 4));
      SQ_CUDA(e, e->bp_newkey.ensure((tot + 1) * 4));
      SQ_CUDA(e, e->bp_ppos.ensure((tot + 2) * 4));
      SQ_CUDA(e, e->bp_kpos.ensure((tot + 2) * 4));
      SQ_CUDA(e, e->scan_tmp.ensure(scan_tmp_words((uint32_t)tot + 1) * 4));
      launch_post_flags(sorted, tot, tbits, e->bp_newpair.as<uint32_t>(), e->bp_newkey.as<uint32_t>(), st, &e->launches);
      launch_exclusive_scan(e->bp_newpair.as<uint32_t>(), e->bp_ppos.as<uint32_t>(), (uint32_t)tot,
                            e->scan_tmp.as<uint32_t>(), st, &e->launches);
      launch_exclusive_scan(e->bp_newkey.as<uint32_t>(), e->bp_kpos.as<uint32_t>(), (uint32_t)tot,
                            e->scan_tmp.as<uint32_t>(), st, &e->launches);
      uint32_t np = 0, nkk = 0;
      SQ_CUDA(e, cudaMemcpyAsync(&np, e->bp_ppos.as<uint32_t>() + tot, 4, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaMemcpyAsync(&nkk, e->bp_kpos.as<uint32_t>() + tot, 4, cudaMemcpyDeviceToHost, st));
      SQ_CUDA(e, cudaStreamSynchronize(st));
      SQ_CUDA(e, e->bp_keys.ensure(((size_t)nkk + 1) * 4));
      SQ_CUDA(e, e->bp_off.ensure(((size_t)nkk + 1) * 8));
      SQ_CUDA(e, e->bp_post.ensure(((size_t)np + 1) * 4));
      launch_post_scatter(sorted, tot, tbits, e->bp_newpair.as<uint32_t>(), e->bp_newkey.as<uint32_t>(),
                          e->bp_ppos.as<uint32_t>(), e->bp_kpos.as<uint32_t>(), e->bp_keys.as<uint32_t>(),
                          e->bp_off.as<unsigned long long>(), e->bp_post.as<uint32_t>(), st, &e->launches);
      SQ_CUDA(e, cudaStreamSynchronize(st));
      SQ_CUDA(e, cudaGetLastError());
      e->bp_nkeys = nkk;
      e->bp_npost = np;
    }
    e->bp_kidx = (int)kidx;
    *nkeys = e->bp_nkeys;
    *npost = e->bp_npost;
    return SQ_OK;
  }
  // second call: copy the cached result out
  if (e->bp_kidx != (int)kidx) return fail(e, SQ_ERR_STATE, "sq_build_postings: sizing call for kidx %u missing", kidx);
  if (!post_off || !post_tid) return fail(e, SQ_ERR_ARG, "NULL output array");
  if (e->bp_nkeys) {
    SQ_CUDA(e, cudaMemcpyAsync(keys, e->bp_keys.p, e->bp_nkeys * 4, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(e, cudaMemcpyAsync(post_off, e->bp_off.p, (e->bp_nkeys + 1) * 8, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(e, cudaMemcpyAsync(post_tid, e->bp_post.p, e->bp_npost * 4, cudaMemcpyDeviceToHost, st));
    SQ_CUDA(e, cudaStreamSynchronize(st));
  } else {
    post_off[0] = 0;
  }
  *nkeys = e->bp_nkeys;
  *npost = e->bp_npost;
  e->bp_kidx = -1;
  return SQ_OK;
}

}  // extern "C"

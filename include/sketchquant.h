/*
 * sketchquant.h -- C ABI of the B200-native quant hot path (sketch -> seed lookup -> vote -> EM/assign).
 *
 * The reference (Codfishz/Sketch-for-RNA-seq) has no FFI: its only boundary is the process
 * (`./build/test -o quant`).  This ABI is what a host program binds instead of calling the reference's
 * in-process functions; every entry point names the reference function(s) it replaces.  Paths below
 * are relative to the reference checkout.
 *
 * Conventions: plain C types only; every function returns 0 (SQ_OK) or a negative SQ_ERR_* code and
 * never throws; the message for the last failure is available from sq_last_error().  One engine per
 * GPU; an engine is used from one host thread at a time.  Host buffers passed to sq_push_reads() may be
 * pageable or pinned (pinned gives asynchronous copies); they can be reused as soon as the call returns.
 *
 * Data layout of a read batch ("2-bit packed"): bases are coded A=0 C=1 G=2 T=3, sixteen per
 * little-endian uint32 word, base i of the batch in bits [2*(i%16), 2*(i%16)+1] of word i/16.
 * Read r occupies bases [base_off[r], base_off[r]+len[r]) of the batch; reads may start at any base.
 * Admission (ACGT-only, len >= max k; src/main.cpp:131-138) is the caller's job: every pushed read
 * counts in R (src/isoform_assignment.cpp:55).
 */
#ifndef SKETCHQUANT_H
#define SKETCHQUANT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_OK 0
#define SQ_ERR_CUDA (-1)      /* a CUDA runtime call or kernel failed */
#define SQ_ERR_ARG (-2)       /* invalid argument */
#define SQ_ERR_STATE (-3)     /* call out of order (e.g. push before load_index) */
#define SQ_ERR_CAPACITY (-4)  /* an internal buffer limit was exceeded; message says which knob */
#define SQ_ERR_NCCL (-5)      /* NCCL missing or a collective failed */
#define SQ_ERR_NO_DEVICE (-6) /* no CUDA device: there is no CPU fallback */

#define SQ_MAX_K_COUNT 8 /* number of k values per index (the reference's -k list) */

typedef struct sq_engine sq_engine;

typedef struct sq_stats {
  uint64_t reads;          /* reads pushed (R local) */
  uint64_t bases;          /* bases pushed */
  uint64_t kmers;          /* sum over reads and k of (len-k+1) */
  uint64_t sketch_hashes;  /* k-mers passing the FracMinHash threshold (before per-read dedup) */
  uint64_t pairs;          /* (read, candidate transcript) pairs kept by the vote */
  uint64_t overflow_reads; /* reads that needed the large-table path */
  uint64_t batches;
  int32_t em_iterations;   /* executed by the last sq_finish */
  int32_t peer_exchange;   /* 1: the last sq_finish exchanged the EM sums over peer memory inside the M-step kernel,
                            * 0: one GPU, or ncclAllReduce per iteration */
  /* device time of the kernels of each stage, ms, accumulated over batches (CUDA events on the engine stream;
   * only filled when profiling was enabled with sq_set_profiling) */
  float ms_sketch, ms_vote, ms_compact /* always 0: nothing is compacted any more */, ms_sort, ms_em, ms_assign;
  uint64_t launches;       /* kernels launched by this engine so far */
  /* vote-kernel work counters (for the roofline's algorithmic bytes) */
  uint64_t queries;        /* (item, k, hash) looked up in the index table */
  uint64_t hits;           /* of which found */
  uint64_t postings;       /* posting-list entries walked */
  float ms_items;          /* item split + counters (everything of the sketch stage except the sketch kernel) */
  uint32_t sketch_launches, vote_launches; /* launches of the two named kernels while profiling was on */
  uint64_t slow_reads;     /* reads voted by the general warp-per-read kernel (third tier) */
  uint64_t mid_reads;      /* short reads the bit-sliced kernel handed to the warp-per-read window kernel (second tier) */
  float ms_vote_main;      /* the first, dominant vote kernel alone (part of ms_vote) */
  float ms_lookup;         /* seed lookup kernels (one per k-index and batch) */
  uint64_t em_classes;     /* equivalence classes (distinct candidate lists) the last sq_finish ran EM on */
  uint64_t em_class_pairs; /* (class, transcript) pairs of those classes */
} sq_stats;

const char* sq_version(void);
/* number of visible CUDA devices, or SQ_ERR_NO_DEVICE */
int sq_device_count(void);
/* message of the most recent failure on this engine (engine may be NULL: last failure of sq_create) */
const char* sq_last_error(const sq_engine* e);

/* FracMinHash threshold exactly as src/sketch.cpp:25-26: (uint32_t)(UINT32_MAX * fraction). */
uint32_t sq_threshold_from_fraction(double fraction);

/* Engine bound to one GPU.  ks[0..nk) is the index's k list in index order (src/data_io.cpp:243-251;
 * duplicates allowed, each is its own k-index like src/sparse_chaining.cpp:48).  threshold is the
 * FracMinHash bound (keep h32 <= threshold, src/sketch.cpp:34), chain_fraction the 0.9 of
 * src/main.cpp:185, n_transcripts the size of the index's transcript section (T of
 * src/isoform_assignment.cpp:17-20). */
int sq_create(sq_engine** out, int device, uint32_t nk, const uint32_t* ks, uint32_t threshold,
              double chain_fraction, uint64_t n_transcripts);
void sq_destroy(sq_engine* e);

/* Run all engine work on this cudaStream_t (default: a stream the engine owns).  Lets a caller time the
 * engine with events on its own stream. */
int sq_set_stream(sq_engine* e, void* cuda_stream);
int sq_set_profiling(sq_engine* e, int enabled);
/* Tunables: "batch_bases" (max bases per internal batch), "cand_per_read" (candidate pairs per read the store makes room for ahead of the first batches),
 * "overflow_workers", "max_read_len", "em_segment", "sub_batch_reads" (reads per internal batch of
 * sq_push_reads_fixed, default 2^20): set before the first push.  "exact_classes" (any time):
 * 1 = reads are merged into one EM term only after comparing their candidate lists element by element,
 * 0 (default) = when the 128-bit fingerprints of the lists agree (~2.5 ms faster per 20 M reads).
 * "peer_exchange" (any time, every rank alike): with a communicator the ranks' EM sums can be exchanged over peer memory
 * inside the M-step kernel (when all ranks could map each other's buffers) instead of one ncclAllReduce per iteration:
 * 0 = never, 1 (default) = where that was measured faster (up to four ranks), 2 = whenever possible.
 * "vote_tier" (any time, tests): 0 = automatic, 1 = every read through the warp-per-read window kernel, 2 = every
 * read through the general warp-per-read kernel; the results do not depend on it. */
int sq_set_option(sq_engine* e, const char* name, int64_t value);

/* Replaces the TranscriptMapping for k-index kidx that load_index() fills (src/data_io.cpp:274-300,
 * include/sketch.h:23): nkeys distinct hashes, CSR offsets post_off[nkeys+1], dense transcript ids
 * post_tid[post_off[nkeys]] (< n_transcripts, each at most once per key).  Host arrays.  Builds the
 * GPU-resident index (key bitmap with ranks, list descriptors, de-duplicated posting lists).  A k-index that is never loaded behaves like a missing map
 * (src/sparse_chaining.cpp:51-53). */
int sq_load_index(sq_engine* e, uint32_t kidx, uint64_t nkeys, const uint32_t* keys, const uint64_t* post_off,
                  const uint32_t* post_tid);

/* Replaces process_fastq_single_pass()'s sketching (src/main.cpp:140-147) + sparse_chain()
 * (src/sparse_chaining.cpp:29-115) for a batch of admitted reads given in HOST memory.  Asynchronous:
 * returns once the batch is copied/enqueued.  base_off may be NULL when the reads are packed back to back,
 * read 0 at base 0 and every read starting at the next multiple of 4 bases after the previous one: then only
 * the lengths are copied to the GPU and the offsets are derived there (such a batch must fit option
 * batch_bases). */
int sq_push_reads(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, const uint32_t* base_off,
                  const uint32_t* len, uint32_t n_reads);
/* Same for reads of one length (untrimmed short-read runs), packed back to back with every read starting at the
 * next multiple of 4 bases: only the packed words travel, lengths and offsets are written on the GPU. */
int sq_push_reads_fixed(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, uint32_t read_len,
                        uint32_t n_reads);
/* Same with the batch already resident in DEVICE memory (the buffers must stay valid until sq_sync).  The packed
 * words must start on a 16-byte boundary and the allocation must be readable up to the next multiple of 4 words
 * behind n_words (any cudaMalloc'ed buffer is): the sketch kernel stages them with 16-byte bulk copies. */
int sq_push_reads_device(sq_engine* e, const uint32_t* d_packed_words, uint64_t n_words,
                         const uint32_t* d_base_off, const uint32_t* d_len, uint32_t n_reads,
                         uint64_t n_bases_hint);
/* Page-locked host memory for the buffers handed to sq_push_reads*(): copies from it are asynchronous and run at the
 * full PCIe rate (from pageable memory they are staged).  NULL when the allocation fails. */
void* sq_host_alloc(size_t bytes);
void sq_host_free(void* p);
/* Wait until every pushed batch has been voted. */
int sq_sync(sq_engine* e);
/* Forget all pushed reads (keeps the index). */
int sq_reset_reads(sq_engine* e);

/* Replaces estimate_isoform_abundance_em(.., em_iters, em_tol) (src/isoform_assignment.cpp:9-68) and
 * assign_reads_to_isoforms() (src/isoform_assignment.cpp:70-97) over everything pushed so far.
 * R_total: homologous_segments.size(); pass 0 to use this engine's read count (all-reduced when a
 * communicator is attached).  Outputs are HOST arrays of n_transcripts entries: pi (EM_Abundance),
 * numreads (NumReads), present (1 iff the transcript has a NumReads entry, i.e. a CSV row,
 * src/data_io.cpp:144-147).  iters_done may be NULL. */
int sq_finish(sq_engine* e, uint64_t R_total, int em_iters, double em_tol, double* pi, double* numreads,
              uint8_t* present, int* iters_done);

/* ---- debug taps used by the parity tests ---- */

/* createSketch_FracMinhash_direct() (src/sketch.cpp:24-39) for every read of a HOST batch and every k:
 * counts[r*nk+i] = number of k-mers of read r with h32 <= threshold for k-index i (a multiset: duplicates
 * are kept, the set is its distinct members), hashes = those values grouped by (r, i) in that order.
 * total receives the number of values; at most cap are written. */
int sq_sketch(sq_engine* e, const uint32_t* packed_words, uint64_t n_words, const uint32_t* base_off,
              const uint32_t* len, uint32_t n_reads, uint32_t* counts, uint32_t* hashes, uint64_t cap,
              uint64_t* total);
/* number of reads pushed and (read, transcript) pairs kept so far (synchronises) */
int sq_num_pairs(sq_engine* e, uint64_t* n_reads, uint64_t* n_pairs);
/* sparse_chain() result (src/sparse_chaining.cpp:111) as CSR in push order: read_off[n_reads+1], then
 * tid/score ordered per read by (score descending, tid ascending). */
int sq_get_candidates(sq_engine* e, uint64_t* read_off, uint32_t* tid, int32_t* score);

/* Replace the candidate store by caller-supplied sparse_chain() output (HOST arrays, CSR over n_reads reads):
 * lets estimate_isoform_abundance_em / assign_reads_to_isoforms be run on arbitrary homologous_segments. */
int sq_set_candidates(sq_engine* e, uint64_t n_reads, const uint64_t* read_off, const uint32_t* tid,
                      const int32_t* score);

/* ---- index construction (build_and_save_index's compute, src/main.cpp:66-85 + src/sketch.cpp:51-74) ----
 * Sequences (transcripts) as a 2-bit packed HOST batch; seq_tid[s] is the dense transcript id the
 * sequence belongs to (several sequences may share one id).  For k-index kidx returns the inverted
 * map hash -> sorted distinct transcript ids.  Two-call protocol: call with keys == NULL to get
 * *nkeys and *npost, then with arrays of that size. */
int sq_build_postings(sq_engine* e, uint32_t kidx, const uint32_t* packed_words, uint64_t n_words,
                      const uint32_t* base_off, const uint32_t* len, const uint32_t* seq_tid, uint32_t n_seqs,
                      uint64_t* nkeys, uint64_t* npost, uint32_t* keys, uint64_t* post_off, uint32_t* post_tid);

/* ---- multi-GPU (one process per GPU; the only collective is an all-reduce of T-vectors) ---- */
#define SQ_NCCL_ID_BYTES 128
int sq_nccl_unique_id(uint8_t id[SQ_NCCL_ID_BYTES]);
int sq_comm_init(sq_engine* e, int nranks, int rank, const uint8_t id[SQ_NCCL_ID_BYTES]);

int sq_get_stats(sq_engine* e, sq_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* SKETCHQUANT_H */

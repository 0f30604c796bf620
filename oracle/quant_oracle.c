/* TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
 *
 * Plain-C restatement of the reference's quant hot path on flat arrays, used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.
 * It travels to the GPU box (where /root/reference does not exist).
 *
 * Pinned against (a) the ntHash known-answer vectors of SURVEY.md Appendix B and
 * (b) the reference's own translation units compiled unmodified into
 * oracle/_ref/libref_oracle.so (tests/test_oracle_vs_reference.py, run where
 * /root/reference is present; golden fixtures produced by it are committed under
 * tests/golden/).
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- ntHash2 forward hash (third-party libnthash; call sites src/sketch.cpp:31-33) ---- */

static uint64_t seed_of(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0x3c8bfbb395c60474ULL;
    case 'C': case 'c': return 0x3193c18562a02b4cULL;
    case 'G': case 'g': return 0x20323ed082572324ULL;
    case 'T': case 't': case 'U': case 'u': return 0x295549f54be24456ULL;
    default: return 0;
  }
}

/* split rotate: bits 63..33 form a 31-bit ring, bits 32..0 a 33-bit ring */
static uint64_t srol1(uint64_t x) {
  uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}

static uint64_t sroln(uint64_t x, unsigned n) {
  while (n--) x = srol1(x);
  return x;
}

/* 64-bit forward hash of exactly k characters (no validity check) */
uint64_t orc_fwd_hash64(const char* s, uint32_t k) {
  uint64_t h = 0;
  for (uint32_t i = 0; i < k; ++i) h = srol1(h) ^ seed_of((unsigned char)s[i]);
  return h;
}

/* Low 32 bits of the forward hash of every usable window, in window order.
 * out_pos (optional) receives the window start.  Returns the number of windows
 * produced (= n-k+1 for an ACGT-only sequence; windows containing any other
 * character are skipped, ntHash2 roll()/init() behaviour).  */
uint64_t orc_hash32_windows(const char* s, uint64_t n, uint32_t k, uint32_t* out, uint64_t* out_pos) {
  if (k == 0 || n < k) return 0;
  uint64_t produced = 0, run = 0, h = 0;
  const uint64_t outrot_k = k;
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t sd = seed_of((unsigned char)s[i]);
    if (sd == 0) { run = 0; h = 0; continue; }
    if (run < k) {
      h = srol1(h) ^ sd;
      ++run;
    } else {
      h = srol1(h) ^ sd ^ sroln(seed_of((unsigned char)s[i - k]), (unsigned)outrot_k);
    }
    if (run == k) {
      if (out) out[produced] = (uint32_t)h;
      if (out_pos) out_pos[produced] = i + 1 - k;
      ++produced;
    }
  }
  return produced;
}

/* src/sketch.cpp:25-26: threshold = (uint32_t)(UINT32_MAX * fraction) with fraction a double
 * (main.cpp:43 passes (double)0.05f). */
uint32_t orc_threshold(double fraction) {
  const uint32_t H = 0xFFFFFFFFu;
  return (uint32_t)(H * fraction);
}

static int cmp_u32(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : x > y;
}

/* src/sketch.cpp:24-39 createSketch_FracMinhash_direct: the SET of h32 <= threshold.
 * Writes the members sorted ascending; returns the set size (may exceed cap; only cap are written). */
uint64_t orc_sketch(const char* s, uint64_t n, uint32_t k, uint32_t threshold, uint32_t* out, uint64_t cap) {
  if (n < k || k == 0) return 0;
  uint64_t nw = n - k + 1;
  uint32_t* all = (uint32_t*)malloc(sizeof(uint32_t) * (nw ? nw : 1));
  uint64_t m = orc_hash32_windows(s, n, k, all, NULL);
  uint64_t sel = 0;
  for (uint64_t i = 0; i < m; ++i)
    if (all[i] <= threshold) all[sel++] = all[i];
  qsort(all, sel, sizeof(uint32_t), cmp_u32);
  uint64_t u = 0;
  for (uint64_t i = 0; i < sel; ++i)
    if (i == 0 || all[i] != all[i - 1]) {
      if (u < cap) out[u] = all[i];
      ++u;
    }
  free(all);
  return u;
}

/* src/data_io.cpp:17-34 is_valid_sequence: upper-case ACGT only */
int orc_is_valid_sequence(const char* s, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    char c = s[i];
    if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return 0;
  }
  return 1;
}

/* ---- index as flat arrays: per k-index i, sorted distinct keys, CSR offsets, dense transcript ids ---- */

typedef struct {
  uint64_t nkeys;
  const uint32_t* keys; /* ascending */
  const uint64_t* off;  /* nkeys+1 */
  const uint32_t* tids;
} orc_postings;

static int64_t find_key(const orc_postings* p, uint32_t h) {
  uint64_t lo = 0, hi = p->nkeys;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    if (p->keys[mid] < h) lo = mid + 1; else hi = mid;
  }
  return (lo < p->nkeys && p->keys[lo] == h) ? (int64_t)lo : -1;
}

typedef struct { uint32_t tid; uint32_t kidx; } hit_t;
static int cmp_hit(const void* a, const void* b) {
  const hit_t* x = (const hit_t*)a; const hit_t* y = (const hit_t*)b;
  if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
  return x->kidx < y->kidx ? -1 : x->kidx > y->kidx;
}
typedef struct { uint32_t tid; int32_t score; } cand_t;
static int cmp_cand(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->score != y->score) return x->score > y->score ? -1 : 1; /* descending, sparse_chaining.cpp:108 */
  return x->tid < y->tid ? -1 : x->tid > y->tid;                 /* tie order is unspecified upstream */
}

/* src/sparse_chaining.cpp:42-112 for ONE read.
 * sk[i]/nsk[i]: the read's sketch SET for k-index i.  Returns the number of candidates and writes
 * (tid, score) ordered by (score desc, tid asc); at most cap are written. */
uint64_t orc_vote_read(uint32_t nk, const uint32_t* const* sk, const uint64_t* nsk, const orc_postings* idx,
                       double fraction, uint32_t* out_tid, int32_t* out_score, uint64_t cap) {
  uint64_t nh = 0, caph = 64;
  hit_t* hits = (hit_t*)malloc(sizeof(hit_t) * caph);
  for (uint32_t i = 0; i < nk; ++i) {
    if (idx[i].nkeys == 0) continue; /* :51-53 missing map contributes nothing */
    for (uint64_t j = 0; j < nsk[i]; ++j) {
      int64_t r = find_key(&idx[i], sk[i][j]); /* :62 */
      if (r < 0) continue;
      for (uint64_t p = idx[i].off[r]; p < idx[i].off[r + 1]; ++p) { /* :64-69 */
        if (nh == caph) { caph *= 2; hits = (hit_t*)realloc(hits, sizeof(hit_t) * caph); }
        hits[nh].tid = idx[i].tids[p]; hits[nh].kidx = i; ++nh;
      }
    }
  }
  qsort(hits, nh, sizeof(hit_t), cmp_hit);
  /* distinct transcripts with per-k counts */
  uint64_t nt = 0;
  for (uint64_t a = 0; a < nh; ++a) if (a == 0 || hits[a].tid != hits[a - 1].tid) ++nt;
  int32_t* counts = (int32_t*)calloc((nt ? nt : 1) * nk, sizeof(int32_t));
  uint32_t* tids = (uint32_t*)malloc(sizeof(uint32_t) * (nt ? nt : 1));
  int32_t* maxc = (int32_t*)calloc(nk, sizeof(int32_t));
  uint64_t t = 0;
  for (uint64_t a = 0; a < nh; ++a) {
    if (a != 0 && hits[a].tid != hits[a - 1].tid) ++t;
    tids[t] = hits[a].tid;
    counts[t * nk + hits[a].kidx]++;
  }
  for (uint64_t a = 0; a < nt; ++a) /* :76-82 */
    for (uint32_t i = 0; i < nk; ++i)
      if (counts[a * nk + i] > maxc[i]) maxc[i] = counts[a * nk + i];
  cand_t* cands = (cand_t*)malloc(sizeof(cand_t) * (nt ? nt : 1));
  uint64_t nc = 0;
  for (uint64_t a = 0; a < nt; ++a) { /* :90-105 */
    int ok = 1; int32_t score = 0;
    for (uint32_t i = 0; i < nk; ++i) {
      double thr = fraction * maxc[i]; /* :86 */
      if (counts[a * nk + i] < thr) { ok = 0; break; } /* int promoted to double, :95 */
      score += counts[a * nk + i];
    }
    if (ok) { cands[nc].tid = tids[a]; cands[nc].score = score; ++nc; }
  }
  qsort(cands, nc, sizeof(cand_t), cmp_cand);
  for (uint64_t a = 0; a < nc && a < cap; ++a) { out_tid[a] = cands[a].tid; out_score[a] = cands[a].score; }
  free(cands); free(maxc); free(tids); free(counts); free(hits);
  return nc;
}

/* src/isoform_assignment.cpp:9-68.  Candidates as CSR over `rows` reads (cand_off has rows+1 entries).
 * R is homologous_segments.size() (:55): every admitted read counts, also those with no candidate
 * (sparse_chaining.cpp:111); rows may include non-admitted reads with empty ranges, so R is passed
 * separately.  pi has T entries.  Returns the number of iterations executed. */
int orc_em(uint64_t rows, const uint64_t* cand_off, const uint32_t* cand_tid, const int32_t* cand_score,
           uint64_t R, uint64_t T, int max_iterations, double tol, double* pi) {
  double* ps = (double*)malloc(sizeof(double) * (T ? T : 1));
  for (uint64_t t = 0; t < T; ++t) pi[t] = 1.0 / (double)T; /* :17-20 */
  int it = 0;
  for (; it < max_iterations; ++it) {
    for (uint64_t t = 0; t < T; ++t) ps[t] = 0.0;
    const double epsilon = 1e-10; /* :28 */
    for (uint64_t r = 0; r < rows; ++r) { /* :30-51 */
      double den = 0.0;
      for (uint64_t j = cand_off[r]; j < cand_off[r + 1]; ++j) den += pi[cand_tid[j]] * (double)cand_score[j];
      if (den > epsilon) {
        double inv = 1.0 / den;
        for (uint64_t j = cand_off[r]; j < cand_off[r + 1]; ++j)
          ps[cand_tid[j]] += (pi[cand_tid[j]] * (double)cand_score[j]) * inv;
      }
    }
    float pseudocount = 0.01f; /* :54 */
    double total_change = 0.0;
    for (uint64_t t = 0; t < T; ++t) { /* :56-60: float/size_t division happens in float */
      double np = ps[t] + (double)(pseudocount / (float)R) + (double)pseudocount;
      total_change += fabs(np - pi[t]);
      pi[t] = np;
    }
    if (total_change < tol) { ++it; break; } /* :62-64 */
  }
  free(ps);
  return it;
}

/* src/isoform_assignment.cpp:70-97 */
void orc_assign(uint64_t rows, const uint64_t* cand_off, const uint32_t* cand_tid, const int32_t* cand_score,
                uint64_t T, const double* pi, double* numreads, uint8_t* present) {
  for (uint64_t t = 0; t < T; ++t) { numreads[t] = 0.0; present[t] = 0; }
  for (uint64_t r = 0; r < rows; ++r) {
    double tot = 0.0;
    for (uint64_t j = cand_off[r]; j < cand_off[r + 1]; ++j) tot += pi[cand_tid[j]] * cand_score[j];
    if (tot > 0.0)
      for (uint64_t j = cand_off[r]; j < cand_off[r + 1]; ++j) {
        numreads[cand_tid[j]] += (pi[cand_tid[j]] * cand_score[j]) / tot; /* :90 divides per term */
        present[cand_tid[j]] = 1;
      }
  }
}

/* ---- whole path on a batch of reads (main.cpp:165-192 minus file I/O) ----
 * reads: ASCII bases concatenated; read r is seq[roff[r] .. roff[r+1]).  Reads failing the admission rule
 * (main.cpp:131-138) are skipped; admitted[r] tells which.  Outputs (caller-allocated):
 *   cand_off[R_in+1] (over ALL input reads; skipped reads get empty ranges), cand_tid/cand_score up to cap.
 * Returns total candidates (may exceed cap -> caller retries), and *R_admitted. */
uint64_t orc_chain_batch(uint32_t nk, const uint32_t* ks, uint32_t threshold, double fraction,
                         const orc_postings* idx, uint64_t R_in, const char* seq, const uint64_t* roff,
                         uint8_t* admitted, uint64_t* R_admitted, uint64_t* cand_off, uint32_t* cand_tid,
                         int32_t* cand_score, uint64_t cap) {
  uint32_t maxk = 0;
  for (uint32_t i = 0; i < nk; ++i) if (ks[i] > maxk) maxk = ks[i];
  uint64_t total = 0, adm = 0;
  uint32_t** sk = (uint32_t**)malloc(sizeof(uint32_t*) * nk);
  uint64_t* nsk = (uint64_t*)malloc(sizeof(uint64_t) * nk);
  for (uint64_t r = 0; r < R_in; ++r) {
    const char* s = seq + roff[r];
    uint64_t n = roff[r + 1] - roff[r];
    cand_off[r] = total;
    int ok = orc_is_valid_sequence(s, n) && n >= maxk;
    if (admitted) admitted[r] = (uint8_t)ok;
    if (!ok) continue;
    ++adm;
    for (uint32_t i = 0; i < nk; ++i) {
      uint64_t nw = n - ks[i] + 1;
      sk[i] = (uint32_t*)malloc(sizeof(uint32_t) * nw);
      nsk[i] = orc_sketch(s, n, ks[i], threshold, sk[i], nw);
    }
    uint64_t room = total < cap ? cap - total : 0;
    uint64_t nc = orc_vote_read(nk, (const uint32_t* const*)sk, nsk, idx, fraction,
                                room ? cand_tid + total : NULL, room ? cand_score + total : NULL, room);
    total += nc;
    for (uint32_t i = 0; i < nk; ++i) free(sk[i]);
  }
  cand_off[R_in] = total;
  if (R_admitted) *R_admitted = adm;
  free(sk); free(nsk);
  return total;
}

/* ---- inverted map for one k from a set of sequences (main.cpp:66-85 + sketch.cpp:51-74) ----
 * seq: ASCII bases concatenated, sequence s = seq[soff[s] .. soff[s+1]) with dense transcript id s.
 * Sequences shorter than min_len (the largest k of the index, main.cpp:67-75) get no sketch.
 * Two calls: with keys == NULL returns sizes through nkeys/npost; then fills keys (ascending),
 * off[nkeys+1], tids (ascending inside a key). */
typedef struct { uint32_t h; uint32_t t; } ht_t;
static int cmp_ht(const void* a, const void* b) {
  const ht_t* x = (const ht_t*)a; const ht_t* y = (const ht_t*)b;
  if (x->h != y->h) return x->h < y->h ? -1 : 1;
  return x->t < y->t ? -1 : x->t > y->t;
}
static ht_t* g_pairs = NULL;
static uint64_t g_npairs = 0;

void orc_build_postings(uint64_t nseq, const char* seq, const uint64_t* soff, uint32_t k, uint32_t min_len,
                        uint32_t threshold, uint64_t* nkeys, uint64_t* npost, uint32_t* keys, uint64_t* off,
                        uint32_t* tids) {
  if (!keys) {
    free(g_pairs); g_pairs = NULL; g_npairs = 0;
    uint64_t cap = 1 << 20, n = 0;
    ht_t* pr = (ht_t*)malloc(sizeof(ht_t) * cap);
    uint32_t* tmp = NULL; uint64_t tmpcap = 0;
    for (uint64_t s = 0; s < nseq; ++s) {
      uint64_t len = soff[s + 1] - soff[s];
      if (len < min_len || len < k) continue;
      if (len > tmpcap) { tmpcap = len * 2; tmp = (uint32_t*)realloc(tmp, sizeof(uint32_t) * tmpcap); }
      uint64_t m = orc_sketch(seq + soff[s], len, k, threshold, tmp, len);
      if (n + m > cap) { while (n + m > cap) cap *= 2; pr = (ht_t*)realloc(pr, sizeof(ht_t) * cap); }
      for (uint64_t i = 0; i < m; ++i) { pr[n].h = tmp[i]; pr[n].t = (uint32_t)s; ++n; }
    }
    free(tmp);
    qsort(pr, n, sizeof(ht_t), cmp_ht);
    uint64_t nk = 0;
    for (uint64_t i = 0; i < n; ++i) if (i == 0 || pr[i].h != pr[i - 1].h) ++nk;
    g_pairs = pr; g_npairs = n;
    *nkeys = nk; *npost = n;
    return;
  }
  uint64_t kpos = 0;
  for (uint64_t i = 0; i < g_npairs; ++i) {
    if (i == 0 || g_pairs[i].h != g_pairs[i - 1].h) { keys[kpos] = g_pairs[i].h; off[kpos] = i; ++kpos; }
    tids[i] = g_pairs[i].t;
  }
  off[kpos] = g_npairs;
  *nkeys = kpos; *npost = g_npairs;
  free(g_pairs); g_pairs = NULL; g_npairs = 0;
}

// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// C-callable harness around the UNMODIFIED reference translation units, which
// are compiled from /root/reference/src where they lie (see oracle/Makefile;
// no reference source is copied into this repository).  It drives exactly the
// functions quantification() calls (/root/reference/src/main.cpp:165-197):
//
//   load_index                       src/data_io.cpp:233
//   process_fastq_single_pass        src/main.cpp:107   (main.cpp is built with -Dmain=ref_main)
//   createSketch_FracMinhash_direct  src/sketch.cpp:24
//   sparse_chain                     src/sparse_chaining.cpp:29
//   estimate_isoform_abundance_em    src/isoform_assignment.cpp:9
//   assign_reads_to_isoforms         src/isoform_assignment.cpp:70
//
// and exposes the in-memory results (sets, candidate lists, doubles) as flat
// arrays, because the CSV only carries 6 significant digits.
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "data_io.h"
#include "isoform_assignment.h"
#include "sketch.h"
#include "sparse_chaining.h"

// defined in /root/reference/src/main.cpp:107 (no header declares it)
std::unordered_map<std::string, MultiKmerSketch> process_fastq_single_pass(
    const std::string& fastq_file, const std::vector<unsigned>& effective_kmer_lengths, double sketch_size);

namespace {

struct RefQuant {
  std::vector<unsigned> ks;
  std::unordered_map<unsigned, TranscriptMapping> index;
  std::unordered_map<std::string, Transcript> transcripts;
  std::vector<std::string> tnames;                       // dense index -> id
  std::unordered_map<std::string, uint32_t> tindex;      // id -> dense index
  std::unordered_map<std::string, MultiKmerSketch> read_sketches;
  std::unordered_map<std::string, std::vector<std::pair<std::string, int>>> segments;
  std::unordered_map<std::string, double> pi, counts;
  double t_sketch = 0, t_chain = 0, t_em = 0, t_assign = 0;
};

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void* refq_create(int nk, const unsigned* ks) {
  auto* q = new RefQuant();
  q->ks.assign(ks, ks + nk);
  return q;
}

void refq_destroy(void* h) { delete static_cast<RefQuant*>(h); }

// Transcript section of the index: ids only matter for quant (sequence is unused there).
void refq_set_transcripts(void* h, uint64_t n, const char* const* names) {
  auto* q = static_cast<RefQuant*>(h);
  q->tnames.clear();
  q->tindex.clear();
  q->transcripts.clear();
  for (uint64_t i = 0; i < n; ++i) {
    std::string id(names[i]);
    q->tnames.push_back(id);
    q->tindex[id] = static_cast<uint32_t>(i);
    q->transcripts[id] = Transcript{id, std::string(), 0};
  }
}

// One k's hash -> posting list map, given as CSR over dense transcript indices.
void refq_set_postings(void* h, unsigned k, uint64_t nkeys, const uint32_t* keys, const uint64_t* off,
                       const uint32_t* tids) {
  auto* q = static_cast<RefQuant*>(h);
  TranscriptMapping& m = q->index[k];
  m.clear();
  m.reserve(nkeys);
  for (uint64_t i = 0; i < nkeys; ++i) {
    auto& vec = m[keys[i]];
    for (uint64_t j = off[i]; j < off[i + 1]; ++j) vec.emplace_back(q->tnames[tids[j]], nullptr);
  }
}

// Reference loader on a reference-format index file (overwrites k list like main.cpp:174).
int refq_load_index_file(void* h, const char* path) {
  auto* q = static_cast<RefQuant*>(h);
  q->index.clear();
  q->transcripts.clear();
  load_index(path, q->ks, q->index, q->transcripts);
  q->tnames.clear();
  q->tindex.clear();
  for (const auto& kv : q->transcripts) q->tnames.push_back(kv.first);
  std::sort(q->tnames.begin(), q->tnames.end());
  for (size_t i = 0; i < q->tnames.size(); ++i) q->tindex[q->tnames[i]] = static_cast<uint32_t>(i);
  return static_cast<int>(q->ks.size());
}

int refq_get_ks(void* h, unsigned* out, int cap) {
  auto* q = static_cast<RefQuant*>(h);
  int n = static_cast<int>(q->ks.size());
  for (int i = 0; i < n && i < cap; ++i) out[i] = q->ks[i];
  return n;
}
uint64_t refq_num_transcripts(void* h) { return static_cast<RefQuant*>(h)->tnames.size(); }
const char* refq_transcript_name(void* h, uint64_t i) { return static_cast<RefQuant*>(h)->tnames[i].c_str(); }

// number of keys of one k's map, and a dump of it as CSR over dense transcript indices
uint64_t refq_postings_size(void* h, unsigned k, uint64_t* n_postings) {
  auto* q = static_cast<RefQuant*>(h);
  auto it = q->index.find(k);
  if (it == q->index.end()) { *n_postings = 0; return 0; }
  uint64_t np = 0;
  for (const auto& kv : it->second) np += kv.second.size();
  *n_postings = np;
  return it->second.size();
}
void refq_get_postings(void* h, unsigned k, uint32_t* keys, uint64_t* off, uint32_t* tids) {
  auto* q = static_cast<RefQuant*>(h);
  auto it = q->index.find(k);
  if (it == q->index.end()) return;
  std::vector<uint32_t> ks;
  for (const auto& kv : it->second) ks.push_back(kv.first);
  std::sort(ks.begin(), ks.end());
  uint64_t o = 0;
  for (size_t i = 0; i < ks.size(); ++i) {
    keys[i] = ks[i];
    off[i] = o;
    std::vector<uint32_t> v;
    for (const auto& pr : it->second.at(ks[i])) v.push_back(q->tindex.at(pr.first));
    std::sort(v.begin(), v.end());
    for (uint32_t t : v) tids[o++] = t;
  }
  off[ks.size()] = o;
}

// createSketch_FracMinhash_direct on one sequence (sketch.cpp:24); returns the set size,
// writes up to cap sorted members.
uint64_t refq_sketch(const char* seq, uint64_t len, int k, double fraction, uint32_t* out, uint64_t cap) {
  std::string s(seq, len);
  auto set = createSketch_FracMinhash_direct(s, k, fraction);
  std::vector<uint32_t> v(set.begin(), set.end());
  std::sort(v.begin(), v.end());
  for (uint64_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return v.size();
}

// process_fastq_single_pass (main.cpp:107) on a FASTQ file; returns number of admitted reads.
uint64_t refq_fastq(void* h, const char* path, double sketch_size) {
  auto* q = static_cast<RefQuant*>(h);
  double t0 = now_s();
  q->read_sketches = process_fastq_single_pass(path, q->ks, sketch_size);
  q->t_sketch = now_s() - t0;
  return q->read_sketches.size();
}

// Same admission + sketching as main.cpp:131-147 for a read handed over in memory.
int refq_add_read(void* h, const char* id, const char* seq, uint64_t len, double sketch_size) {
  auto* q = static_cast<RefQuant*>(h);
  std::string s(seq, len);
  if (!is_valid_sequence(s)) return 0;
  unsigned max_k = *std::max_element(q->ks.begin(), q->ks.end());
  if (s.size() < max_k) return 0;
  MultiKmerSketch mks;
  for (unsigned k : q->ks) mks.sketches[k] = createSketch_FracMinhash_direct(s, k, sketch_size);
  q->read_sketches[id] = std::move(mks);
  return 1;
}

uint64_t refq_num_reads(void* h) { return static_cast<RefQuant*>(h)->read_sketches.size(); }

void refq_chain(void* h, double fraction) {
  auto* q = static_cast<RefQuant*>(h);
  double t0 = now_s();
  q->segments = sparse_chain(q->read_sketches, q->index, q->transcripts, q->ks, fraction);
  q->t_chain = now_s() - t0;
}

void refq_em(void* h, int max_iterations, double tol) {
  auto* q = static_cast<RefQuant*>(h);
  double t0 = now_s();
  q->pi = estimate_isoform_abundance_em(q->segments, q->transcripts, max_iterations, tol);
  q->t_em = now_s() - t0;
}

void refq_assign(void* h) {
  auto* q = static_cast<RefQuant*>(h);
  double t0 = now_s();
  q->counts = assign_reads_to_isoforms(q->segments, q->pi, q->transcripts);
  q->t_assign = now_s() - t0;
}

void refq_times(void* h, double* out4) {
  auto* q = static_cast<RefQuant*>(h);
  out4[0] = q->t_sketch; out4[1] = q->t_chain; out4[2] = q->t_em; out4[3] = q->t_assign;
}

// sketch set of one admitted read for k (sorted); returns size or -1 if the read id is unknown
int64_t refq_read_sketch(void* h, const char* id, unsigned k, uint32_t* out, uint64_t cap) {
  auto* q = static_cast<RefQuant*>(h);
  auto it = q->read_sketches.find(id);
  if (it == q->read_sketches.end()) return -1;
  auto st = it->second.sketches.find(k);
  if (st == it->second.sketches.end()) return 0;
  std::vector<uint32_t> v(st->second.begin(), st->second.end());
  std::sort(v.begin(), v.end());
  for (uint64_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
  return static_cast<int64_t>(v.size());
}

// candidate list of one read as (dense transcript index, score), sorted by (score desc, index asc);
// returns length or -1 if the read id is unknown
int64_t refq_read_candidates(void* h, const char* id, uint32_t* tid, int32_t* score, uint64_t cap) {
  auto* q = static_cast<RefQuant*>(h);
  auto it = q->segments.find(id);
  if (it == q->segments.end()) return -1;
  std::vector<std::pair<int32_t, uint32_t>> v;
  for (const auto& pr : it->second) v.emplace_back(-pr.second, q->tindex.at(pr.first));
  std::sort(v.begin(), v.end());
  for (uint64_t i = 0; i < v.size() && i < cap; ++i) { tid[i] = v[i].second; score[i] = -v[i].first; }
  return static_cast<int64_t>(v.size());
}

// candidate lists of the reads named <prefix><i>, i = 0..n-1, as one CSR (each list ordered like
// refq_read_candidates); a read without an entry gets an empty list.  Returns the total number of
// pairs; tid/score receive at most cap of them.
uint64_t refq_candidates_csr(void* h, const char* prefix, uint64_t n, uint64_t* off, uint32_t* tid, int32_t* score,
                             uint64_t cap) {
  auto* q = static_cast<RefQuant*>(h);
  uint64_t tot = 0;
  std::vector<std::pair<int32_t, uint32_t>> v;
  for (uint64_t i = 0; i < n; ++i) {
    off[i] = tot;
    auto it = q->segments.find(std::string(prefix) + std::to_string(i));
    if (it == q->segments.end()) continue;
    v.clear();
    for (const auto& pr : it->second) v.emplace_back(-pr.second, q->tindex.at(pr.first));
    std::sort(v.begin(), v.end());
    for (const auto& e : v) {
      if (tot < cap) { tid[tot] = e.second; score[tot] = -e.first; }
      ++tot;
    }
  }
  off[n] = tot;
  return tot;
}

void refq_get_pi(void* h, double* out) {
  auto* q = static_cast<RefQuant*>(h);
  for (size_t i = 0; i < q->tnames.size(); ++i) {
    auto it = q->pi.find(q->tnames[i]);
    out[i] = it == q->pi.end() ? 0.0 : it->second;
  }
}

void refq_get_counts(void* h, double* out, uint8_t* present) {
  auto* q = static_cast<RefQuant*>(h);
  for (size_t i = 0; i < q->tnames.size(); ++i) {
    auto it = q->counts.find(q->tnames[i]);
    present[i] = it != q->counts.end();
    out[i] = present[i] ? it->second : 0.0;
  }
}

// output_to_csv (data_io.cpp:133) with the harness state
void refq_write_csv(void* h, const char* path) {
  auto* q = static_cast<RefQuant*>(h);
  output_to_csv(path, q->counts, q->pi, q->transcripts);
}

}  // extern "C"

// Drop-in command line for the reference's `./build/test` (src/main.cpp:212-276): same options, positional
// arguments, stdout lines, exit codes, index file format and CSV; the quant hot path (and the sketching of the
// index build) runs on the GPU through the C ABI of include/sketchquant.h.  There is no CPU fallback: without a
// CUDA device the program stops with an error.
//
//   test [-k 21,25,31] -o index <reference.fasta> <index_out>
//   test [-k ...]      -o quant <index_file> <reads.fastq> <output.csv>
//
// Extras that default to the reference's behaviour: --gpus N (shard the reads over N GPUs of this box, one host
// thread and one engine per GPU, NCCL all-reduce of the per-transcript vectors), --threads N (host parsing
// threads), env SQ_SKETCH_SIZE / SQ_CHAIN_FRACTION / SQ_EM_ITERS / SQ_EM_TOL (the constants hard-coded at
// src/main.cpp:43,185,188), --report FILE (JSON timing report), --index-cache / SQ_INDEX_CACHE=1 (keep a
// parsed copy of the index next to it as <index>.sqidx, used only while the index file is unchanged).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sketchquant.h"
#include "fastx.hpp"
#include "index_file.hpp"

using namespace sqhost;

namespace {

struct Options {
  float sketch_size = 0.05f;  // const float sketch_size = 0.05f (main.cpp:43)
  double chain_fraction = 0.9;
  int em_iters = 20;
  double em_tol = 0.01;
  int gpus = 1;
  int threads = 0;
  bool index_cache = false;
  std::string report;
};

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void print_help(const std::string& program_name) {  // main.cpp:24-40, verbatim (including the "default: 81")
  std::cout << "Usage: " << program_name << " [OPTIONS] <mode> [arguments]" << std::endl;
  std::cout << "Modes:" << std::endl;
  std::cout << "  index   Build index from reference genome" << std::endl;
  std::cout << "  quant   Quantify using pre-built index and reads" << std::endl;
  std::cout << std::endl;
  std::cout << "Options:" << std::endl;
  std::cout << "  -h, --help              Show this help message and exit" << std::endl;
  std::cout << "  -k, --kmer-length SIZE  Comma separated list of k-mer lengths (default: 81)" << std::endl;
  std::cout << "  -o, --mode MODE         Mode: index or quant (default: quant)" << std::endl;
  std::cout << std::endl;
  std::cout << "Index mode usage:" << std::endl;
  std::cout << "  " << program_name << " index <reference_genome.fasta> <index_output>" << std::endl;
  std::cout << std::endl;
  std::cout << "Quant mode usage:" << std::endl;
  std::cout << "  " << program_name << " quant <index_file> <reads.fastq> <output>" << std::endl;
}

[[noreturn]] void die(sq_engine* e, int rc, const char* what) {  // may be called from a worker thread
  std::cerr << "sketchquant: " << what << " failed (" << rc << "): " << sq_last_error(e) << std::endl;
  std::cout.flush();
  _exit(2);
}
#define SQ(e, call)                    \
  do {                                 \
    int _rc = (call);                  \
    if (_rc != SQ_OK) die(e, _rc, #call); \
  } while (0)

// split a sequence at characters ntHash cannot hash (anything but ACGTU/acgtu): windows containing them are
// skipped by ntHash2, which is the same as hashing every maximal clean run on its own
void clean_runs(const std::string& s, uint32_t min_len, std::vector<std::pair<uint32_t, uint32_t>>* runs) {
  auto ok = [](char c) {
    switch (c) {
      case 'A': case 'C': case 'G': case 'T': case 'U': case 'a': case 'c': case 'g': case 't': case 'u': return true;
      default: return false;
    }
  };
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && !ok(s[i])) ++i;
    size_t j = i;
    while (j < s.size() && ok(s[j])) ++j;
    if (j - i >= min_len) runs->emplace_back((uint32_t)i, (uint32_t)(j - i));
    i = j;
  }
}

// build_and_save_index (main.cpp:56-92)
void build_and_save_index(const std::string& fasta, const std::string& out_path, std::vector<unsigned>& ks,
                          const Options& opt) {
  const double t0 = now();
  std::vector<FastaRecord> recs = load_fasta(fasta);
  if (ks.empty()) throw std::runtime_error("no k-mer length given");
  const uint32_t kmax = *std::max_element(ks.begin(), ks.end());
  const uint32_t kmin = *std::min_element(ks.begin(), ks.end());
  std::vector<std::string> names, seqs;
  names.reserve(recs.size());
  seqs.reserve(recs.size());
  // sequences to sketch: transcripts at least as long as every k (main.cpp:66-75), cut into clean runs
  std::vector<const char*> ptr;
  std::vector<uint32_t> len, tid;
  std::vector<std::pair<uint32_t, uint32_t>> runs;
  for (size_t i = 0; i < recs.size(); ++i) {
    names.push_back(recs[i].id);
    seqs.push_back(std::move(recs[i].sequence));
  }
  for (size_t i = 0; i < seqs.size(); ++i) {
    if (seqs[i].size() < kmax) continue;
    runs.clear();
    clean_runs(seqs[i], kmin, &runs);
    for (auto& r : runs) {
      ptr.push_back(seqs[i].data() + r.first);
      len.push_back(r.second);
      tid.push_back((uint32_t)i);
    }
  }
  std::unordered_map<uint32_t, Postings> maps;
  if (!names.empty()) {
    std::vector<uint32_t> k32(ks.begin(), ks.end());
    sq_engine* e = nullptr;
    int rc = sq_create(&e, 0, (uint32_t)k32.size(), k32.data(), sq_threshold_from_fraction((double)opt.sketch_size),
                       opt.chain_fraction, names.size());
    if (rc != SQ_OK) die(nullptr, rc, "sq_create");
    PackedBatch pb;
    pack_sequences(ptr.data(), len.data(), ptr.size(), opt.threads, &pb);
    for (size_t ki = 0; ki < k32.size(); ++ki) {
      // a clean run shorter than this k has no window for it; the kernel skips it (L < k)
      uint64_t nkeys = 0, npost = 0;
      SQ(e, sq_build_postings(e, (uint32_t)ki, pb.words.data(), pb.words.size(), pb.base_off.data(), pb.len.data(),
                              tid.data(), (uint32_t)ptr.size(), &nkeys, &npost, nullptr, nullptr, nullptr));
      Postings& P = maps[k32[ki]];  // duplicate k values share one map, like kmer_to_transcripts[k] upstream
      P.keys.resize(nkeys);
      P.off.resize(nkeys + 1);
      P.tid.resize(npost);
      SQ(e, sq_build_postings(e, (uint32_t)ki, pb.words.data(), pb.words.size(), pb.base_off.data(), pb.len.data(),
                              tid.data(), (uint32_t)ptr.size(), &nkeys, &npost, P.keys.data(), P.off.data(),
                              P.tid.data()));
    }
    sq_destroy(e);
  }
  std::cout << "Index built in " << (now() - t0) << " seconds." << std::endl;  // main.cpp:88
  std::vector<uint32_t> k32(ks.begin(), ks.end());
  write_index(out_path, k32, names, seqs, maps);
}

void output_to_csv(const std::string& path, const std::vector<std::string>& names, const std::vector<double>& numreads,
                   const std::vector<double>& pi, const std::vector<uint8_t>& present) {  // data_io.cpp:133-152
  std::ofstream out(path);
  if (!out.is_open()) throw std::runtime_error("Could not open file for writing: " + path);
  out << "Name,NumReads,EM_Abundance\n";
  for (size_t i = 0; i < names.size(); ++i)
    if (present[i]) out << names[i] << "," << numreads[i] << "," << pi[i] << "\n";
  out.close();
}

// small blocking queue for the ingest pipeline
template <class T>
class Channel {
 public:
  explicit Channel(size_t cap) : cap_(cap) {}
  bool push(T v) {  // false when the channel was aborted
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return q_.size() < cap_ || abort_; });
    if (abort_) return false;
    q_.push_back(std::move(v));
    cv_.notify_all();
    return true;
  }
  bool pop(T* out) {  // false when closed and drained, or aborted
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || closed_ || abort_; });
    if (abort_ || q_.empty()) return false;
    *out = std::move(q_.front());
    q_.pop_front();
    cv_.notify_all();
    return true;
  }
  void close() { std::lock_guard<std::mutex> lk(mu_); closed_ = true; cv_.notify_all(); }
  void abort() { std::lock_guard<std::mutex> lk(mu_); abort_ = true; cv_.notify_all(); }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
  size_t cap_;
  bool closed_ = false, abort_ = false;
};

// a chunk of admitted, packed reads in page-locked memory (sq_host_alloc)
struct PinnedBatch {
  void* mem = nullptr;
  size_t cap = 0;
  PackedView v;
  void ensure(uint64_t seq_bytes, size_t n_recs) {
    const size_t words = (size_t)((seq_bytes + 4 * n_recs) / 16 + 16);
    const size_t need = words * 4 + n_recs * 8 + 256;
    if (need > cap) {
      if (mem) sq_host_free(mem);
      cap = need + need / 4;
      mem = sq_host_alloc(cap);
      if (!mem) throw std::runtime_error("cannot allocate page-locked host memory");
    }
    v = PackedView();
    v.words = static_cast<uint32_t*>(mem);
    v.base_off = v.words + words;
    v.len = v.base_off + n_recs;
  }
  void release() { if (mem) sq_host_free(mem); mem = nullptr; cap = 0; }
};

// quantification (main.cpp:165-197)
void quantification(const std::string& index_path, const std::string& reads_path, const std::string& output_path,
                    std::vector<unsigned>& kmer_lengths, const Options& opt) {
  const double t_start = now();
  // The CUDA driver and the first context take longer to come up than a cached index takes to read: start them now,
  // beside the index file, instead of when the first engine is created (a page-locked allocation is the lightest
  // call of the ABI that needs a context).  Without a device this does nothing; the error comes from sq_create.
  std::thread cuda_warmup([] {
    if (sq_device_count() > 0) {
      void* p = sq_host_alloc(4096);
      if (p) sq_host_free(p);
    }
  });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } warmup_joiner{cuda_warmup};
  IndexData idx;
  const bool have_index = read_index(index_path, &idx, false, opt.index_cache);
  if (have_index) kmer_lengths.assign(idx.ks.begin(), idx.ks.end());  // load_index overwrites the -k list (main.cpp:174)
  std::cout << "Loading index completed" << std::endl;
  const double t_index = now();
  if (kmer_lengths.empty()) throw std::runtime_error("no k-mer length available");
  const uint32_t kmax = *std::max_element(kmer_lengths.begin(), kmer_lengths.end());
  const bool trace = getenv("SQ_TRACE") != nullptr;

  if (trace) fprintf(stderr, "[sq trace] index file parsed            %.3f s\n", now() - t_start);
  int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const size_t T = idx.names.size();
  std::vector<double> pi(T, 0.0), numreads(T, 0.0);
  std::vector<uint8_t> present(T, 0);
  double t_reads = 0, t_chain = 0, t_em = 0;
  sq_stats stats;
  memset(&stats, 0, sizeof(stats));
  uint64_t n_seen = 0, R = 0;
  bool exact_path = false;

  if (T == 0) {
    // unreadable/empty index: upstream carries on with empty maps and writes a header-only CSV
    FastqFile fq(reads_path);
    R = fq.admitted_records(kmax, threads, &n_seen).size();
    std::cout << "Loading read completed" << std::endl;
    std::cout << "Sparse chaining completed" << std::endl;
    std::cout << "EM estimation completed" << std::endl;
    std::cout << "Read assignment completed" << std::endl;
    t_reads = t_chain = t_em = now();
  } else {
    const int G = std::max(1, opt.gpus);
    const int ndev = sq_device_count();
    if (ndev < G) {
      std::cerr << "sketchquant: " << G << " GPU(s) requested, " << (ndev < 0 ? 0 : ndev)
                << " visible; this program has no CPU fallback" << std::endl;
      std::exit(2);
    }
    std::vector<uint32_t> k32(kmer_lengths.begin(), kmer_lengths.end());
    std::vector<sq_engine*> eng(G, nullptr);
    uint8_t uid[SQ_NCCL_ID_BYTES];
    if (G > 1) SQ(nullptr, sq_nccl_unique_id(uid));
    FastqFile fq(reads_path);  // an unopenable file throws like upstream

    // ---- pipeline: the file is cut into segments at line starts (line counts taken in parallel, fastx.hpp); every
    // worker scans segments into chunks of records (by reads AND by bases: a chunk of long reads stays far below
    // the engine's batch limits), admits + packs each chunk into a page-locked buffer; one pusher per GPU creates
    // its engine and index replica, then hands ready chunks to sq_push_reads while scanning and packing go on.
    // A file that is not a regular four-line FASTQ is redone through the sequential scan, like one with duplicate ids.
    const size_t chunk_reads = 1u << 18;
    const uint64_t chunk_bases = 1ull << 26;
    const int W = std::max(1, threads - 1);
    Channel<PinnedBatch*> ready((size_t)2 * G + 2);
    Channel<PinnedBatch*> pool((size_t)W + 2 * G + 4);
    std::vector<PinnedBatch> buffers((size_t)W + 2 * G + 2);
    for (auto& bf : buffers) pool.push(&bf);
    uint64_t fq_lines = 0;
    const std::vector<FastqSegment> segs = split_fastq(fq, (size_t)32 << 20, W, &fq_lines);
    IdSet ids(fq_lines / 4 + 1024);  // sized from the line count: a regular file has exactly lines / 4 records
    std::atomic<bool> ids_unique{true}, regular{true};
    std::atomic<uint64_t> admitted{0}, seen{0};
    std::atomic<size_t> next_seg{0};
    std::mutex err_mu;
    std::string first_error;
    auto fail_all = [&](const std::string& what) {
      { std::lock_guard<std::mutex> lk(err_mu); if (first_error.empty()) first_error = what; }
      ready.abort(); pool.abort();
    };
    std::atomic<int> workers_left{W};
    std::vector<std::thread> workers;
    for (int w = 0; w < W; ++w)
      workers.emplace_back([&] {
        try {
          RawChunk c;
          bool go = true;
          for (size_t si; go && regular && (si = next_seg++) < segs.size();) {
            FastqSegmentScanner sc(fq, segs[si]);
            while (go && sc.next(chunk_reads, chunk_bases, &c)) {
              if (c.recs.empty()) continue;
              PinnedBatch* bf = nullptr;
              if (!pool.pop(&bf)) { go = false; break; }
              bf->ensure(c.seq_bytes, c.recs.size());
              if (!admit_and_pack(fq.data(), c, kmax, &ids, &bf->v)) ids_unique = false;
              admitted += bf->v.n_reads;
              if (bf->v.n_reads == 0) { pool.push(bf); continue; }
              if (!ready.push(bf)) go = false;
            }
            seen += sc.records_seen();
            if (!sc.regular()) regular = false;
          }
        } catch (const std::exception& ex) { fail_all(ex.what()); }
        if (--workers_left == 0) ready.close();
      });
    std::vector<std::thread> pushers;
    for (int g = 0; g < G; ++g)
      pushers.emplace_back([&, g] {
        sq_engine* e = nullptr;
        int rc = sq_create(&e, g, (uint32_t)k32.size(), k32.data(), sq_threshold_from_fraction((double)opt.sketch_size),
                           opt.chain_fraction, T);
        if (rc != SQ_OK) die(nullptr, rc, "sq_create");
        eng[g] = e;
        if (opt.report.size()) sq_set_profiling(e, 1);
        for (size_t ki = 0; ki < k32.size(); ++ki) {
          auto it = idx.maps.find(k32[ki]);
          if (it == idx.maps.end()) continue;  // no map for this k: contributes nothing (sparse_chaining.cpp:51-53)
          const Postings& P = it->second;
          SQ(e, sq_load_index(e, (uint32_t)ki, P.keys.size(), P.keys.data(), P.off.data(), P.tid.data()));
        }
        if (G > 1) SQ(e, sq_comm_init(e, G, g, uid));
        if (trace) fprintf(stderr, "[sq trace] gpu %d engine+index ready      %.3f s since start\n", g, now() - t_start);
        PinnedBatch* bf = nullptr;
        while (ready.pop(&bf)) {
          SQ(e, sq_push_reads(e, bf->v.words, bf->v.n_words, bf->v.base_off, bf->v.len, bf->v.n_reads));
          pool.push(bf);
        }
      });
    for (auto& w : workers) w.join();
    for (auto& p : pushers) p.join();
    if (!first_error.empty()) throw std::runtime_error(first_error);
    n_seen = seen;
    R = admitted;
    if (trace) fprintf(stderr, "[sq trace] streamed %llu reads              %.3f s since start\n", (unsigned long long)R, now() - t_start);

    auto on_all_gpus = [&](auto&& fn) {
      std::vector<std::thread> th;
      for (int g = 0; g < G; ++g) th.emplace_back([&, g] { fn(g); });
      for (auto& t : th) t.join();
    };
    if (!ids_unique || !regular) {
      // some read id came twice (or the id set overflowed): later duplicates replace earlier ones
      // (read_sketches[read.id] = ..., main.cpp:147), a whole-file rule -- or the file is not a regular four-line
      // FASTQ, so the segments could not be scanned on their own: redo the reads through the exact sequential scan
      exact_path = true;
      if (trace) fprintf(stderr, "[sq trace] %s: exact path\n", regular ? "duplicate read ids" : "irregular FASTQ");
      std::vector<FastqFile::Rec> recs = fq.admitted_records(kmax, threads, &n_seen);
      R = recs.size();
      on_all_gpus([&](int g) {
        sq_engine* e = eng[g];
        SQ(e, sq_reset_reads(e));
        const uint64_t per = (R + G - 1) / G, lo = std::min<uint64_t>(R, g * per), hi = std::min<uint64_t>(R, lo + per);
        PackedBatch pb;
        std::vector<const char*> ptr;
        std::vector<uint32_t> len;
        for (uint64_t b = lo; b < hi;) {
          ptr.clear();
          len.clear();
          uint64_t bases = 0;
          while (b < hi && ptr.size() < chunk_reads && bases < chunk_bases) {  // chunks by reads and by bases
            ptr.push_back(fq.data() + recs[b].seq_off);
            len.push_back(recs[b].seq_len);
            bases += recs[b].seq_len;
            ++b;
          }
          pack_sequences(ptr.data(), len.data(), ptr.size(), std::max(1, threads / G), &pb);
          SQ(e, sq_push_reads(e, pb.words.data(), pb.words.size(), pb.base_off.data(), pb.len.data(), (uint32_t)ptr.size()));
        }
      });
    }
    for (auto& bf : buffers) bf.release();
    std::cout << "Loading read completed" << std::endl;  // upstream's read stage includes the sketching (main.cpp:182)
    t_reads = now();
    on_all_gpus([&](int g) { SQ(eng[g], sq_sync(eng[g])); });
    std::cout << "Sparse chaining completed" << std::endl;
    t_chain = now();
    std::vector<std::vector<double>> pis(G), nrs(G);
    std::vector<std::vector<uint8_t>> prs(G);
    std::vector<sq_stats> sts(G);
    on_all_gpus([&](int g) {
      pis[g].resize(T);
      nrs[g].resize(T);
      prs[g].resize(T);
      int iters = 0;
      SQ(eng[g], sq_finish(eng[g], G > 1 ? 0 : R, opt.em_iters, opt.em_tol, pis[g].data(), nrs[g].data(), prs[g].data(), &iters));
      sq_get_stats(eng[g], &sts[g]);
    });
    std::cout << "EM estimation completed" << std::endl;
    std::cout << "Read assignment completed" << std::endl;
    t_em = now();
    pi = pis[0];
    numreads = nrs[0];
    present = prs[0];
    stats = sts[0];
    for (int g = 1; g < G; ++g) {
      stats.reads += sts[g].reads; stats.bases += sts[g].bases; stats.sketch_hashes += sts[g].sketch_hashes;
      stats.pairs += sts[g].pairs; stats.launches += sts[g].launches;
    }
    for (auto* e : eng) sq_destroy(e);
  }
  output_to_csv(output_path, idx.names, numreads, pi, present);
  std::cout << "Output written to " << output_path << std::endl;
  if (opt.report.size()) {
    std::ofstream r(opt.report);
    const double t_end = now();
    r << "{\"records_seen\": " << n_seen << ", \"reads_admitted\": " << R << ", \"transcripts\": " << T
      << ", \"gpus\": " << opt.gpus << ", \"host_threads\": " << threads << ", \"duplicate_id_path\": " << (exact_path ? "true" : "false")
      << ", \"s_load_index\": " << (t_index - t_start) << ", \"s_parse_pack_push\": " << (t_reads - t_index)
      << ", \"s_vote_drain\": " << (t_chain - t_reads) << ", \"s_em_assign\": " << (t_em - t_chain)
      << ", \"s_total\": " << (t_end - t_start) << ", \"reads_per_s_quant\": " << (R / std::max(1e-9, t_em - t_index))
      << ", \"pairs\": " << stats.pairs << ", \"sketch_hashes\": " << stats.sketch_hashes
      << ", \"kernel_launches\": " << stats.launches << ", \"ms_sketch\": " << stats.ms_sketch
      << ", \"ms_vote\": " << stats.ms_vote << ", \"ms_sort\": " << stats.ms_sort << ", \"ms_em\": " << stats.ms_em
      << ", \"ms_assign\": " << stats.ms_assign << "}\n";
  }
}

}  // namespace

int main(int argc, char* argv[]) {
  std::string mode = "quant";
  std::vector<unsigned> kmer_lengths = {31};  // main.cpp:215
  Options opt;
  if (const char* v = getenv("SQ_SKETCH_SIZE")) opt.sketch_size = strtof(v, nullptr);
  if (const char* v = getenv("SQ_CHAIN_FRACTION")) opt.chain_fraction = atof(v);
  if (const char* v = getenv("SQ_EM_ITERS")) opt.em_iters = atoi(v);
  if (const char* v = getenv("SQ_EM_TOL")) opt.em_tol = atof(v);
  if (const char* v = getenv("SQ_INDEX_CACHE")) opt.index_cache = atoi(v) != 0;

  static struct option long_options[] = {{"help", no_argument, 0, 'h'},
                                         {"kmer-length", required_argument, 0, 'k'},
                                         {"mode", required_argument, 0, 'o'},
                                         {"gpus", required_argument, 0, 1000},
                                         {"threads", required_argument, 0, 1001},
                                         {"report", required_argument, 0, 1002},
                                         {"index-cache", no_argument, 0, 1003},
                                         {0, 0, 0, 0}};
  int opt_c, option_index = 0;
  while ((opt_c = getopt_long(argc, argv, "hk:o:", long_options, &option_index)) != -1) {
    switch (opt_c) {
      case 'h':
        print_help(argv[0]);
        return 0;
      case 'k': {
        std::string s(optarg);
        kmer_lengths.clear();
        std::istringstream iss(s);
        std::string token;
        while (std::getline(iss, token, ','))
          if (!token.empty()) kmer_lengths.push_back(std::stoi(token));  // main.cpp:231-241
        break;
      }
      case 'o':
        mode = std::string(optarg);
        break;
      case 1000: opt.gpus = atoi(optarg); break;
      case 1001: opt.threads = atoi(optarg); break;
      case 1002: opt.report = optarg; break;
      case 1003: opt.index_cache = true; break;
      default:
        print_help(argv[0]);
        return 1;
    }
  }
  if (mode == "index") {
    if (optind + 2 > argc) {
      std::cerr << "Usage: " << argv[0] << " index <reference_genome.fasta> <index_output>" << std::endl;
      return 1;
    }
    build_and_save_index(argv[optind], argv[optind + 1], kmer_lengths, opt);
  } else if (mode == "quant") {
    if (optind + 3 > argc) {
      std::cerr << "Usage: " << argv[0] << " quant <index_file> <reads.fastq> <output>" << std::endl;
      return 1;
    }
    quantification(argv[optind], argv[optind + 1], argv[optind + 2], kmer_lengths, opt);
  } else if (mode == "selftest-admit" && optind + 1 <= argc) {
    // host-logic probes used by tests/ (no GPU needed): what the FASTQ scan admits for this -k list
    FastqFile fq(argv[optind]);
    uint64_t seen = 0;
    auto recs = fq.admitted_records(*std::max_element(kmer_lengths.begin(), kmer_lengths.end()),
                                    opt.threads > 0 ? opt.threads : 4, &seen);
    std::cout << "records " << seen << " admitted " << recs.size() << "\n";
    for (auto& r : recs)
      std::cout << std::string(fq.data() + r.id_off, r.id_len) << "\t" << std::string(fq.data() + r.seq_off, r.seq_len) << "\n";
  } else if (mode == "selftest-stream" && optind + 1 <= argc) {
    // the streaming scanner + admission + packing of quant mode, chunked small on purpose (no GPU needed)
    FastqFile fq(argv[optind]);
    FastqScanner sc(fq);
    IdSet ids(fq.size() / 150 + 1024);
    const uint32_t kmax = *std::max_element(kmer_lengths.begin(), kmer_lengths.end());
    RawChunk c;
    bool unique = true;
    uint64_t admitted = 0;
    std::ostringstream body;
    while (sc.next(3, 200, &c)) {
      if (c.recs.empty()) continue;
      std::vector<uint32_t> mem((c.seq_bytes + 4 * c.recs.size()) / 16 + 16 + 2 * c.recs.size());
      PackedView v;
      v.words = mem.data();
      v.base_off = mem.data() + (c.seq_bytes + 4 * c.recs.size()) / 16 + 16;
      v.len = v.base_off + c.recs.size();
      unique &= admit_and_pack(fq.data(), c, kmax, &ids, &v);
      admitted += v.n_reads;
      for (uint32_t i = 0; i < v.n_reads; ++i) {
        std::string sq(v.len[i], '?');
        for (uint32_t j = 0; j < v.len[i]; ++j) {
          const uint64_t b = (uint64_t)v.base_off[i] + j;
          sq[j] = "ACGT"[(v.words[b >> 4] >> (2 * (b & 15))) & 3];
        }
        body << sq << "\n";
      }
    }
    std::cout << "records " << sc.records_seen() << " admitted " << admitted << " unique " << (unique ? 1 : 0) << "\n" << body.str();
  } else if (mode == "selftest-segments" && optind + 2 <= argc) {
    // the parallel scan of quant mode (segments at line starts, scanned independently), segment by segment in file
    // order: prints what selftest-stream prints when the file is a regular four-line FASTQ, "regular 0" otherwise
    FastqFile fq(argv[optind]);
    const size_t target = (size_t)std::max(1L, atol(argv[optind + 1]));
    const std::vector<FastqSegment> segs = split_fastq(fq, target, opt.threads > 0 ? opt.threads : 4);
    IdSet ids(fq.size() / 150 + 1024);
    const uint32_t kmax = *std::max_element(kmer_lengths.begin(), kmer_lengths.end());
    RawChunk c;
    bool unique = true, regular = true;
    uint64_t admitted = 0, seen = 0;
    std::ostringstream body;
    for (const FastqSegment& sg : segs) {
      FastqSegmentScanner sc(fq, sg);
      while (sc.next(3, 200, &c)) {
        if (c.recs.empty()) continue;
        std::vector<uint32_t> mem((c.seq_bytes + 4 * c.recs.size()) / 16 + 16 + 2 * c.recs.size());
        PackedView v;
        v.words = mem.data();
        v.base_off = mem.data() + (c.seq_bytes + 4 * c.recs.size()) / 16 + 16;
        v.len = v.base_off + c.recs.size();
        unique &= admit_and_pack(fq.data(), c, kmax, &ids, &v);
        admitted += v.n_reads;
        for (uint32_t i = 0; i < v.n_reads; ++i) {
          std::string sq(v.len[i], '?');
          for (uint32_t j = 0; j < v.len[i]; ++j) {
            const uint64_t b = (uint64_t)v.base_off[i] + j;
            sq[j] = "ACGT"[(v.words[b >> 4] >> (2 * (b & 15))) & 3];
          }
          body << sq << "\n";
        }
      }
      seen += sc.records_seen();
      regular &= sc.regular();
    }
    std::cout << "segments " << segs.size() << " regular " << (regular ? 1 : 0) << "\n";
    if (regular) std::cout << "records " << seen << " admitted " << admitted << " unique " << (unique ? 1 : 0) << "\n" << body.str();
  } else if (mode == "selftest-ingest" && optind + 1 <= argc) {
    // host side of quant mode's ingest pipeline alone (scan -> admit + pack in workers, batches dropped): reads/s
    const double t0 = now();
    FastqFile fq(argv[optind]);
    const uint32_t kmax = *std::max_element(kmer_lengths.begin(), kmer_lengths.end());
    const int W = std::max(1, (opt.threads > 0 ? opt.threads : (int)std::thread::hardware_concurrency()) - 1);
    uint64_t fq_lines = 0;
    const std::vector<FastqSegment> segs = split_fastq(fq, (size_t)32 << 20, W, &fq_lines);
    IdSet ids(fq_lines / 4 + 1024);
    std::atomic<uint64_t> admitted{0}, bases{0}, seen{0};
    std::atomic<bool> unique{true}, regular{true};
    std::atomic<size_t> next_seg{0};
    std::vector<std::thread> workers;
    for (int w = 0; w < W; ++w)
      workers.emplace_back([&] {
        RawChunk c;
        std::vector<uint32_t> mem;
        for (size_t si; (si = next_seg++) < segs.size();) {
          FastqSegmentScanner sc(fq, segs[si]);
          while (sc.next(1u << 18, 1ull << 26, &c)) {
            if (c.recs.empty()) continue;
            if (getenv("SQ_SCAN_ONLY")) { admitted += c.recs.size(); continue; }
            const size_t words = (c.seq_bytes + 4 * c.recs.size()) / 16 + 16;
            mem.resize(words + 2 * c.recs.size());
            PackedView v;
            v.words = mem.data();
            v.base_off = mem.data() + words;
            v.len = v.base_off + c.recs.size();
            if (!admit_and_pack(fq.data(), c, kmax, &ids, &v)) unique = false;
            admitted += v.n_reads;
            bases += v.n_bases;
          }
          seen += sc.records_seen();
          if (!sc.regular()) regular = false;
        }
      });
    for (auto& w : workers) w.join();
    const double dt = now() - t0;
    std::cout << "records " << seen << " admitted " << admitted << " unique " << (unique ? 1 : 0) << " regular " << (regular ? 1 : 0) << " threads " << W
              << " seconds " << dt << " reads_per_s " << (admitted / dt) << " MB_per_s " << (fq.size() / dt / 1e6) << "\n";
  } else if (mode == "selftest-fasta" && optind + 1 <= argc) {
    for (auto& r : load_fasta(argv[optind])) std::cout << r.id << "\t" << r.sequence << "\n";
  } else if (mode == "selftest-index" && optind + 1 <= argc) {
    IndexData idx;
    if (!read_index(argv[optind], &idx, true)) return 3;
    std::cout << "ks";
    for (auto k : idx.ks) std::cout << " " << k;
    std::cout << "\nT " << idx.names.size() << "\n";
    for (auto& kv : idx.maps) std::cout << "map " << kv.first << " keys " << kv.second.keys.size() << " postings " << kv.second.tid.size() << "\n";
    if (optind + 2 <= argc) write_index(argv[optind + 1], idx.ks, idx.names, idx.sequences, idx.maps);
  } else {
    std::cerr << "Invalid mode. Please choose 'index' or 'quant'." << std::endl;
    return 1;
  }
  return 0;
}

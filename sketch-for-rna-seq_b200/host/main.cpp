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

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <condition_variable>
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

[[noreturn]] void die(sq_engine* e, int rc, const char* what) {
  std::cerr << "sketchquant: " << what << " failed (" << rc << "): " << sq_last_error(e) << std::endl;
  std::exit(2);
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

// quantification (main.cpp:165-197)
void quantification(const std::string& index_path, const std::string& reads_path, const std::string& output_path,
                    std::vector<unsigned>& kmer_lengths, const Options& opt) {
  const double t_start = now();
  IndexData idx;
  const bool have_index = read_index(index_path, &idx, false, opt.index_cache);
  if (have_index) kmer_lengths.assign(idx.ks.begin(), idx.ks.end());  // load_index overwrites the -k list (main.cpp:174)
  std::cout << "Loading index completed" << std::endl;
  const double t_index = now();
  if (kmer_lengths.empty()) throw std::runtime_error("no k-mer length available");
  const uint32_t kmax = *std::max_element(kmer_lengths.begin(), kmer_lengths.end());

  if (getenv("SQ_TRACE")) fprintf(stderr, "[sq trace] index file parsed            %.3f s\n", now() - t_start);
  int threads = opt.threads > 0 ? opt.threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const size_t T = idx.names.size();
  std::vector<double> pi(T, 0.0), numreads(T, 0.0);
  std::vector<uint8_t> present(T, 0);
  double t_reads = 0, t_chain = 0, t_em = 0;
  sq_stats stats;
  memset(&stats, 0, sizeof(stats));
  uint64_t n_seen = 0, R = 0;

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
    // One host thread per GPU.  Each creates its engine and index replica right away (CUDA context, table
    // build) while the main thread scans the FASTQ; then it packs and pushes its contiguous share of records.
    const size_t chunk_reads = 1u << 21;
    std::vector<std::thread> workers;
    std::vector<double> tr(G, 0), tc(G, 0), te(G, 0);
    std::vector<std::vector<double>> pis(G), nrs(G);
    std::vector<std::vector<uint8_t>> prs(G);
    std::vector<sq_stats> sts(G);
    std::mutex mu;
    std::condition_variable cv;
    bool reads_ready = false, reads_failed = false;
    const FastqFile* fqp = nullptr;
    std::vector<FastqFile::Rec> recs;
    for (int g = 0; g < G; ++g) {
      workers.emplace_back([&, g] {
        sq_engine* e = nullptr;
        int rc = sq_create(&e, g, (uint32_t)k32.size(), k32.data(), sq_threshold_from_fraction((double)opt.sketch_size),
                           opt.chain_fraction, T);
        if (rc != SQ_OK) die(nullptr, rc, "sq_create");
        eng[g] = e;
        if (getenv("SQ_TRACE")) fprintf(stderr, "[sq trace] gpu %d engine created           %.3f s since start\n", g, now() - t_start);
        if (opt.report.size()) sq_set_profiling(e, 1);
        for (size_t ki = 0; ki < k32.size(); ++ki) {
          auto it = idx.maps.find(k32[ki]);
          if (it == idx.maps.end()) continue;  // no map for this k: contributes nothing (sparse_chaining.cpp:51-53)
          const Postings& P = it->second;
          SQ(e, sq_load_index(e, (uint32_t)ki, P.keys.size(), P.keys.data(), P.off.data(), P.tid.data()));
        }
        if (G > 1) SQ(e, sq_comm_init(e, G, g, uid));
        if (getenv("SQ_TRACE")) fprintf(stderr, "[sq trace] gpu %d engine+index ready      %.3f s since start\n", g, now() - t_start);
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return reads_ready || reads_failed; });
          if (reads_failed) return;
        }
        const FastqFile& fq = *fqp;
        const uint64_t per = (R + G - 1) / G, lo = std::min<uint64_t>(R, g * per), hi = std::min<uint64_t>(R, lo + per);
        PackedBatch pb[2];
        std::vector<const char*> ptr;
        std::vector<uint32_t> len;
        int which = 0;
        for (uint64_t b = lo; b < hi; b += chunk_reads) {
          const uint64_t n = std::min<uint64_t>(chunk_reads, hi - b);
          ptr.resize(n);
          len.resize(n);
          for (uint64_t i = 0; i < n; ++i) {
            ptr[i] = fq.data() + recs[b + i].seq_off;
            len[i] = recs[b + i].seq_len;
          }
          PackedBatch& P = pb[which];
          which ^= 1;
          pack_sequences(ptr.data(), len.data(), n, std::max(1, threads / G), &P);
          SQ(e, sq_push_reads(e, P.words.data(), P.words.size(), P.base_off.data(), P.len.data(), (uint32_t)n));
        }
        tr[g] = now();
        SQ(e, sq_sync(e));
        tc[g] = now();
        pis[g].resize(T);
        nrs[g].resize(T);
        prs[g].resize(T);
        int iters = 0;
        SQ(e, sq_finish(e, G > 1 ? 0 : R, opt.em_iters, opt.em_tol, pis[g].data(), nrs[g].data(), prs[g].data(), &iters));
        te[g] = now();
        sq_get_stats(e, &sts[g]);
      });
    }
    // main thread: FASTQ scan (an unopenable file throws like upstream; release the workers first)
    std::unique_ptr<FastqFile> fq;
    try {
      fq.reset(new FastqFile(reads_path));
      recs = fq->admitted_records(kmax, threads, &n_seen);
    } catch (...) {
      { std::lock_guard<std::mutex> lk(mu); reads_failed = true; }
      cv.notify_all();
      for (auto& w : workers) w.join();
      throw;
    }
    R = recs.size();
    fqp = fq.get();
    { std::lock_guard<std::mutex> lk(mu); reads_ready = true; }
    cv.notify_all();
    for (auto& w : workers) w.join();
    std::cout << "Loading read completed" << std::endl;
    std::cout << "Sparse chaining completed" << std::endl;
    std::cout << "EM estimation completed" << std::endl;
    std::cout << "Read assignment completed" << std::endl;
    pi = pis[0];
    numreads = nrs[0];
    present = prs[0];
    t_reads = *std::max_element(tr.begin(), tr.end());
    t_chain = *std::max_element(tc.begin(), tc.end());
    t_em = *std::max_element(te.begin(), te.end());
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
      << ", \"gpus\": " << opt.gpus << ", \"host_threads\": " << threads
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

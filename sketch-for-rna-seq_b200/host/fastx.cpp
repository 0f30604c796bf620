#include "fastx.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace sqhost {

namespace {
struct Tables {
  bool valid[256];
  uint8_t code[256];
  Tables() {
    memset(valid, 0, sizeof(valid));
    memset(code, 0, sizeof(code));
    valid[(int)'A'] = valid[(int)'C'] = valid[(int)'G'] = valid[(int)'T'] = true;
    code[(int)'C'] = code[(int)'c'] = 1;
    code[(int)'G'] = code[(int)'g'] = 2;
    code[(int)'T'] = code[(int)'t'] = code[(int)'U'] = code[(int)'u'] = 3;
  }
};
const Tables kTab;

template <class F>
void parallel_for(size_t n, int n_threads, F f) {
  n_threads = std::max(1, std::min<int>(n_threads, (int)std::max<size_t>(1, n / 4096)));
  if (n_threads == 1) { f(0, n, 0); return; }
  std::vector<std::thread> th;
  const size_t per = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const size_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
    th.emplace_back([=] { f(lo, hi, t); });
  }
  for (auto& t : th) t.join();
}
}  // namespace

bool is_valid_sequence(const char* s, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (!kTab.valid[(unsigned char)s[i]]) return false;
  return true;
}

std::vector<FastaRecord> load_fasta(const std::string& path) {
  std::ifstream in(path, std::ios::in | std::ios::binary);
  if (!in) throw std::runtime_error("Could not open FASTA file: " + path);
  std::vector<FastaRecord> out;
  std::unordered_set<std::string> seen;
  std::string line, seq, id;
  auto flush = [&](bool check) {
    if (id.empty()) return;
    if (check && !is_valid_sequence(seq.data(), seq.size())) return;
    if (seen.insert(id).second) out.push_back({id, seq});  // emplace: the first record of an id wins
  };
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    if (line[0] == '>') {
      flush(true);
      id = line.substr(1, line.find(' ') - 1);  // data_io.cpp:67
      seq.clear();
    } else {
      seq += line;
    }
  }
  flush(false);  // the last record is stored without the validity check (data_io.cpp:75-77)
  return out;
}

FastqFile::FastqFile(const std::string& path) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) throw std::runtime_error("Could not open FASTQ file: " + path);
  struct stat st;
  fstat(fd, &st);
  size_ = (size_t)st.st_size;
  if (size_) {
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { close(fd); throw std::runtime_error("Could not open FASTQ file: " + path); }
    madvise(m, size_, MADV_SEQUENTIAL);
    data_ = static_cast<const char*>(m);
  }
  close(fd);
}

FastqFile::~FastqFile() {
  if (data_) munmap(const_cast<char*>(data_), size_);
}

namespace {
double wall() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct Trace {
  bool on = getenv("SQ_TRACE") != nullptr;
  double t = wall();
  void lap(const char* what) {
    if (!on) return;
    const double n = wall();
    fprintf(stderr, "[sq trace] %-28s %.3f s\n", what, n - t);
    t = n;
  }
};
}  // namespace

std::vector<FastqFile::Rec> FastqFile::admitted_records(uint32_t max_k, int n_threads, uint64_t* n_seen) const {
  Trace tr;
  // 1. line starts: newline positions are found in parallel, the record state machine (which depends on every
  //    earlier line) then runs over the line table only
  const char* d = data_;
  const size_t n = size_;
  int nt = std::max(1, n_threads);
  std::vector<std::vector<uint64_t>> nl(nt);
  parallel_for(n, nt, [&](size_t lo, size_t hi, int t) {
    auto& v = nl[t];
    v.reserve((hi - lo) / 64 + 16);
    const char* p = d + lo;
    const char* e = d + hi;
    while (p < e) {
      const char* q = static_cast<const char*>(memchr(p, '\n', e - p));
      if (!q) break;
      v.push_back((uint64_t)(q - d));
      p = q + 1;
    }
  });
  tr.lap("fastq newline scan");
  std::vector<uint64_t> line_end;  // position of the terminating '\n' (or n for an unterminated last line)
  size_t total = 0;
  for (auto& v : nl) total += v.size();
  line_end.reserve(total + 1);
  for (auto& v : nl) line_end.insert(line_end.end(), v.begin(), v.end());
  if (n && (line_end.empty() || line_end.back() != n - 1)) line_end.push_back(n);
  const size_t n_lines = line_end.size();
  auto line_begin = [&](size_t i) { return i == 0 ? (uint64_t)0 : line_end[i - 1] + 1; };
  tr.lap("fastq line table");
  // 2. records (sequential over lines, O(1) per line)
  std::vector<Rec> recs;
  recs.reserve(n_lines / 4 + 1);
  for (size_t i = 0; i < n_lines;) {
    const uint64_t b = line_begin(i), e = line_end[i];
    ++i;
    if (e == b || d[b] != '@') continue;  // main.cpp:121-123
    Rec r;
    r.id_off = b + 1;
    r.id_len = (uint32_t)(e - b - 1);
    if (i < n_lines) {
      r.seq_off = line_begin(i);
      r.seq_len = (uint32_t)(line_end[i] - r.seq_off);
    } else {
      r.seq_off = n;
      r.seq_len = 0;  // getline on EOF leaves an empty sequence
    }
    i += 3;  // sequence, '+', quality
    recs.push_back(r);
  }
  if (n_seen) *n_seen = recs.size();
  tr.lap("fastq record scan");
  // 3. admission (main.cpp:131-138), in parallel
  std::vector<uint8_t> ok(recs.size());
  parallel_for(recs.size(), nt, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i)
      ok[i] = recs[i].seq_len >= max_k && is_valid_sequence(d + recs[i].seq_off, recs[i].seq_len);
  });
  tr.lap("fastq admission");
  // 4. duplicate ids: the LAST admitted record of an id survives (read_sketches[read.id] = ..., main.cpp:147),
  //    at the position of the FIRST occurrence is irrelevant (unordered_map has no order): we keep file order of
  //    the surviving records.  Sharded by a hash of the id: every thread owns the ids of one shard, walks the
  //    records in file order and remembers the last record of each id, so no locks are needed.
  const size_t nrec = recs.size();
  std::vector<uint64_t> idh(nrec);
  parallel_for(nrec, nt, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      uint64_t h = 0xcbf29ce484222325ull;
      const unsigned char* p = reinterpret_cast<const unsigned char*>(d + recs[i].id_off);
      for (uint32_t j = 0; j < recs[i].id_len; ++j) { h ^= p[j]; h *= 0x100000001b3ull; }
      idh[i] = h ^ (h >> 29);
    }
  });
  std::vector<uint8_t> keep(nrec, 0);
  const int shards = nt;
  {
    std::vector<std::thread> th;
    for (int sh = 0; sh < shards; ++sh)
      th.emplace_back([&, sh] {
        std::unordered_map<std::string_view, uint32_t> last;  // id -> index of its last admitted record
        last.reserve(nrec / shards * 2 + 16);
        for (size_t i = 0; i < nrec; ++i) {
          if (!ok[i] || (int)((idh[i] >> 40) % shards) != sh) continue;
          std::string_view id(d + recs[i].id_off, recs[i].id_len);
          auto ins = last.emplace(id, (uint32_t)i);
          if (!ins.second) {
            keep[ins.first->second] = 0;
            ins.first->second = (uint32_t)i;
          }
          keep[i] = 1;
        }
      });
    for (auto& t : th) t.join();
  }
  std::vector<Rec> out;
  out.reserve(nrec);
  for (size_t i = 0; i < nrec; ++i)
    if (keep[i]) out.push_back(recs[i]);
  tr.lap("fastq duplicate ids");
  return out;
}

void pack_sequences(const char* const* seqs, const uint32_t* lens, size_t n, int n_threads, PackedBatch* out) {
  out->base_off.resize(n);
  out->len.assign(lens, lens + n);
  uint64_t pos = 0;
  for (size_t i = 0; i < n; ++i) {
    out->base_off[i] = (uint32_t)pos;
    pos += ((uint64_t)lens[i] + 3) & ~3ull;
    if (pos >= 0xFFFFFFF0ull) throw std::runtime_error("pack_sequences: batch exceeds 2^32 bases");
  }
  out->n_bases = n ? (uint64_t)out->base_off[n - 1] + lens[n - 1] : 0;
  const size_t n_words = (((out->n_bases + 15) / 16 + 3) & ~(size_t)3) + 4;
  out->words.assign(n_words, 0);
  uint8_t* bytes = reinterpret_cast<uint8_t*>(out->words.data());  // 4 bases per byte, sequences start on bytes
  parallel_for(n, n_threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      const unsigned char* s = reinterpret_cast<const unsigned char*>(seqs[i]);
      uint8_t* dst = bytes + out->base_off[i] / 4;
      const uint32_t L = lens[i];
      uint32_t j = 0;
      for (; j + 4 <= L; j += 4)
        dst[j / 4] = (uint8_t)(kTab.code[s[j]] | kTab.code[s[j + 1]] << 2 | kTab.code[s[j + 2]] << 4 | kTab.code[s[j + 3]] << 6);
      if (j < L) {
        uint8_t v = 0;
        for (uint32_t q = 0; j + q < L; ++q) v |= (uint8_t)(kTab.code[s[j + q]] << (2 * q));
        dst[j / 4] = v;
      }
    }
  });
}

}  // namespace sqhost

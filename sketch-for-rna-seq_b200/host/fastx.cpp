#include "fastx.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
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

bool FastqScanner::next(size_t max_records, uint64_t max_seq_bytes, RawChunk* out) {
  out->recs.clear();
  out->seq_bytes = 0;
  const char* d = d_;
  const size_t n = n_;
  auto line_end = [&](size_t p) {  // position of the line's '\n', or n
    const void* q = p < n ? memchr(d + p, '\n', n - p) : nullptr;
    return q ? (size_t)(static_cast<const char*>(q) - d) : n;
  };
  while (pos_ < n && out->recs.size() < max_records && out->seq_bytes < max_seq_bytes) {
    const size_t b = pos_, e = line_end(b);
    pos_ = e + 1;
    if (e == b || d[b] != '@') continue;  // main.cpp:121-123
    FastqFile::Rec r;
    r.id_off = b + 1;
    r.id_len = (uint32_t)(e - b - 1);
    if (pos_ < n) {
      const size_t se = line_end(pos_);
      r.seq_off = pos_;
      r.seq_len = (uint32_t)(se - pos_);
      pos_ = se + 1;
    } else {
      r.seq_off = n;
      r.seq_len = 0;  // getline on EOF leaves an empty sequence
    }
    for (int skip = 0; skip < 2 && pos_ < n; ++skip) pos_ = line_end(pos_) + 1;  // '+', quality
    out->recs.push_back(r);
    out->seq_bytes += r.seq_len;
    ++seen_;
  }
  if (pos_ > n) pos_ = n;
  return !out->recs.empty() || pos_ < n;
}

// ---- parallel scanning (see fastx.hpp)
#if defined(__x86_64__)
__attribute__((target("avx2"))) static uint64_t count_newlines_avx2(const char* p, size_t n) {
  uint64_t c = 0;
  size_t i = 0;
  const __m256i nl = _mm256_set1_epi8('\n');
  for (; i + 32 <= n; i += 32)
    c += (uint64_t)__builtin_popcount((unsigned)_mm256_movemask_epi8(
        _mm256_cmpeq_epi8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i)), nl)));
  for (; i < n; ++i) c += p[i] == '\n';
  return c;
}
#endif
static uint64_t count_newlines(const char* p, size_t n) {
#if defined(__x86_64__)
  if (__builtin_cpu_supports("avx2")) return count_newlines_avx2(p, n);
#endif
  uint64_t c = 0;
  for (size_t i = 0; i < n; ++i) c += p[i] == '\n';
  return c;
}

std::vector<FastqSegment> split_fastq(const FastqFile& f, size_t target_bytes, int n_threads, uint64_t* total_lines) {
  const char* d = f.data();
  const size_t n = f.size();
  const size_t S = std::max<size_t>(1, (n + std::max<size_t>(target_bytes, 1) - 1) / std::max<size_t>(target_bytes, 1));
  std::vector<FastqSegment> seg(S);
  // boundaries: the first line start at or after i * n / S
  std::vector<size_t> cut(S + 1, n);
  cut[0] = 0;
  for (size_t i = 1; i < S; ++i) {
    size_t o = (size_t)((unsigned __int128)n * i / S);
    if (o <= cut[i - 1]) o = cut[i - 1];
    if (o == 0 || d[o - 1] == '\n') { cut[i] = o; continue; }
    const void* q = memchr(d + o, '\n', n - o);
    cut[i] = q ? (size_t)(static_cast<const char*>(q) - d) + 1 : n;
  }
  for (size_t i = 1; i <= S; ++i) cut[i] = std::max(cut[i], cut[i - 1]);
  std::vector<uint64_t> lines(S, 0);
  const unsigned nt = (unsigned)std::max(1, std::min<int>(n_threads, (int)S));
  std::atomic<size_t> next{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&] {
      for (size_t i; (i = next++) < S;) lines[i] = count_newlines(d + cut[i], cut[i + 1] - cut[i]);
    });
  for (auto& t : th) t.join();
  uint64_t before = 0;
  for (size_t i = 0; i < S; ++i) {
    seg[i].begin = cut[i];
    seg[i].end = cut[i + 1];
    seg[i].first_line = before;
    before += lines[i];  // every segment but the last ends behind a newline: whole lines only
  }
  if (total_lines) *total_lines = before + (n > 0 && d[n - 1] != '\n' ? 1 : 0);
  return seg;
}

FastqSegmentScanner::FastqSegmentScanner(const FastqFile& f, const FastqSegment& seg)
    : d_(f.data()), n_(f.size()), pos_(seg.begin), end_(seg.end) {
  // lines 4i+1 .. 4i+3 at the head of the segment belong to a record whose header is in the segment before
  for (uint64_t phase = seg.first_line & 3; phase != 0 && pos_ < end_; phase = (phase + 1) & 3) {
    const void* q = memchr(d_ + pos_, '\n', n_ - pos_);
    pos_ = q ? (size_t)(static_cast<const char*>(q) - d_) + 1 : n_;
  }
}

bool FastqSegmentScanner::next(size_t max_records, uint64_t max_seq_bytes, RawChunk* out) {
  out->recs.clear();
  out->seq_bytes = 0;
  const char* d = d_;
  const size_t n = n_;
  auto line_end = [&](size_t p) {  // position of the line's '\n', or n
    const void* q = p < n ? memchr(d + p, '\n', n - p) : nullptr;
    return q ? (size_t)(static_cast<const char*>(q) - d) : n;
  };
  while (regular_ && pos_ < end_ && out->recs.size() < max_records && out->seq_bytes < max_seq_bytes) {
    const size_t b = pos_, e = line_end(b);
    if (e == b || d[b] != '@') {  // a header position without a header: not a regular four-line FASTQ
      regular_ = false;
      break;
    }
    pos_ = e + 1;
    FastqFile::Rec r;
    r.id_off = b + 1;
    r.id_len = (uint32_t)(e - b - 1);
    if (pos_ < n) {
      const size_t se = line_end(pos_);
      r.seq_off = pos_;
      r.seq_len = (uint32_t)(se - pos_);
      pos_ = se + 1;
    } else {
      r.seq_off = n;
      r.seq_len = 0;  // getline on EOF leaves an empty sequence
    }
    for (int skip = 0; skip < 2 && pos_ < n; ++skip) pos_ = line_end(pos_) + 1;  // '+', quality
    out->recs.push_back(r);
    out->seq_bytes += r.seq_len;
    ++seen_;
  }
  if (pos_ > n) pos_ = n;
  return regular_ && (!out->recs.empty() || pos_ < end_);
}

// The fill level is kept per 1/64 of the table (64 counters, a cache line each): one shared counter would be the
// one line every inserting thread fights for.
IdSet::IdSet(uint64_t expected) {
  uint64_t cap = 1 << 16;
  while (cap < expected * 2 + 16) cap <<= 1;
  mask_ = cap - 1;
  limit_ = (cap - cap / 4) / 64;  // per stripe
  stripe_shift_ = 0;
  while ((cap >> stripe_shift_) > 64) ++stripe_shift_;
  slots_ = calloc(cap, sizeof(uint64_t));
  count_ = aligned_alloc(64, 64 * 64);
  if (!slots_ || !count_) throw std::runtime_error("IdSet: out of memory");
  memset(count_, 0, 64 * 64);
}
IdSet::~IdSet() {
  free(slots_);
  free(count_);
}
void IdSet::prefetch(uint64_t h) const {
  if (h == 0) h = 0x9E3779B97F4A7C15ull;
  __builtin_prefetch(static_cast<const uint64_t*>(slots_) + (((h * 0x9E3779B97F4A7C15ull) >> 20) & mask_), 1, 0);
}
bool IdSet::insert(uint64_t h) {
  if (h == 0) h = 0x9E3779B97F4A7C15ull;  // 0 marks a free slot
  uint64_t* slots = static_cast<uint64_t*>(slots_);
  const uint64_t s0 = ((h * 0x9E3779B97F4A7C15ull) >> 20) & mask_;
  uint64_t* count = static_cast<uint64_t*>(count_) + (s0 >> stripe_shift_) * 8;
  if (__atomic_load_n(count, __ATOMIC_RELAXED) >= limit_) return false;
  for (uint64_t s = s0;; ++s) {
    uint64_t* slot = slots + (s & mask_);
    uint64_t cur = __atomic_load_n(slot, __ATOMIC_RELAXED);
    if (cur == 0) {
      if (__atomic_compare_exchange_n(slot, &cur, h, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
        __atomic_fetch_add(count, 1, __ATOMIC_RELAXED);
        return true;
      }
    }
    if (cur == h) return false;
  }
}

#if defined(__x86_64__)
// 32 bases -> 8 packed bytes; returns false when a byte is not one of A C G T.  code = x ^ (x >> 1) with
// x = (c >> 1) & 3 maps A C G T to 0 1 2 3; maddubs / madd fold four 2-bit codes into one byte.
__attribute__((target("avx2"))) static inline bool pack32_avx2(const unsigned char* s, uint8_t* dst) {
  const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s));
  const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(c, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(c, _mm256_set1_epi8('C'))),
                                     _mm256_or_si256(_mm256_cmpeq_epi8(c, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(c, _mm256_set1_epi8('T'))));
  if (_mm256_movemask_epi8(ok) != -1) return false;
  const __m256i x = _mm256_and_si256(_mm256_srli_epi16(c, 1), _mm256_set1_epi8(3));
  const __m256i code = _mm256_xor_si256(x, _mm256_and_si256(_mm256_srli_epi16(x, 1), _mm256_set1_epi8(1)));
  const __m256i p16 = _mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401));    // c0 + 4*c1 per 16-bit lane
  const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi32(0x00100001));    // + 16*(c2 + 4*c3) per 32-bit lane
  const __m256i sh = _mm256_shuffle_epi8(p32, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                               0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
  const uint32_t lo = (uint32_t)_mm256_extract_epi32(sh, 0), hi = (uint32_t)_mm256_extract_epi32(sh, 4);
  memcpy(dst, &lo, 4);
  memcpy(dst + 4, &hi, 4);
  return true;
}
static const bool kHaveAvx2 = __builtin_cpu_supports("avx2");
#else
static const bool kHaveAvx2 = false;
static inline bool pack32_avx2(const unsigned char*, uint8_t*) { return false; }
#endif

bool admit_and_pack(const char* d, const RawChunk& c, uint32_t max_k, IdSet* ids, PackedView* out) {
  // code table with an "invalid" bit: anything but upper-case ACGT refuses the read (data_io.cpp:17-34)
  static const struct Adm {
    uint8_t code[256];
    Adm() {
      memset(code, 0x80, sizeof(code));
      code[(int)'A'] = 0; code[(int)'C'] = 1; code[(int)'G'] = 2; code[(int)'T'] = 3;
    }
  } adm;
  bool unique = true;
  std::vector<uint64_t> hs;  // id hashes of the admitted reads: inserted at the end, slots prefetched ahead
  if (ids) hs.reserve(c.recs.size());
  uint64_t pos = 0;  // next free base (multiple of 4)
  uint32_t nr = 0;
  uint8_t* bytes = reinterpret_cast<uint8_t*>(out->words);
  for (const FastqFile::Rec& r : c.recs) {
    const uint32_t L = r.seq_len;
    if (L < max_k) continue;  // main.cpp:136-138
    const unsigned char* s = reinterpret_cast<const unsigned char*>(d + r.seq_off);
    uint8_t* dst = bytes + pos / 4;
    uint32_t bad = 0, j = 0;
    if (kHaveAvx2)
      for (; j + 32 <= L; j += 32)
        if (!pack32_avx2(s + j, dst + j / 4)) { bad = 0x80; break; }
    for (; j + 4 <= L && !(bad & 0x80); j += 4) {
      const uint32_t a = adm.code[s[j]], b = adm.code[s[j + 1]], cc = adm.code[s[j + 2]], e = adm.code[s[j + 3]];
      bad |= a | b | cc | e;
      dst[j / 4] = (uint8_t)(a | b << 2 | cc << 4 | e << 6);
    }
    if (j < L && !(bad & 0x80)) {
      uint32_t v = 0;
      for (uint32_t q = 0; j + q < L; ++q) { const uint32_t a = adm.code[s[j + q]]; bad |= a; v |= (a & 3) << (2 * q); }
      dst[j / 4] = (uint8_t)v;
    }
    if (bad & 0x80) continue;  // refused: the bytes written are overwritten by the next read
    if (ids) {
      uint64_t h = 0xcbf29ce484222325ull;
      const unsigned char* p = reinterpret_cast<const unsigned char*>(d + r.id_off);
      for (uint32_t q = 0; q < r.id_len; ++q) { h ^= p[q]; h *= 0x100000001b3ull; }
      hs.push_back(h ^ (h >> 29));
    }
    out->base_off[nr] = (uint32_t)pos;
    out->len[nr] = L;
    ++nr;
    out->n_bases = pos + L;
    pos += ((uint64_t)L + 3) & ~3ull;
  }
  for (size_t i = 0; i < hs.size(); ++i) {
    if (i + 12 < hs.size()) ids->prefetch(hs[i + 12]);
    if (!ids->insert(hs[i])) unique = false;
  }
  // zero the tail so that the word count can be rounded up to a multiple of 4 (+4 of padding)
  const uint64_t used_bytes = pos / 4;
  out->n_words = (((out->n_bases + 15) / 16 + 3) & ~(uint64_t)3) + 4;
  memset(bytes + used_bytes, 0, out->n_words * 4 - used_bytes);
  out->n_reads = nr;
  if (nr == 0) { out->n_bases = 0; }
  return unique;
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

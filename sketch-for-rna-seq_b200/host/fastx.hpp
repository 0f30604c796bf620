// FASTA / FASTQ ingestion with the reference's exact record and admission semantics, and 2-bit packing.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sqhost {

// reference src/data_io.cpp:17-34: true iff every character is one of A C G T (upper case)
bool is_valid_sequence(const char* s, size_t n);

struct FastaRecord {
  std::string id;
  std::string sequence;
};
// reference src/data_io.cpp:47-80: id = header up to the first space; records with an invalid character are
// dropped, EXCEPT the last record of the file, which is stored unchecked; duplicate ids: first wins.
// Throws std::runtime_error("Could not open FASTA file: ...") like the reference.
std::vector<FastaRecord> load_fasta(const std::string& path);

// A whole FASTQ file mapped into memory and cut into the records process_fastq_single_pass() would see
// (src/main.cpp:120-148): a non-empty line starting with '@' opens a record, id = rest of that line, the next
// three lines are sequence, '+', quality whatever they contain.
class FastqFile {
 public:
  explicit FastqFile(const std::string& path);  // throws std::runtime_error("Could not open FASTQ file: ...")
  ~FastqFile();
  struct Rec {
    uint64_t id_off, seq_off;
    uint32_t id_len, seq_len;
  };
  // records that pass admission (ACGT only, length >= max_k), with later duplicates of an id replacing earlier
  // ones (read_sketches[read.id] = ..., main.cpp:147).  n_threads host threads.
  std::vector<Rec> admitted_records(uint32_t max_k, int n_threads, uint64_t* n_records_seen) const;
  const char* data() const { return data_; }
  size_t size() const { return size_; }

 private:
  const char* data_ = nullptr;
  size_t size_ = 0;
};

// ---- streaming ingest: the file is cut into chunks of records while it is being scanned, every chunk is
// admitted + packed by a worker into a caller-supplied (page-locked) buffer and can be pushed to a GPU right
// away.  Duplicate read ids (last one wins, main.cpp:147) are a whole-file property: the chunks only RECORD every
// id in a concurrent set and report whether any id came twice; the caller then redoes the file through
// admitted_records().  (FASTQ ids are unique in practice; the exact path stays for files where they are not.)
struct RawChunk {
  std::vector<FastqFile::Rec> recs;
  uint64_t seq_bytes = 0;
};
// Sequential record scanner over a mapped FASTQ file (same state machine as admitted_records()).
class FastqScanner {
 public:
  explicit FastqScanner(const FastqFile& f) : d_(f.data()), n_(f.size()) {}
  // next chunk of at most max_records records / max_seq_bytes sequence bytes; false when the file is exhausted
  bool next(size_t max_records, uint64_t max_seq_bytes, RawChunk* out);
  uint64_t records_seen() const { return seen_; }

 private:
  const char* d_;
  size_t n_, pos_ = 0;
  uint64_t seen_ = 0;
};

// ---- scanning in parallel.  One sequential scanner tops out at ~3 GB/s (10 M reads/s), far below what the
// packing workers and the GPU take.  The file is therefore cut into segments at line starts; the number of lines
// before each segment is counted in parallel, and a segment is scanned on its own ASSUMING the file is a regular
// FASTQ: every record four lines, i.e. every line whose index is a multiple of 4 is a non-empty line starting
// with '@'.  Under that assumption upstream's state machine (main.cpp:107-151: skip lines until one starts with
// '@', take the next line as the sequence, skip two more) consumes exactly four lines per record, so the record
// at lines 4i .. 4i+3 is what it produces.  A segment scanner that meets a header position without a header
// reports the file as irregular, and the caller falls back to the sequential scan of the whole file (the same
// fall-back that duplicate read ids take): the assumption is checked, never trusted.
struct FastqSegment {
  size_t begin = 0, end = 0;  // bytes; begin is the start of a line, end the start of the next segment's first line (or EOF)
  uint64_t first_line = 0;    // index of the line that starts at `begin`
};
// segments of about target_bytes (at least one, even for an empty file), line counts taken by n_threads threads
// (total_lines, optional: newlines in the file + 1 for an unterminated last line -- four per record of a regular file)
std::vector<FastqSegment> split_fastq(const FastqFile& f, size_t target_bytes, int n_threads, uint64_t* total_lines = nullptr);

class FastqSegmentScanner {
 public:
  FastqSegmentScanner(const FastqFile& f, const FastqSegment& seg);
  // next chunk of the records whose header line starts inside the segment; false when the segment is exhausted
  // (or the file turned out irregular: check regular())
  bool next(size_t max_records, uint64_t max_seq_bytes, RawChunk* out);
  uint64_t records_seen() const { return seen_; }
  bool regular() const { return regular_; }

 private:
  const char* d_;
  size_t n_, pos_, end_;
  uint64_t seen_ = 0;
  bool regular_ = true;
};

// concurrent set of 64-bit id hashes (open addressing, lock-free)
class IdSet {
 public:
  explicit IdSet(uint64_t expected);
  ~IdSet();
  // false when h was already there (duplicate id, or a 64-bit collision: the caller falls back to the exact path
  // either way) or when the set is full
  bool insert(uint64_t h);
  void prefetch(uint64_t h) const;  // pull h's slot towards the cache ahead of insert()

 private:
  void* slots_;
  uint64_t mask_, limit_;
  unsigned stripe_shift_;
  void* count_;
};

// admitted reads of a chunk, packed into caller-supplied buffers (capacity: words for `seq_bytes` bases + 8,
// one base_off / len entry per record)
struct PackedView {
  uint32_t* words = nullptr;
  uint32_t* base_off = nullptr;
  uint32_t* len = nullptr;
  uint64_t n_words = 0, n_bases = 0;
  uint32_t n_reads = 0;
};
// admission (ACGT only, length >= max_k) + packing in one pass over the sequence bytes; ids of admitted reads go
// into `ids` (may be NULL).  Returns false if an admitted id was already in the set.
bool admit_and_pack(const char* file_data, const RawChunk& c, uint32_t max_k, IdSet* ids, PackedView* out);

// 2-bit packing (A=0 C=1 G=2 T=3, 16 bases per uint32, include/sketchquant.h).  Sequence i starts at base
// base_off[i] = next multiple of 4 after the previous one.  Lower-case acgt/u pack like upper case (the hash
// treats them alike); callers split sequences at other characters first.
struct PackedBatch {
  std::vector<uint32_t> words;     // padded to a multiple of 4 words plus 4
  std::vector<uint32_t> base_off, len;
  uint64_t n_bases = 0;
};
void pack_sequences(const char* const* seqs, const uint32_t* lens, size_t n, int n_threads, PackedBatch* out);

}  // namespace sqhost

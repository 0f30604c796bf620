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

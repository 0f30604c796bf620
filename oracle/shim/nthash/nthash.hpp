// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Header-only stand-in for the third-party ntHash library, which the reference
// links as -lnthash (/root/reference/build.sh:34) and includes as
// <nthash/nthash.hpp> (/root/reference/src/sketch.cpp:7, src/kmer.cpp:4,
// src/main.cpp:13) but does not vendor.  No version is pinned upstream; the
// symbol table of the reference's stale build/test binary shows the ntHash 2.x
// API (NtHash(const std::string&, uint8_t, uint16_t, size_t), roll(),
// get_forward_hash()).  This file restates only what the reference calls:
//
//   nthash::NtHash nth(sequence, 1, k);          (sketch.cpp:31)
//   while (nth.roll()) nth.get_forward_hash();   (sketch.cpp:32-33)
//
// Published algorithm (ntHash2, Kazemi et al. 2022): every base has a 64-bit
// seed; the forward hash of s_0..s_{k-1} is XOR_i srol^{k-1-i}(seed[s_i]) where
// srol rotates bits 63..33 and bits 32..0 as two independent lanes (31 and 33
// bits wide) by one position.  Rolling one base to the right:
//   fh' = srol(fh) ^ seed[in] ^ srol^k(seed[out]).
// Windows that contain a character outside ACGTU/acgtu are skipped.
//
// Pins: the four seeds below are the SEED_TAB entries of
// the reference's own build/test binary; tests/test_oracle_kats.py checks the
// forward-hash known-answer vectors listed there.
#ifndef ORACLE_NTHASH_STANDIN_HPP
#define ORACLE_NTHASH_STANDIN_HPP

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace nthash {

namespace standin {

inline uint64_t seed_of(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0x3c8bfbb395c60474ULL;
    case 'C': case 'c': return 0x3193c18562a02b4cULL;
    case 'G': case 'g': return 0x20323ed082572324ULL;
    case 'T': case 't': case 'U': case 'u': return 0x295549f54be24456ULL;
    default: return 0;  // SEED_N: marks the window as unusable
  }
}

// one step of the split rotation: [63..33] is a 31-bit ring, [32..0] a 33-bit ring
inline uint64_t srol1(uint64_t x) {
  const uint64_t carry = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
  return ((x << 1) & 0xFFFFFFFDFFFFFFFFULL) | carry;
}

inline uint64_t sroln(uint64_t x, unsigned n) {
  // lanes have coprime periods 31 and 33; reduce each separately
  const unsigned nh = n % 31, nl = n % 33;
  uint64_t hi = x >> 33;                 // 31 bits
  uint64_t lo = x & 0x1FFFFFFFFULL;      // 33 bits
  if (nh) hi = ((hi << nh) | (hi >> (31 - nh))) & 0x7FFFFFFFULL;
  if (nl) lo = ((lo << nl) | (lo >> (33 - nl))) & 0x1FFFFFFFFULL;
  return (hi << 33) | lo;
}

}  // namespace standin

class NtHash {
 public:
  NtHash(const std::string& seq, unsigned num_hashes, unsigned k, size_t pos = 0)
      : seq_(seq.data()), len_(seq.size()), k_(k), pos_(pos), started_(false), fwd_(0) {
    (void)num_hashes;
    if (k == 0) throw std::invalid_argument("[ntHash::stand-in] k must be greater than 0");
    if (len_ < k)
      throw std::invalid_argument("[ntHash::stand-in] sequence length (" + std::to_string(len_) +
                                  ") is smaller than k (" + std::to_string(k) + ")");
  }

  // Advance to the next usable window; false when the sequence is exhausted.
  bool roll() {
    if (!started_) return seek_and_init();
    if (pos_ + k_ >= len_) return false;  // current window is the last one
    const unsigned char in = static_cast<unsigned char>(seq_[pos_ + k_]);
    if (standin::seed_of(in) == 0) {
      pos_ += k_;  // every window touching the bad character is unusable
      return seek_and_init();
    }
    const unsigned char out = static_cast<unsigned char>(seq_[pos_]);
    fwd_ = standin::srol1(fwd_) ^ standin::seed_of(in) ^ standin::sroln(standin::seed_of(out), k_);
    ++pos_;
    return true;
  }

  uint64_t get_forward_hash() const { return fwd_; }
  size_t get_pos() const { return pos_; }

 private:
  bool seek_and_init() {
    // find the first window at or after pos_ made of valid characters only
    while (pos_ + k_ <= len_) {
      size_t bad = k_;
      for (size_t i = k_; i-- > 0;) {
        if (standin::seed_of(static_cast<unsigned char>(seq_[pos_ + i])) == 0) { bad = i; break; }
      }
      if (bad == k_) break;
      pos_ += bad + 1;
    }
    if (pos_ + k_ > len_) return false;
    uint64_t h = 0;
    for (size_t i = 0; i < k_; ++i)
      h = standin::srol1(h) ^ standin::seed_of(static_cast<unsigned char>(seq_[pos_ + i]));
    fwd_ = h;
    started_ = true;
    return true;
  }

  const char* seq_;
  size_t len_;
  unsigned k_;
  size_t pos_;
  bool started_;
  uint64_t fwd_;
};

}  // namespace nthash

#endif  // ORACLE_NTHASH_STANDIN_HPP

#include "index_file.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string_view>

namespace sqhost {

namespace {
struct Cursor {
  const unsigned char* p;
  const unsigned char* end;
  template <class T> T get() {
    T v{};
    if (p + sizeof(T) > end) throw std::runtime_error("index file is truncated");
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  std::string_view str(uint64_t n) {
    if (n > (uint64_t)(end - p)) throw std::runtime_error("index file is truncated");
    std::string_view s(reinterpret_cast<const char*>(p), n);
    p += n;
    return s;
  }
};
}  // namespace

bool read_index(const std::string& path, IndexData* out, bool keep_sequences) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    std::cerr << "Error: Unable to open file for reading: " << path << std::endl;  // data_io.cpp:239
    return false;
  }
  struct stat st;
  fstat(fd, &st);
  const size_t size = (size_t)st.st_size;
  void* map = size ? mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
  close(fd);
  if (size && map == MAP_FAILED) throw std::runtime_error("mmap failed for " + path);
  Cursor c{static_cast<const unsigned char*>(map), static_cast<const unsigned char*>(map) + size};
  out->ks.clear();
  out->names.clear();
  out->sequences.clear();
  out->maps.clear();
  const uint64_t nk = c.get<uint64_t>();
  for (uint64_t i = 0; i < nk; ++i) out->ks.push_back(c.get<uint32_t>());
  const uint64_t T = c.get<uint64_t>();
  // one dictionary id -> dense index built from the transcript section; postings are resolved through it
  std::unordered_map<std::string_view, uint32_t> dict;
  dict.reserve(T * 2);
  out->names.reserve(T);
  std::vector<std::string_view> views;
  views.reserve(T);
  for (uint64_t i = 0; i < T; ++i) {
    const uint64_t idlen = c.get<uint64_t>();
    std::string_view id = c.str(idlen);
    const uint64_t seqlen = c.get<uint64_t>();
    std::string_view seq = c.str(seqlen);
    (void)c.get<int32_t>();  // length: 0 when written by a libstdc++ build, never read by quant
    auto ins = dict.emplace(id, (uint32_t)out->names.size());
    if (ins.second) {
      out->names.emplace_back(id);
      if (keep_sequences) out->sequences.emplace_back(seq);
    } else if (keep_sequences) {
      out->sequences[ins.first->second] = std::string(seq);  // transcripts[id] = ... overwrites (data_io.cpp:271)
    }
  }
  const uint64_t nmaps = c.get<uint64_t>();
  for (uint64_t m = 0; m < nmaps; ++m) {
    const uint32_t k = c.get<uint32_t>();
    const uint64_t nkeys = c.get<uint64_t>();
    Postings& P = out->maps[k];
    P = Postings();
    P.keys.reserve(nkeys);
    P.off.reserve(nkeys + 1);
    P.off.push_back(0);
    // mapping[kmer] = vec (data_io.cpp:297): a repeated key would overwrite; files written by the reference
    // have distinct keys, which is what we rely on
    for (uint64_t j = 0; j < nkeys; ++j) {
      P.keys.push_back(c.get<uint32_t>());
      const uint64_t deg = c.get<uint64_t>();
      for (uint64_t d = 0; d < deg; ++d) {
        const uint64_t n = c.get<uint64_t>();
        std::string_view tid = c.str(n);
        auto it = dict.find(tid);
        if (it == dict.end()) {
          // a posting naming a transcript that is not in the transcript section cannot come from an index the
          // reference wrote (save_index writes both from the same run); skip it loudly
          std::cerr << "Warning: index posting names unknown transcript '" << tid << "', ignored" << std::endl;
          continue;
        }
        P.tid.push_back(it->second);
      }
      P.off.push_back(P.tid.size());
    }
  }
  if (map) munmap(map, size);
  std::cout << "Index loaded from " << path << std::endl;  // data_io.cpp:303
  return true;
}

bool write_index(const std::string& path, const std::vector<uint32_t>& ks, const std::vector<std::string>& names,
                 const std::vector<std::string>& sequences, const std::unordered_map<uint32_t, Postings>& maps) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) {
    std::cerr << "Error: Unable to open file for writing: " << path << std::endl;  // data_io.cpp:171
    return false;
  }
  std::vector<char> buf(1 << 22);
  setvbuf(f, buf.data(), _IOFBF, buf.size());
  auto w64 = [&](uint64_t v) { fwrite(&v, 8, 1, f); };
  auto w32 = [&](uint32_t v) { fwrite(&v, 4, 1, f); };
  w64(ks.size());
  for (uint32_t k : ks) w32(k);
  w64(names.size());
  for (size_t i = 0; i < names.size(); ++i) {
    w64(names[i].size());
    fwrite(names[i].data(), 1, names[i].size(), f);
    const std::string& s = sequences[i];
    w64(s.size());
    fwrite(s.data(), 1, s.size(), f);
    w32(0);  // Transcript.length as a libstdc++ build of the reference writes it (moved-from size, data_io.cpp:64)
  }
  w64(maps.size());
  for (const auto& kv : maps) {
    const Postings& P = kv.second;
    w32(kv.first);
    w64(P.keys.size());
    for (size_t i = 0; i < P.keys.size(); ++i) {
      w32(P.keys[i]);
      w64(P.off[i + 1] - P.off[i]);
      for (uint64_t j = P.off[i]; j < P.off[i + 1]; ++j) {
        const std::string& nm = names[P.tid[j]];
        w64(nm.size());
        fwrite(nm.data(), 1, nm.size(), f);
      }
    }
  }
  fclose(f);
  std::cout << "Index saved to " << path << std::endl;  // data_io.cpp:219
  return true;
}

}  // namespace sqhost

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

// ---- optional sidecar (<index>.sqidx): the parsed index as flat arrays, valid only for the exact file it was
// made from (size + mtime).  The reference format stores every posting as a length-prefixed id STRING, which
// costs ~1.5 s of dictionary lookups per run at human scale; the sidecar loads in ~0.1 s.  Opt-in
// (SQ_INDEX_CACHE=1 or --index-cache), never required, ignored when stale or unreadable.
namespace {
struct SidecarHeader {
  char magic[8];
  uint64_t src_size, src_mtime_ns, nk, T, names_bytes, nmaps;
};
const char kMagic[8] = {'S', 'Q', 'I', 'D', 'X', '0', '0', '1'};

bool stat_sig(const std::string& path, uint64_t* size, uint64_t* mtime_ns) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return false;
  *size = (uint64_t)st.st_size;
  *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
  return true;
}

bool load_sidecar(const std::string& path, IndexData* out) {
  uint64_t size, mtime;
  if (!stat_sig(path, &size, &mtime)) return false;
  FILE* f = fopen((path + ".sqidx").c_str(), "rb");
  if (!f) return false;
  bool ok = false;
  SidecarHeader h;
  do {
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0) break;
    if (h.src_size != size || h.src_mtime_ns != mtime) break;
    out->ks.resize(h.nk);
    if (h.nk && fread(out->ks.data(), 4, h.nk, f) != h.nk) break;
    std::vector<uint32_t> name_len(h.T);
    if (h.T && fread(name_len.data(), 4, h.T, f) != h.T) break;
    std::string blob(h.names_bytes, '\0');
    if (h.names_bytes && fread(&blob[0], 1, h.names_bytes, f) != h.names_bytes) break;
    out->names.clear();
    out->names.reserve(h.T);
    size_t pos = 0;
    for (uint64_t i = 0; i < h.T; ++i) { out->names.emplace_back(blob.data() + pos, name_len[i]); pos += name_len[i]; }
    bool bad = false;
    for (uint64_t m = 0; m < h.nmaps && !bad; ++m) {
      uint64_t hdr[3];
      if (fread(hdr, 8, 3, f) != 3) { bad = true; break; }
      Postings& P = out->maps[(uint32_t)hdr[0]];
      P.keys.resize(hdr[1]);
      P.off.resize(hdr[1] + 1);
      P.tid.resize(hdr[2]);
      if (hdr[1] && fread(P.keys.data(), 4, hdr[1], f) != hdr[1]) bad = true;
      if (!bad && fread(P.off.data(), 8, hdr[1] + 1, f) != hdr[1] + 1) bad = true;
      if (!bad && hdr[2] && fread(P.tid.data(), 4, hdr[2], f) != hdr[2]) bad = true;
    }
    ok = !bad;
  } while (false);
  fclose(f);
  if (!ok) { out->ks.clear(); out->names.clear(); out->maps.clear(); }
  return ok;
}

void save_sidecar(const std::string& path, const IndexData& idx) {
  uint64_t size, mtime;
  if (!stat_sig(path, &size, &mtime)) return;
  const std::string tmp = path + ".sqidx.tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  SidecarHeader h;
  memcpy(h.magic, kMagic, 8);
  h.src_size = size; h.src_mtime_ns = mtime; h.nk = idx.ks.size(); h.T = idx.names.size(); h.nmaps = idx.maps.size();
  h.names_bytes = 0;
  std::vector<uint32_t> name_len;
  for (auto& n : idx.names) { name_len.push_back((uint32_t)n.size()); h.names_bytes += n.size(); }
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  if (h.nk) ok &= fwrite(idx.ks.data(), 4, h.nk, f) == h.nk;
  if (h.T) ok &= fwrite(name_len.data(), 4, h.T, f) == h.T;
  for (auto& n : idx.names) ok &= fwrite(n.data(), 1, n.size(), f) == n.size();
  for (auto& kv : idx.maps) {
    const Postings& P = kv.second;
    uint64_t hdr[3] = {kv.first, P.keys.size(), P.tid.size()};
    ok &= fwrite(hdr, 8, 3, f) == 3;
    if (hdr[1]) ok &= fwrite(P.keys.data(), 4, hdr[1], f) == hdr[1];
    ok &= fwrite(P.off.data(), 8, hdr[1] + 1, f) == hdr[1] + 1;
    if (hdr[2]) ok &= fwrite(P.tid.data(), 4, hdr[2], f) == hdr[2];
  }
  fclose(f);
  if (ok) rename(tmp.c_str(), (path + ".sqidx").c_str()); else remove(tmp.c_str());
}
}  // namespace

bool read_index(const std::string& path, IndexData* out, bool keep_sequences, bool use_cache) {
  if (use_cache && !keep_sequences) {
    out->ks.clear(); out->names.clear(); out->sequences.clear(); out->maps.clear();
    if (load_sidecar(path, out)) {
      std::cout << "Index loaded from " << path << std::endl;  // data_io.cpp:303
      return true;
    }
  }
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
  if (use_cache && !keep_sequences) save_sidecar(path, *out);
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

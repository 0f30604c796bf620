// On-disk index of the reference (src/data_io.cpp:165-220 writer, :233-304 reader), field-compatible both ways:
//   u64 nK; u32 k[nK];
//   u64 T;  T x { u64 idLen; id; u64 seqLen; seq; i32 length }
//   u64 nMaps; nMaps x { u32 k; u64 nKeys; nKeys x { u32 hash; u64 deg; deg x { u64 tidLen; tid } } }
// native little-endian, no magic, any record order.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace sqhost {

struct Postings {
  std::vector<uint32_t> keys;
  std::vector<uint64_t> off;  // keys.size()+1
  std::vector<uint32_t> tid;  // dense transcript ids (position in IndexData::names)
};

struct IndexData {
  std::vector<uint32_t> ks;
  std::vector<std::string> names;      // transcript ids in file order (distinct)
  std::vector<std::string> sequences;  // only filled when keep_sequences
  std::unordered_map<uint32_t, Postings> maps;  // k -> inverted map (missing k: no map, sparse_chaining.cpp:51-53)
};

// false when the file cannot be opened (the reference prints an error and carries on, data_io.cpp:238-241)
// use_cache: read / write the optional sidecar <path>.sqidx (flat arrays, validated by size+mtime of <path>)
bool read_index(const std::string& path, IndexData* out, bool keep_sequences, bool use_cache = false);
bool write_index(const std::string& path, const std::vector<uint32_t>& ks, const std::vector<std::string>& names,
                 const std::vector<std::string>& sequences, const std::unordered_map<uint32_t, Postings>& maps);

}  // namespace sqhost

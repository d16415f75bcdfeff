// Host-side build of the BED name table (perfect hash, displacements, overflow list, word blob) of one side.
// Plain C++ so that the CPU test-suite can build tables without a GPU (jx_hostprobe.cpp) with the very code
// jx_bed_load runs.  The hash is jx_fast5.cuh's slot_hash_* / slot_index -- the same functions the kernels use.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jcvi_b200.h"
#include "jx_fast5.cuh"

namespace jxnt {
struct HostReader {
  const unsigned char* p;
  inline uint32_t operator()(int i) const { return p[i]; }
};
// Host tables of the last jx_bed_load (perfect hash, displacements, overflow lists, word blob): the same name
// list is often loaded on both sides (self comparisons; formats.blast cscore / filter --ids), and building the
// tables is O(names) hashing and sorting on one host thread.
struct BuiltSide {
  bool valid = false;
  std::vector<char> names; std::vector<uint32_t> off, name_id; std::vector<uint8_t> indexable; bool has_indexable = false;
  std::vector<uint32_t> tab; std::vector<uint16_t> disp; std::vector<uint32_t> ovf_hash, ovf_rank, words, meta;
  uint32_t bshift = 31, cshift = 28;
  bool matches(const jx_bed* bed) const {
    const size_t G = (size_t)bed->n_genes;
    if (!valid || off.size() != G + 1) return false;
    if (memcmp(off.data(), bed->name_off, (G + 1) * 4) != 0) return false;
    const size_t blob = G ? bed->name_off[G] : 0;
    if (names.size() != blob || (blob && memcmp(names.data(), bed->names, blob) != 0)) return false;
    if (G && memcmp(name_id.data(), bed->name_id, G * 4) != 0) return false;
    if (has_indexable != (bed->indexable != nullptr)) return false;
    if (has_indexable && G && memcmp(indexable.data(), bed->indexable, G) != 0) return false;
    return true;
  }
};

// returns "" or an error message
inline std::string build_side(const jx_bed* bed, BuiltSide& B) {
  const int64_t G = bed->n_genes;
  const size_t blob = G ? bed->name_off[G] : 0;
  {
    B.valid = false;
    HostReader rd{(const unsigned char*)bed->names};
    std::vector<std::pair<uint32_t, int64_t>> keys;     // (hash, rank); the last rank of a repeated name wins
    {
      std::unordered_map<std::string, int64_t> uniq;
      for (int64_t r = 0; r < G; ++r) {
        if (bed->indexable && !bed->indexable[r]) continue;
        const uint32_t a = bed->name_off[r], b = bed->name_off[r + 1];
        if (b - a >= 128u) continue;   // can never match: BLAST names of 128+ bytes are rejected
        uniq[std::string(bed->names + a, bed->names + b)] = r;
      }
      keys.reserve(uniq.size());
      for (const auto& kv : uniq) {
        const int64_t r = kv.second;
        keys.emplace_back(jx::slot_hash_bytes(rd, (int)bed->name_off[r], (int)bed->name_off[r + 1]), r);
      }
    }
    // names with identical 32-bit hashes cannot be separated by any displacement (expected from ~10^5
    // names on): the first keeps its place in the table, the others go to a small sorted overflow
    // list that lookups consult only after a miss
    std::sort(keys.begin(), keys.end());
    std::vector<uint32_t> ovf_hash, ovf_rank;
    {
      size_t w = 0;
      for (size_t k = 0; k < keys.size(); ++k) {
        if (w && keys[w - 1].first == keys[k].first) { ovf_hash.push_back(keys[k].first); ovf_rank.push_back((uint32_t)keys[k].second); }
        else keys[w++] = keys[k];
      }
      keys.resize(w);
    }
    const size_t n = keys.size();
    uint32_t cbits = 4;
    while ((1ull << cbits) < 2 * (uint64_t)n + 2) ++cbits;
    std::vector<uint32_t>& tab = B.tab;
    std::vector<uint16_t>& disp = B.disp;
    uint32_t bbits = 1;
    for (;; ++cbits) {
      if (cbits > 30) return "BED name table does not fit";
      bbits = 1;
      while ((1ull << bbits) < (uint64_t)n / 8 + 1 && bbits < 28) ++bbits;   // ~8 names per bucket: the displacement table stays L1-resident
      const uint32_t nb = 1u << bbits, cap = 1u << cbits, cshift = 32 - cbits, bshift = 32 - bbits;
      std::vector<std::vector<uint32_t>> bucket(nb);
      for (uint32_t k = 0; k < n; ++k) bucket[keys[k].first >> bshift].push_back(k);
      std::vector<uint32_t> order(nb);
      for (uint32_t i = 0; i < nb; ++i) order[i] = i;
      std::stable_sort(order.begin(), order.end(),
                       [&](uint32_t x, uint32_t y) { return bucket[x].size() > bucket[y].size(); });
      std::vector<int64_t> owner((size_t)cap, -1);
      disp.assign(nb, 0);
      bool ok = true;
      std::vector<uint32_t> pos;
      for (uint32_t bi : order) {
        const auto& ks = bucket[bi];
        if (ks.empty()) break;
        uint32_t d = 0;
        for (; d < 65536u; ++d) {
          pos.clear();
          bool free_all = true;
          for (uint32_t k : ks) {
            const uint32_t p = jx::slot_index(keys[k].first, d, cshift);
            if (owner[p] >= 0 || std::find(pos.begin(), pos.end(), p) != pos.end()) { free_all = false; break; }
            pos.push_back(p);
          }
          if (free_all) break;
        }
        if (d == 65536u) { ok = false; break; }       // (two names with the same 32-bit hash, or bad luck): grow
        disp[bi] = (uint16_t)d;
        for (size_t i = 0; i < ks.size(); ++i) owner[pos[i]] = keys[ks[i]].second;
      }
      if (!ok) continue;
      tab.assign((size_t)cap * 8, 0u);
      for (uint32_t p = 0; p < cap; ++p) {
        const int64_t r = owner[p];
        if (r < 0) continue;
        uint32_t* slot = &tab[(size_t)p * 8];
        const uint32_t a = bed->name_off[r], b = bed->name_off[r + 1];
        slot[0] = (uint32_t)(r + 1);
        slot[1] = bed->name_id[r];
        for (uint32_t i = 0; i < 4 * JX_SLOT_WORDS && a + i < b; ++i)
          slot[2 + (i >> 2)] |= (uint32_t)(unsigned char)bed->names[a + i] << (8 * (i & 3));
      }
      B.bshift = bshift; B.cshift = cshift;
      break;
    }
    B.ovf_hash.swap(ovf_hash); B.ovf_rank.swap(ovf_rank);

    // word-aligned, zero-padded copy of the names so that the kernel verifies a probe with
    // 32-bit compares against shared-memory words
    std::vector<uint32_t>& meta = B.meta; std::vector<uint32_t>& words = B.words;
    meta.assign((size_t)G, 0u); words.clear();
    words.reserve(blob / 4 + (size_t)G + 4);
    for (int64_t r = 0; r < G; ++r) {
      uint32_t a = bed->name_off[r], b = bed->name_off[r + 1], len = b - a;
      if (len >= 128u) { meta[(size_t)r] = 0; continue; }
      if (words.size() >= (1u << 25)) return "BED names exceed 128 MiB";
      meta[(size_t)r] = ((uint32_t)words.size() << 7) | len;
      for (uint32_t i = 0; i < len; i += 4) {
        uint32_t w = 0;
        for (uint32_t k = 0; k < 4 && i + k < len; ++k) w |= (uint32_t)(unsigned char)bed->names[a + i + k] << (8 * k);
        words.push_back(w);
      }
    }
    words.push_back(0);
    B.names.assign(bed->names, bed->names + blob);
    B.off.assign(bed->name_off, bed->name_off + G + 1);
    B.name_id.assign(bed->name_id, bed->name_id + G);
    B.has_indexable = bed->indexable != nullptr;
    if (B.has_indexable) B.indexable.assign(bed->indexable, bed->indexable + G); else B.indexable.clear();
    B.valid = true;
    }
  return "";
}
}  // namespace jxnt

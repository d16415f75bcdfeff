// Context, BED side tables and the sort/scan library primitives of libjcvi_b200.so.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include "jx_internal.cuh"
#include "jx_fast5.cuh"

namespace {
struct HostReader {
  const unsigned char* p;
  __host__ __device__ inline uint32_t operator()(int i) const { return p[i]; }
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
BuiltSide g_built;          // guarded by g_built_mu: contexts of different threads may load concurrently
std::mutex g_built_mu;

template <class T>
int upload(jx_ctx* ctx, T** dst, const T* src, size_t n) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  JX_CK(cudaMalloc((void**)dst, (n ? n : 1) * sizeof(T)));
  if (n) JX_CK(cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  return JX_OK;
}
}  // namespace

extern "C" {

void jx_free(void* p) { free(p); }

void jx_abi_sizes(int32_t out[6]) {
  out[0] = (int32_t)sizeof(jx_bed); out[1] = (int32_t)sizeof(jx_filter_opts); out[2] = (int32_t)sizeof(jx_filter_result);
  out[3] = (int32_t)sizeof(jx_scan_opts); out[4] = (int32_t)sizeof(jx_scan_result); out[5] = (int32_t)sizeof(jx_liftover_result);
}

const char* jx_version(void) { return "jcvi_b200 0.1 (sm_100a)"; }

int jx_ctx_create(int device, jx_ctx** out) {
  if (!out) return JX_EARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return JX_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return JX_ECUDA;
  jx_ctx* ctx = new jx_ctx();
  ctx->device = device;
  ctx->trace = getenv("JX_TRACE") ? atoi(getenv("JX_TRACE")) : 0;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return JX_ECUDA;
  }
  // keep freed scratch memory in the stream-ordered pool between calls
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = ctx;
  return JX_OK;
}

void jx_ctx_destroy(jx_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int s = 0; s < 2; ++s) {
    SideDev& d = ctx->side[s];
    cudaFree(d.table); cudaFree(d.disp); cudaFree(d.ovf_hash); cudaFree(d.ovf_rank); cudaFree(d.blobw); cudaFree(d.name_meta); cudaFree(d.chr_lex); cudaFree(d.chr_begin);
    cudaFree(d.chr_end); cudaFree(d.length); cudaFree(d.lexrank); cudaFree(d.name_id);
  }
  cudaFree(ctx->q_same_s);
  cudaStreamSynchronize(ctx->copy_stream);
  for (cudaEvent_t ev : ctx->held_events) cudaEventDestroy(ev);
  cudaFree(ctx->held_text);
  cudaFree(ctx->lines_dev);
  for (auto& b : ctx->pin_pool) cudaFreeHost(b.p);
  if (ctx->pin_small) cudaFreeHost(ctx->pin_small);
  if (ctx->ws) cudaFree(ctx->ws);
  cudaFree(ctx->cache.recs); cudaFree(ctx->cache.desc); cudaFree(ctx->sess_best);
  cudaFree(ctx->lastf.q); cudaFree(ctx->lastf.s); cudaFree(ctx->lastf.score);
  for (auto& ev : ctx->tok_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

const char* jx_last_error(jx_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t jx_launch_count(jx_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t jx_stream_handle(jx_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

int jx_tokenizer_timing(jx_ctx* ctx, int reset, double* ms, int64_t* launches, int64_t* bytes, int64_t* lines) {
  if (!ctx) return JX_EARG;
  cudaSetDevice(ctx->device);
  JX_CK(cudaStreamSynchronize(ctx->stream));
  jx_harvest_tok_events(ctx);
  if (ms) *ms = ctx->tok_ms_acc;
  if (launches) *launches = ctx->tok_launches_acc;
  if (bytes) *bytes = ctx->tok_bytes;
  if (lines) *lines = ctx->tok_lines;
  if (reset) { ctx->tok_ms_acc = 0; ctx->tok_launches_acc = 0; ctx->tok_bytes = 0; ctx->tok_lines = 0; }
  return JX_OK;
}

int jx_bed_load(jx_ctx* ctx, int side, const jx_bed* bed) {
  if (!ctx || !bed || side < 0 || side > 1 || bed->n_genes < 0) return JX_EARG;
  if (bed->n_genes >= (1ll << 31) - 2) { ctx->err = "too many genes"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  SideDev& d = ctx->side[side];
  const int64_t G = bed->n_genes;
  d.G = G;
  d.n_chr = bed->n_chr;
  const size_t blob = G ? bed->name_off[G] : 0;
  d.h_names.assign(bed->names, bed->names + blob);
  d.h_off.assign(bed->name_off, bed->name_off + G + 1);
  d.h_chr_lex.assign(bed->chr_lex, bed->chr_lex + G);
  // perfect hash table of 32-byte slots at load factor <= 0.5 (hash-and-displace), built on the host
  // with the same hash the kernels use (jx_fast5.cuh slot_hash_* / slot_index)
  int rc;
  std::lock_guard<std::mutex> built_lock(g_built_mu);
  BuiltSide& B = g_built;
  const bool hit = B.matches(bed);
  if (!hit) {
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
      if (cbits > 30) { ctx->err = "BED name table does not fit"; return JX_EARG; }
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
      if (words.size() >= (1u << 25)) { ctx->err = "BED names exceed 128 MiB"; return JX_EARG; }
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
  d.bshift = B.bshift; d.cshift = B.cshift;
  if ((rc = upload(ctx, &d.table, reinterpret_cast<const uint4*>(B.tab.data()), B.tab.size() / 4))) return rc;
  if ((rc = upload(ctx, &d.disp, B.disp.data(), B.disp.size()))) return rc;
  d.n_ovf = (uint32_t)B.ovf_hash.size();
  if ((rc = upload(ctx, &d.ovf_hash, B.ovf_hash.data(), B.ovf_hash.size()))) return rc;
  if ((rc = upload(ctx, &d.ovf_rank, B.ovf_rank.data(), B.ovf_rank.size()))) return rc;
  if ((rc = upload(ctx, &d.blobw, B.words.data(), B.words.size()))) return rc;
  if ((rc = upload(ctx, &d.name_meta, B.meta.data(), B.meta.size()))) return rc;
  if ((rc = upload(ctx, &d.chr_lex, bed->chr_lex, (size_t)G))) return rc;
  if ((rc = upload(ctx, &d.chr_begin, bed->chr_begin, (size_t)G))) return rc;
  if ((rc = upload(ctx, &d.chr_end, bed->chr_end, (size_t)G))) return rc;
  if ((rc = upload(ctx, &d.length, bed->length, (size_t)G))) return rc;
  if ((rc = upload(ctx, &d.lexrank, bed->accn_lexrank, (size_t)G))) return rc;
  if ((rc = upload(ctx, &d.name_id, bed->name_id, (size_t)G))) return rc;
  JX_CK(cudaStreamSynchronize(ctx->stream));
  d.loaded = true;
  ctx->pair_ready = false;
  ctx->cache.valid = false;
  return JX_OK;
}

int jx_pair_setup(jx_ctx* ctx, int is_self, const int32_t* q_same_s, uint32_t n_name_ids) {
  if (!ctx || !q_same_s) return JX_EARG;
  if (!ctx->side[0].loaded || !ctx->side[1].loaded) { ctx->err = "jx_bed_load both sides first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  ctx->is_self = is_self != 0;
  ctx->n_name_ids = n_name_ids;
  int rc;
  if ((rc = upload(ctx, &ctx->q_same_s, q_same_s, (size_t)ctx->side[0].G))) return rc;
  JX_CK(cudaStreamSynchronize(ctx->stream));
  ctx->pair_ready = true;
  return JX_OK;
}

}  // extern "C"

// ---- CUB-backed primitives -------------------------------------------------------------------
int jx_sort_pairs(jx_ctx* ctx, Arena& arena, uint64_t*& keys, uint32_t*& vals, size_t n, int begin_bit, int end_bit) {
  if (n == 0) return JX_OK;
  if (n >= (1ull << 31)) { ctx->err = "sort of >= 2^31 items"; return JX_EARG; }
  JX_ALLOC(k2, uint64_t, n, false);
  JX_ALLOC(v2, uint32_t, n, false);
  cub::DoubleBuffer<uint64_t> dk(keys, k2);
  cub::DoubleBuffer<uint32_t> dv(vals, v2);
  size_t tmp = 0;
  JX_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n, begin_bit, end_bit, ctx->stream));
  JX_ALLOC(t, uint8_t, tmp, false);
  JX_CK(cub::DeviceRadixSort::SortPairs(t, tmp, dk, dv, (int)n, begin_bit, end_bit, ctx->stream));
  ++ctx->launches;
  keys = dk.Current();
  vals = dv.Current();
  return JX_OK;
}

int jx_sort_keys(jx_ctx* ctx, Arena& arena, uint64_t*& keys, size_t n, int begin_bit, int end_bit) {
  if (n == 0) return JX_OK;
  if (n >= (1ull << 31)) { ctx->err = "sort of >= 2^31 items"; return JX_EARG; }
  JX_ALLOC(k2, uint64_t, n, false);
  cub::DoubleBuffer<uint64_t> dk(keys, k2);
  size_t tmp = 0;
  JX_CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, dk, (int)n, begin_bit, end_bit, ctx->stream));
  JX_ALLOC(t, uint8_t, tmp, false);
  JX_CK(cub::DeviceRadixSort::SortKeys(t, tmp, dk, (int)n, begin_bit, end_bit, ctx->stream));
  ++ctx->launches;
  keys = dk.Current();
  return JX_OK;
}

int jx_exclusive_sum(jx_ctx* ctx, Arena& arena, const uint32_t* in, uint32_t* out, size_t n) {
  if (n == 0) return JX_OK;
  size_t tmp = 0;
  JX_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream));
  JX_ALLOC(t, uint8_t, tmp, false);
  JX_CK(cub::DeviceScan::ExclusiveSum(t, tmp, in, out, (int)n, ctx->stream));
  ++ctx->launches;
  return JX_OK;
}

int jx_exclusive_sum64(jx_ctx* ctx, Arena& arena, const uint32_t* in, uint64_t* out, size_t n) {
  if (n == 0) return JX_OK;
  size_t tmp = 0;
  JX_CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, in, out, cub::Sum(), (uint64_t)0, (int)n, ctx->stream));
  JX_ALLOC(t, uint8_t, tmp, false);
  JX_CK(cub::DeviceScan::ExclusiveScan(t, tmp, in, out, cub::Sum(), (uint64_t)0, (int)n, ctx->stream));
  ++ctx->launches;
  return JX_OK;
}

// Large device -> pageable host copy: 16 MiB chunks through two pinned staging halves; while chunk c + 1 crosses
// PCIe, chunk c is copied from staging into `host` by a few host threads (a plain cudaMemcpy to pageable memory
// runs at a fraction of the link rate and first-touches every page on one thread).
int jx_d2h_big(jx_ctx* ctx, char* host, const uint8_t* dev, size_t bytes) {
  constexpr size_t CH = 16u << 20;
  if (bytes <= (1u << 20)) {
    JX_CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    return JX_OK;
  }
  char* stage = nullptr;
  int rc;
  if ((rc = jx_pin_small(ctx, 2 * CH, &stage))) return rc;
  cudaEvent_t ev[2];
  JX_CK(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  JX_CK(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  const size_t nch = (bytes + CH - 1) / CH;
  const unsigned hw = std::thread::hardware_concurrency();
  const unsigned nt = std::max(1u, std::min(4u, hw ? hw : 1u));
  auto drain = [&](size_t c) {                     // staging half of chunk c -> host
    const size_t o = c * CH, len = std::min(CH, bytes - o);
    const char* src = stage + (c & 1) * CH;
    if (nt == 1 || len < (4u << 20)) { memcpy(host + o, src, len); return; }
    std::vector<std::thread> th;
    const size_t part = ((len + nt - 1) / nt + 4095) & ~(size_t)4095;
    size_t done = 0;                                 // bytes handed to a thread so far
    try {
      for (unsigned t = 0; t < nt; ++t) {
        const size_t a = (size_t)t * part;
        if (a >= len) break;
        const size_t l = std::min(part, len - a);
        th.emplace_back([=]() { memcpy(host + o + a, src + a, l); });
        done = a + l;
      }
    } catch (...) {}                                 // no more threads: this one copies the rest
    if (done < len) memcpy(host + o + done, src + done, len - done);
    for (auto& x : th) x.join();
  };
  cudaError_t err = cudaSuccess;
  for (size_t c = 0; c < nch && err == cudaSuccess; ++c) {
    const size_t o = c * CH, len = std::min(CH, bytes - o);
    if (c >= 2) { /* the half is free: chunk c - 2 was drained before chunk c - 1 was waited for */ }
    err = cudaMemcpyAsync(stage + (c & 1) * CH, dev + o, len, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaEventRecord(ev[c & 1], ctx->stream);
    if (c >= 1 && err == cudaSuccess) {
      err = cudaEventSynchronize(ev[(c - 1) & 1]);
      if (err == cudaSuccess) drain(c - 1);
    }
  }
  if (err == cudaSuccess) err = cudaEventSynchronize(ev[(nch - 1) & 1]);
  if (err == cudaSuccess) drain(nch - 1);
  cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
  if (err != cudaSuccess) { ctx->err = std::string("jx_d2h_big: ") + cudaGetErrorString(err); return JX_ECUDA; }
  return JX_OK;
}

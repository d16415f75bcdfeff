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
#include "jx_nametable.h"

namespace {
using jxnt::BuiltSide;
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

int jx_host_query_runs(const char* text, int64_t nbytes) {
  return (text && nbytes > 0) ? jx_sample_query_runs((const uint8_t*)text, (size_t)nbytes) : 0;
}

int32_t jx_abi_size_synfind(void) { return (int32_t)sizeof(jx_synfind_result); }

const char* jx_version(void) { return "jcvi_b200 0.2 (sm_100a)"; }

int jx_ctx_create(int device, jx_ctx** out) {
  if (!out) return JX_EARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return JX_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return JX_ECUDA;
  jx_ctx* ctx = new jx_ctx();
  ctx->device = device;
  ctx->trace = getenv("JX_TRACE") ? atoi(getenv("JX_TRACE")) : 0;
  ctx->tok_v5 = getenv("JX_TOK_V5") && atoi(getenv("JX_TOK_V5")) != 0;
  ctx->tok_v6 = getenv("JX_TOK_V6") && atoi(getenv("JX_TOK_V6")) != 0;
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
  if (ctx->io_pin) cudaFreeHost(ctx->io_pin);
  cudaFree(ctx->ftext_dev);
  if (ctx->ev_text) cudaEventDestroy(ctx->ev_text);
  if (ctx->pin_scan) cudaFreeHost(ctx->pin_scan);
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
    const std::string e = jxnt::build_side(bed, B);
    if (!e.empty()) { ctx->err = e; return JX_EARG; }
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

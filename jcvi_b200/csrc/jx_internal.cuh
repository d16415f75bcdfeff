// Internal declarations shared by the translation units of libjcvi_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/jcvi_b200.h"
#include "jx_parse.cuh"

// ---- records ---------------------------------------------------------------------------------
// One 16-byte record per input line, indexed by the line's ordinal in the text (so the record
// index IS the file order the reference's first-seen / stable-sort rules refer to).
//   x = query rank  (0xFFFFFFFF: no record -- '#' line, unmapped name, dropped self hit ...)
//   y = subject rank
//   z = float32 score bits (cblast.pyx:88)
//   w = (byte offset of the line inside its tile (JX_TILE bytes) << 1) | (evalue < 1e-10)
#define JX_NOREC 0xFFFFFFFFu

constexpr int JX_TOK_THREADS = 128;                      // threads per CTA of the tokenizer; they cooperate on one tile
#define JX_GROUP 128
constexpr int JX_SEG = 64;                               // bytes of the staged window classified by one thread
// A CTA stages a WINDOW of 8 KiB = 128 threads x 64 bytes: the tile (the lines that START in it are the CTA's)
// plus 256 bytes of look-ahead, so that every thread classifies exactly one segment and a line that starts near
// the end of the tile is still inside the window (longer lines fall back to reading global memory).
constexpr int JX_OVER = 256;
constexpr int JX_TILE = JX_TOK_THREADS * JX_SEG - JX_OVER;   // 7936 (a multiple of 16: bulk-copy source alignment)
constexpr int JX_WIN = JX_TILE + JX_OVER;                // 8192

struct SideDev {
  int64_t G = 0;
  int32_t n_chr = 0;
  uint4* table = nullptr;      // perfect hash, two uint4 per 32-byte slot: {rank + 1 (0 = empty), name id, name[24]}
  uint16_t* disp = nullptr;    // per bucket displacement (hash-and-displace)
  uint32_t bshift = 31, cshift = 28;   // bucket = hash >> bshift, slot = slot_index(hash, disp[bucket], cshift)
  uint32_t* ovf_hash = nullptr;  // names whose 32-bit hash equals that of a name in the table (sorted by hash; rare)
  uint32_t* ovf_rank = nullptr;
  uint32_t n_ovf = 0;
  uint32_t* blobw = nullptr;     // accn bytes, each name padded with zeros to a 4-byte boundary
  uint32_t* name_meta = nullptr; // per rank: (word offset into blobw << 7) | byte length (< 128)
  int32_t *chr_lex = nullptr, *chr_begin = nullptr, *chr_end = nullptr, *length = nullptr;
  uint32_t *lexrank = nullptr, *name_id = nullptr;
  std::vector<char> h_names;
  std::vector<uint32_t> h_off;
  std::vector<int32_t> h_chr_lex;
  bool loaded = false;
};

struct jx_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  SideDev side[2];
  bool is_self = false;
  bool pair_ready = false;
  int32_t* q_same_s = nullptr;
  uint32_t n_name_ids = 0;
  std::string err;
  int trace = 0;                 // JX_TRACE=1: print phase timings (synchronising; debugging aid)
  double trace_t0 = 0;
  int64_t launches = 0;
  // tokenizer timing (CUDA events on `stream`)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tok_events;
  bool record_reuse = true;      // jx_set_record_reuse
  // per-call scratch: one grow-only device workspace handed out by bump pointer (Arena); requests that
  // do not fit fall back to cudaMallocAsync and make the workspace grow for the next call
  char* ws = nullptr;
  size_t ws_cap = 0, ws_used = 0, ws_virtual = 0, ws_need = 0;
  int ws_depth = 0;
  int64_t tok_resident = 0;      // persistent tokenizer CTAs that fit on this device at once
  bool tok_v5 = false, tok_v6 = false;   // JX_TOK_V5=1 / JX_TOK_V6=1: run the path through an older kernel (A/B measurements)
  uint32_t tok_cap = 64;         // version-7 kernel: record slots per sub-tile that fitted the last text (sticky guess)
  int tok_grouped = 0;           // the held text's lines come in runs of equal queries (sampled at upload)
  int64_t tok_bytes = 0;
  // text uploaded with jx_text_upload (device copy + one event per 64 MiB chunk)
  // records of the last jx_blastfilter pass over the held text, for jx_liftover_text
  struct RecCache {
    uint4* recs = nullptr; size_t recs_cap = 0;       // context-owned, grow-only
    uint64_t* desc = nullptr; size_t desc_cap = 0;
    uint32_t cap = 0, sub_bytes = 0;                  // record slot layout (see LineIndex)
    bool valid = false;
    bool plain = false;       // records of a non-self pass without exclusions or the near-diagonal rule: any stage may reuse them
    const uint8_t* text = nullptr; int64_t nbytes = 0; int strip = 0;
    int64_t ntiles = 0; uint32_t n_lines = 0;
    unsigned long long stats[12] = {0};
  } cache;
  // device-resident survivors of the last jx_blastfilter (for jx_scan_filtered)
  struct LastFilter { int32_t* q = nullptr; int32_t* s = nullptr; float* score = nullptr; size_t cap = 0; int64_t n = 0; int64_t token = 0; } lastf;
  int64_t token_counter = 0;
  uint32_t* sess_best = nullptr; size_t sess_best_cap = 0;   // jx_cscore_begin/finish session
  // one genome pair sharded over ranks (jx_shard_*): device state that lives from one stage call to the next
  struct Shard {
    std::vector<void*> allocs;                         // stream-ordered allocations of this session
    const uint4* C = nullptr; uint32_t nC = 0;         // owner side: the candidates received for this rank's chromosome pairs
    uint32_t* cnt = nullptr;                           // device counters
    unsigned long long *K1 = nullptr, *V1 = nullptr; uint32_t* W1 = nullptr; uint32_t hmask = 0;
    uint32_t* Cslot = nullptr; uint8_t* win = nullptr;
    int32_t *mq = nullptr, *ms = nullptr;              // mother maps (identical on every rank)
  } sh;
  bool sess_open = false;
  uint8_t* lines_dev = nullptr;  // grow-only device buffer of the last jx_*_device result (packed raw lines)
  size_t lines_dev_cap = 0;
  // pinned buffers for `.filtered` texts: a result borrows one and gives it back in jx_filter_result_free, so
  // every result owns its text (no staleness when a later call runs) and steady state allocates nothing
  struct PinBuf { char* p; size_t cap; bool busy; };
  std::vector<PinBuf> pin_pool;
  char* pin_small = nullptr;     // grow-only pinned staging for small result arrays (one D2H instead of several pageable ones)
  size_t pin_small_cap = 0;
  char* pin_scan = nullptr;      // the same for scan results (their copy may still be in flight while liftover stages its own)
  size_t pin_scan_cap = 0;
  // jx_pipeline: the `.filtered` text is formatted into a context-owned device buffer and copied to the host on the copy
  // stream while scan and liftover run (the stage's own scratch is recycled by then)
  uint8_t* ftext_dev = nullptr; size_t ftext_dev_cap = 0; bool defer_text = false, text_in_flight = false;
  cudaEvent_t ev_text = nullptr;
  void* io_pin = nullptr;        // jx_file_upload: ring of pinned pieces (jx_io.cu), grow-never, freed with the context
  uint8_t* held_text = nullptr;
  int64_t held_n = 0;
  int64_t held_cap = 0;
  std::vector<cudaEvent_t> held_events;
  int64_t tok_lines = 0;
  double tok_ms_acc = 0;
  int64_t tok_launches_acc = 0;
};

#define JX_CK(call)                                                                         \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) {                                                                \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
      return JX_ECUDA;                                                                      \
    }                                                                                       \
  } while (0)

#include <chrono>
inline void jx_trace_point(jx_ctx* ctx, const char* label) {
  if (!ctx->trace) return;
  cudaStreamSynchronize(ctx->stream);
  const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  fprintf(stderr, "[jx trace] %-28s +%.3f ms\n", label, ctx->trace_t0 ? t - ctx->trace_t0 : 0.0);
  ctx->trace_t0 = t;
}
#define JX_TRACE(label) jx_trace_point(ctx, label)

// Do the lines of a table come in runs of equal queries (lastal / BLAST write all alignments of a query together)?
// Looks at the first 256 KiB: 1 when at least half of the lines repeat the first field of the line before them.
inline int jx_sample_query_runs(const uint8_t* p, size_t n) {
  if (n > (256u << 10)) n = 256u << 10;
  size_t lines = 0, same = 0, prev_a = 0, prev_len = 0, i = 0;
  bool have_prev = false;
  while (i < n) {
    size_t a = i, e = i;
    while (e < n && p[e] != '\t' && p[e] != '\n') ++e;
    const size_t len = e - a;
    if (len && p[a] != '#') {
      ++lines;
      if (have_prev && len == prev_len && memcmp(p + a, p + prev_a, len) == 0) ++same;
      prev_a = a; prev_len = len; have_prev = true;
    }
    while (e < n && p[e] != '\n') ++e;
    i = e + 1;
  }
  return lines >= 32 && same * 2 >= lines ? 1 : 0;
}

// fold finished tokenizer launch timings into the accumulators (call after a stream synchronise);
// keeps the event list from growing when nobody asks for jx_tokenizer_timing
inline void jx_harvest_tok_events(jx_ctx* ctx) {
  for (auto& ev : ctx->tok_events) {
    float t = 0;
    if (cudaEventElapsedTime(&t, ev.first, ev.second) == cudaSuccess) { ctx->tok_ms_acc += t; ++ctx->tok_launches_acc; }
    cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
  }
  ctx->tok_events.clear();
}

inline char* jx_pin_acquire(jx_ctx* ctx, size_t bytes) {
  int best = -1, spare = -1;
  for (int i = 0; i < (int)ctx->pin_pool.size(); ++i) {
    auto& b = ctx->pin_pool[(size_t)i];
    if (b.busy) continue;
    if (b.cap >= bytes) { if (best < 0 || b.cap < ctx->pin_pool[(size_t)best].cap) best = i; }
    else spare = i;
  }
  if (best >= 0) { ctx->pin_pool[(size_t)best].busy = true; return ctx->pin_pool[(size_t)best].p; }
  const size_t cap = bytes + (bytes >> 2) + (1u << 20);
  char* p = nullptr;
  if (spare >= 0) {                                 // a free buffer that is too small: replace it
    cudaFreeHost(ctx->pin_pool[(size_t)spare].p);
    ctx->pin_pool.erase(ctx->pin_pool.begin() + spare);
  }
  if (cudaMallocHost((void**)&p, cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ctx->pin_pool.push_back({p, cap, true});
  return p;
}
inline void jx_pin_release(jx_ctx* ctx, char* p) {
  for (auto& b : ctx->pin_pool) if (b.p == p) { b.busy = false; return; }
}

// pinned staging of at least `bytes` (valid until the next call of this function on the context)
inline int jx_pin_small(jx_ctx* ctx, size_t bytes, char** out) {
  if (ctx->pin_small_cap < bytes) {
    if (ctx->pin_small) cudaFreeHost(ctx->pin_small);
    ctx->pin_small = nullptr; ctx->pin_small_cap = 0;
    const size_t cap = bytes + bytes / 2 + 4096;
    if (cudaMallocHost((void**)&ctx->pin_small, cap) != cudaSuccess) { ctx->err = "cudaMallocHost failed"; return JX_ECUDA; }
    ctx->pin_small_cap = cap;
  }
  *out = ctx->pin_small;
  return JX_OK;
}

inline int jx_pin_scan(jx_ctx* ctx, size_t bytes, char** out) {
  if (ctx->pin_scan_cap < bytes) {
    if (ctx->pin_scan) { cudaStreamSynchronize(ctx->copy_stream); cudaFreeHost(ctx->pin_scan); }
    ctx->pin_scan = nullptr; ctx->pin_scan_cap = 0;
    const size_t cap = bytes + bytes / 2 + 4096;
    if (cudaMallocHost((void**)&ctx->pin_scan, cap) != cudaSuccess) { ctx->err = "cudaMallocHost failed"; return JX_ECUDA; }
    ctx->pin_scan_cap = cap;
  }
  *out = ctx->pin_scan;
  return JX_OK;
}

// session memory of the sharded stages: stream-ordered, released by jx_shard_reset / the next session
template <class T>
inline T* jx_shard_alloc(jx_ctx* ctx, size_t n, bool zero = false) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
  if (cudaMallocAsync(&p, bytes, ctx->stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ctx->sh.allocs.push_back(p);
  if (zero) cudaMemsetAsync(p, 0, bytes, ctx->stream);
  return (T*)p;
}
#define JX_SALLOC(var, T, n, zero)                                                          \
  T* var = jx_shard_alloc<T>(ctx, (n), (zero));                                             \
  if (!var) {                                                                               \
    ctx->err = "device allocation failed: " #var;                                           \
    return JX_ECUDA;                                                                        \
  }

#define JX_LAUNCH(kernel, grid, block, smem, ...)                                           \
  do {                                                                                      \
    kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);                          \
    ++ctx->launches;                                                                        \
  } while (0)

// stream-ordered scratch allocations, released when the Arena goes out of scope
struct Arena {
  jx_ctx* ctx;
  std::vector<void*> ptrs;       // fall-back allocations (cudaMallocAsync)
  size_t mark, vmark;
  explicit Arena(jx_ctx* c) : ctx(c), mark(c->ws_used), vmark(c->ws_virtual) { ++c->ws_depth; }
  ~Arena() {
    for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
    ctx->ws_used = mark;
    ctx->ws_virtual = vmark;
    if (--ctx->ws_depth == 0 && ctx->ws_need > ctx->ws_cap) {
      // grow to what this call would have needed (bounded: a quarter of the device); every kernel that
      // used the old block was enqueued on ctx->stream
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      size_t want = ctx->ws_need + ctx->ws_need / 4 + (64u << 20);
      if (want > total_b / 4) want = total_b / 4;
      if (want > ctx->ws_cap && want - ctx->ws_cap < free_b / 2) {
        cudaStreamSynchronize(ctx->stream);
        if (ctx->ws) cudaFree(ctx->ws);
        ctx->ws = nullptr; ctx->ws_cap = 0;
        if (cudaMalloc((void**)&ctx->ws, want) == cudaSuccess) ctx->ws_cap = want;
        else cudaGetLastError();
      }
      ctx->ws_need = 0;
    }
  }
  template <class T>
  T* alloc(size_t n, bool zero = false) {
    void* p = nullptr;
    const size_t bytes = (((n ? n : 1) * sizeof(T)) + 255) & ~(size_t)255;
    ctx->ws_virtual += bytes;
    if (ctx->ws_virtual > ctx->ws_need) ctx->ws_need = ctx->ws_virtual;
    if (ctx->ws && ctx->ws_used + bytes <= ctx->ws_cap) {
      p = ctx->ws + ctx->ws_used;
      ctx->ws_used += bytes;
    } else {
      if (cudaMallocAsync(&p, bytes, ctx->stream) != cudaSuccess) return nullptr;
      ptrs.push_back(p);
    }
    if (zero) cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream);
    return (T*)p;
  }
};

#define JX_ALLOC(var, T, n, zero)                                                           \
  T* var = arena.alloc<T>((n), (zero));                                                     \
  if (!var) {                                                                               \
    ctx->err = "device allocation failed: " #var;                                           \
    return JX_ECUDA;                                                                        \
  }

// ---- library primitives (jx_core.cu; CUB) ---------------------------------------------------
// LSD radix sort of (key, value) pairs on bits [begin_bit, end_bit); stable.  keys/vals are
// updated to point at the sorted buffers (which live in the arena).
int jx_sort_pairs(jx_ctx* ctx, Arena& arena, uint64_t*& keys, uint32_t*& vals, size_t n, int begin_bit, int end_bit);
int jx_sort_keys(jx_ctx* ctx, Arena& arena, uint64_t*& keys, size_t n, int begin_bit, int end_bit);
int jx_exclusive_sum(jx_ctx* ctx, Arena& arena, const uint32_t* in, uint32_t* out, size_t n);
int jx_exclusive_sum64(jx_ctx* ctx, Arena& arena, const uint32_t* in, uint64_t* out, size_t n);
// device -> pageable host, pipelined through pinned staging (stream-synchronous on return)
int jx_d2h_big(jx_ctx* ctx, char* host, const uint8_t* dev, size_t bytes);
inline int jx_bits(uint64_t maxval) { int b = 1; while (b < 64 && (maxval >> b)) ++b; return b; }

// ---- tokenizer (jx_tokenize.cu) -----------------------------------------------------------------
struct TokMode {
  int strip = 1;
  int neardiag = 0;            // self comparison: drop same-seqid hits with si - qi < 40 (synteny.py:280)
  uint32_t* best = nullptr;    // per name-id best score bits (filter_cscore pass 1) or null
  const uint64_t* excl = nullptr;
  int64_t n_excl = 0;
  bool fill_cache = false;     // keep records/descriptors in ctx->cache (held text only)
  bool persist = false;        // like fill_cache but also for self comparisons / exclusions (sharded session)
  bool names = false;          // no BED: emit a 64-bit hash of both (stripped) names per line instead of ranks
  // formats.blast filter (blast.py:620-709): keep / drop every line by a predicate on its fields
  bool filter = false;
  double f_score = 0, f_pctid = 0, f_evalue = 0;
  long long f_hitlen = 0;
  bool f_noself = false, f_inverse = false, f_ids = false;
  uint32_t seed = 0;           // of that hash
};
// How a record's index maps to its line's byte offset.  Version-7 records live in fixed slots (`cap` per sub-tile of
// `sub_bytes` bytes): offset = index / cap * sub_bytes + (rec.w >> 1).  cap == 0: dense line ordinals, the per-tile
// inclusive line counts (`desc`, tiles of JX_TILE bytes) are searched.
struct LineIndex {
  const uint64_t* desc = nullptr;
  int64_t ntiles = 0;
  uint32_t cap = 0, sub_bytes = 0;
};
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t jx_line_offset(const LineIndex& L, uint64_t i, uint32_t w) {
  if (L.cap) return (i / L.cap) * (uint64_t)L.sub_bytes + (w >> 1);
  int64_t lo = 0, hi = L.ntiles - 1;              // tile of line i: first tile whose inclusive line count exceeds i
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((L.desc[mid] & ((1ull << 62) - 1)) > i) hi = mid; else lo = mid + 1;
  }
  return (uint64_t)lo * JX_TILE + (w >> 1);
}
#endif
struct TokOut {
  LineIndex li;
  uint4* recs = nullptr;       // [n_lines] (n_lines = record SLOTS: lines, or sub-tiles x cap with "no record" fillers)
  uint4* hashes = nullptr;     // [n_lines] names mode: {hq.lo, hq.hi, hs.lo, hs.hi}, 0 = no name
  uint64_t* desc = nullptr;    // per tile: (2 << 62) | inclusive line count
  int64_t ntiles = 0;
  const uint8_t* d_text = nullptr;
  int64_t nbytes = 0;
  uint32_t n_lines = 0;
  // 0 lines 1 comment 2 mapped 3 unmapped 4 hard 5 longname 6 overflow 7 first bad offset+1
  // 8 excluded 9 self/near-diagonal drops
  unsigned long long stats[12] = {0};
};
int jx_run_tokenizer(jx_ctx* ctx, Arena& arena, const char* text, int64_t nbytes, int on_device,
                     const TokMode& mode, TokOut& out);

// raw text lines of the records rec_idx[0..cnt) (device array, file order) -> malloc'ed host buffer
int jx_gather_lines(jx_ctx* ctx, Arena& arena, const uint8_t* d_text, int64_t nbytes, const LineIndex& li,
                    const uint4* recs, const uint32_t* rec_idx, uint32_t cnt, char** out, int64_t* out_len, int rstrip = 0,
                    uint64_t* dev_out = nullptr);   // dev_out: leave the packed lines in ctx->lines_dev instead (no host copy)

// ---- stage cores shared by the C-ABI entry points (jx_filter.cu / jx_scan.cu / jx_liftover.cu / jx_pipeline.cu) ----
// device-side anchors of a scan (arena memory of the caller): members cluster by cluster, block id per member
struct ScanDev {
  int32_t* q = nullptr; int32_t* s = nullptr; float* score = nullptr; uint32_t* block = nullptr;
  uint32_t* count = nullptr;     // device: [0] = anchors
  uint32_t n = 0, nb = 0;        // host copies (nb only when the host half has run)
  bool defer_copy = false;       // in: leave the result copy in flight (`stage`), the caller finishes with scan_finish_host
  bool no_host = false;          // in: the anchors stay on the device only (no result copy; counts are still returned)
  char* stage = nullptr;
};
void scan_finish_host(jx_scan_result* out, const char* stage, uint32_t m);
int jx_scan_survivors(jx_ctx* ctx, Arena& arena, const jx_scan_opts* opts, jx_scan_result* out, ScanDev* dev);   // of the last jx_blastfilter
int jx_liftover_core(jx_ctx* ctx, Arena& arena, const uint4* recs, uint32_t n_slots, const unsigned long long* tstats, int dist,
                     const int32_t* d_aq, const int32_t* d_as, const uint32_t* d_ablock, uint32_t nal,
                     const int32_t* h_aq, const int32_t* h_as, const int32_t* h_ablock, jx_liftover_result* out);

// ---- shared device helpers ---------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t score_ord(uint32_t bits) {   // monotone float32 -> uint32
  return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  for (;;) {
    uint32_t p = __ldcg(parent + x);
    if (p == x) return x;
    uint32_t gp = __ldcg(parent + p);
    if (gp != p) parent[x] = gp;   // path halving: benign race (any ancestor is valid)
    x = p;
  }
}
// lock-free union: hook the larger root under the smaller one, so root == min member index
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { uint32_t t = a; a = b; b = t; }
    if (atomicCAS(parent + b, b, a) == b) return;
  }
}
// append `pred` items of a warp to a global list; returns the slot (valid when pred)
__device__ __forceinline__ uint32_t warp_append(bool pred, uint32_t* counter) {
  unsigned m = __ballot_sync(0xFFFFFFFFu, pred);
  if (!m) return 0;
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  return base + __popc(m & ((1u << lane) - 1));
}
// U predicates per lane, ONE atomic per CTA: every warp of the CTA must call this the same number of
// times (uniform trip counts); `s_cnt` is a shared array of blockDim.x / 32 + 1 words
template <int U>
__device__ __forceinline__ void block_append_multi(const bool (&pred)[U], uint32_t* counter, uint32_t (&slot)[U],
                                                   uint32_t* s_cnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t total = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred[u]);
    slot[u] = total + __popc(m & lt);
    total += __popc(m);
  }
  if (lane == 0) s_cnt[warp] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t sum = 0;
    for (int w = 0; w < nwarp; ++w) { const uint32_t c = s_cnt[w]; s_cnt[w] = sum; sum += c; }
    s_cnt[nwarp] = sum ? atomicAdd(counter, sum) : 0u;
  }
  __syncthreads();
  const uint32_t base = s_cnt[nwarp] + s_cnt[warp];
  __syncthreads();                       // s_cnt is reused by the next call
#pragma unroll
  for (int u = 0; u < U; ++u) slot[u] += base;
}
// the same for U predicates per lane with ONE atomic per warp: the counter is a single address, and
// 20 M records / 32 atomics on it (0.5 ns each, serialised in L2) used to be most of the selection
// kernels' time
template <int U>
__device__ __forceinline__ void warp_append_multi(const bool (&pred)[U], uint32_t* counter, uint32_t (&slot)[U]) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t total = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, pred[u]);
    slot[u] = total + __popc(m & lt);
    total += __popc(m);
  }
  uint32_t base = 0;
  if (total) {
    if (lane == 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) slot[u] += base;
}
#endif

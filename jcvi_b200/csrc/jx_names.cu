// Name handling for tables that come WITHOUT a BED: `jcvi.formats.blast cscore`
// (src/jcvi/formats/blast.py:1098-1189) keys its dictionaries by whatever names the BLAST table
// holds.  jx_distinct_names finds one representative line per distinct (stripped) name on the
// device; the host turns the handful of representatives into a name list, loads it as the name
// table of both sides (jx_bed_load) and the rest of `cscore` is the existing C-score machinery
// (jx_cscore_begin / jx_cscore_finish).  Also here: the host-side BlastLine helpers the Python
// layer uses on the few surviving lines (the same glibc sscanf / sprintf calls cblast.pyx makes).
#include <cstdio>
#include "jx_internal.cuh"

namespace {

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

// two sort keys per line: the hash of the query name and of the subject name (~0 = no name)
__global__ void k_name_keys(const uint4* __restrict__ hashes, uint32_t n, uint64_t* keys, uint32_t* vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 h = hashes[i];
  const uint64_t hq = (uint64_t)h.x | ((uint64_t)h.y << 32), hs = (uint64_t)h.z | ((uint64_t)h.w << 32);
  keys[2 * (size_t)i] = hq ? hq : ~0ull;      vals[2 * (size_t)i] = 2u * i;
  keys[2 * (size_t)i + 1] = hs ? hs : ~0ull;  vals[2 * (size_t)i + 1] = 2u * i + 1u;
}
__global__ void k_name_heads(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* flag) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t k = keys[t];
  flag[t] = (k != ~0ull && (t == 0 || keys[t - 1] != k)) ? 1u : 0u;
}
// the radix sort is stable and vals start ascending: the head of a run is the first occurrence
__global__ void k_take_reps(const uint32_t* __restrict__ vals, const uint32_t* __restrict__ flag,
                            const uint32_t* __restrict__ pos, uint32_t n, uint32_t* line, uint8_t* side) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || !flag[t]) return;
  const uint32_t v = vals[t];
  line[pos[t]] = v >> 1;
  side[pos[t]] = (uint8_t)(v & 1u);
}

}  // namespace

extern "C" {

int jx_distinct_names(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, uint32_t seed,
                      char** lines, int64_t* lines_len, uint8_t** side, int64_t* n_names, int64_t stats[4]) {
  if (!ctx || !lines || !lines_len || !side || !n_names) return JX_EARG;
  *lines = nullptr; *lines_len = 0; *side = nullptr; *n_names = 0;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  int rc;
  TokMode mode;
  mode.strip = strip_names;
  mode.names = true;
  mode.seed = seed;
  TokOut tok;
  if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
  if (stats) {
    stats[0] = (int64_t)(tok.stats[0] - tok.stats[1]);   // lines that are not comments
    stats[1] = (int64_t)tok.stats[2];                    // lines with two names
    stats[2] = (int64_t)tok.stats[3];                    // lines with fewer than two fields
    stats[3] = (int64_t)tok.stats[1];                    // '#' lines
  }
  const uint32_t n = tok.n_lines;
  if (n >= (1u << 31)) { ctx->err = "more than 2^31 lines"; return JX_EUNSUPPORTED; }
  if (!n) { *lines = (char*)calloc(1, 1); *side = (uint8_t*)calloc(1, 1); return JX_OK; }
  const uint32_t m = 2u * n;
  JX_ALLOC(keys, uint64_t, m, false);
  JX_ALLOC(vals, uint32_t, m, false);
  JX_LAUNCH(k_name_keys, nblk(n), 256, 0, tok.hashes, n, keys, vals);
  uint64_t* ks = keys; uint32_t* vs = vals;
  if ((rc = jx_sort_pairs(ctx, arena, ks, vs, m, 0, 64))) return rc;
  JX_ALLOC(flag, uint32_t, m, false);
  JX_ALLOC(pos, uint32_t, m, false);
  JX_LAUNCH(k_name_heads, nblk(m), 256, 0, ks, m, flag);
  if ((rc = jx_exclusive_sum(ctx, arena, flag, pos, m))) return rc;
  uint32_t h[2];
  JX_CK(cudaMemcpyAsync(&h[0], pos + m - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&h[1], flag + m - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const uint32_t g = h[0] + h[1];
  *n_names = g;
  *side = (uint8_t*)malloc((size_t)g + 1);
  if (!g) { *lines = (char*)calloc(1, 1); return JX_OK; }
  JX_ALLOC(rep_line, uint32_t, g, false);
  JX_ALLOC(rep_side, uint8_t, g, false);
  JX_LAUNCH(k_take_reps, nblk(m), 256, 0, vs, flag, pos, m, rep_line, rep_side);
  JX_CK(cudaMemcpyAsync(*side, rep_side, g, cudaMemcpyDeviceToHost, ctx->stream));
  rc = jx_gather_lines(ctx, arena, tok.d_text, tok.nbytes, tok.desc, tok.ntiles, tok.recs, rep_line, g, lines, lines_len);
  JX_CK(cudaStreamSynchronize(ctx->stream));
  return rc;
}

// per-name best scores of the open jx_cscore_begin session, as float32 values
int jx_best_table_copy(jx_ctx* ctx, float* out, uint32_t n) {
  if (!ctx || !out) return JX_EARG;
  if (!ctx->sess_best || n > ctx->n_name_ids) { ctx->err = "jx_best_table_copy without jx_cscore_begin"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  JX_CK(cudaMemcpyAsync(out, ctx->sess_best, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  return JX_OK;
}

// tokenizer counters of the records retained by jx_cscore_begin: 0 lines (non-'#'), 1 mapped, 2 unmapped
int jx_session_counts(jx_ctx* ctx, int64_t out[3]) {
  if (!ctx || !out) return JX_EARG;
  if (!ctx->cache.valid) { ctx->err = "no open session"; return JX_EARG; }
  out[0] = (int64_t)(ctx->cache.stats[0] - ctx->cache.stats[1]);
  out[1] = (int64_t)ctx->cache.stats[2];
  out[2] = (int64_t)ctx->cache.stats[3];
  return JX_OK;
}

// BlastLine(line) and str(BlastLine) exactly as cblast.pyx does them (format strings cblast.pyx:15,17):
// glibc sscanf into char[128] names / float / int / double, start<=stop swap with orientation, sprintf.
// Host only; for the few lines that survive a filter.  Returns the number of converted fields.
int jx_host_blastline(const char* line, char* query128, char* subject128, float* pctid, float* score, char* str_out,
                      int str_cap) {
  char q[128] = {0}, s[128] = {0};
  float pct = 0.f, sc = 0.f;
  int hitlen = 0, nmismatch = 0, ngaps = 0, qstart = 0, qstop = 0, sstart = 0, sstop = 0;
  double evalue = 0.0;
  const int nconv = sscanf(line, "%127s\t%127s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f", q, s, &pct, &hitlen, &nmismatch,
                           &ngaps, &qstart, &qstop, &sstart, &sstop, &evalue, &sc);
  char orientation = '+';
  if (qstart > qstop) { int t = qstart; qstart = qstop; qstop = t; orientation = '-'; }
  if (sstart > sstop) { int t = sstart; sstart = sstop; sstop = t; orientation = '-'; }
  if (query128) memcpy(query128, q, 128);
  if (subject128) memcpy(subject128, s, 128);
  if (pctid) *pctid = pct;
  if (score) *score = sc;
  (void)orientation;   // __str__ builds a swapped-back args list but prints the object's own fields (cblast.pyx:144-156)
  if (str_out && str_cap > 0)
    snprintf(str_out, (size_t)str_cap, "%s\t%s\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.3g", q, s, (double)pct, hitlen,
             nmismatch, ngaps, qstart, qstop, sstart, sstop, evalue, (double)sc);
  return nconv;
}

}  // extern "C"

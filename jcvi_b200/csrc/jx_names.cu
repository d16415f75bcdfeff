// Name handling for tables that come WITHOUT a BED: `jcvi.formats.blast cscore`
// (src/jcvi/formats/blast.py:1098-1189) keys its dictionaries by whatever names the BLAST table
// holds.  jx_distinct_names finds one representative line per distinct (stripped) name on the
// device; the host turns the handful of representatives into a name list, loads it as the name
// table of both sides (jx_bed_load) and the rest of `cscore` is the existing C-score machinery
// (jx_cscore_begin / jx_cscore_finish).  Also here: the host-side BlastLine helpers the Python
// layer uses on the few surviving lines (the same glibc sscanf / sprintf calls cblast.pyx makes).
#include <algorithm>
#include <cstdio>
#include <map>
#include <string>
#include <vector>
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

// formats.blast filter: the tokenizer's filter-mode records are {keep, rstripped length, 64-bit row offset}
__global__ void k_kept_len(const uint4* __restrict__ recs, uint32_t n, uint32_t* len) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint4 r = recs[i]; len[i] = r.x == 1u ? r.y + 1u : 0u; }      // + '\n'
}
// one warp per row
__global__ void k_copy_kept(const uint8_t* __restrict__ text, const uint4* __restrict__ recs, const uint64_t* __restrict__ pos,
                            uint32_t n, uint8_t* out) {
  const uint32_t w = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= n) return;
  const uint4 r = recs[w];
  if (r.x != 1u) return;
  const uint8_t* src = text + (((uint64_t)r.w << 32) | r.z);
  uint8_t* dst = out + pos[w];
  const uint32_t L = r.y;
  for (uint32_t i = lane; i < L; i += 32) dst[i] = src[i];
  if (lane == 0) dst[L] = '\n';
}

}  // namespace

extern "C" {

// `jcvi.formats.blast filter` (src/jcvi/formats/blast.py:620-709): the rows whose BlastLine passes
//   remove = score < o.score or pctid < o.pctid or hitlen < o.hitlen or evalue > o.evalue or noids
//   if inverse: remove = not remove;  remove = remove or (noself and query == subject)
// in file order, each as row.rstrip() + "\n".  use_ids: "c.query in ids and c.subject in ids" against the
// names loaded with jx_bed_load (the ids file as a name list on both sides, raw names).
int jx_blast_filter(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, double score, double pctid,
                    int64_t hitlen, double evalue, int noself, int inverse, int use_ids, char** out, int64_t* out_len,
                    int64_t stats[4]) {
  if (!ctx || !out || !out_len) return JX_EARG;
  *out = nullptr; *out_len = 0;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  int rc;
  TokMode mode;
  mode.strip = 0;
  mode.filter = true;
  mode.f_score = score; mode.f_pctid = pctid; mode.f_hitlen = (long long)hitlen; mode.f_evalue = evalue;
  mode.f_noself = noself != 0; mode.f_inverse = inverse != 0; mode.f_ids = use_ids != 0;
  TokOut tok;
  if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
  if (stats) {
    stats[0] = (int64_t)(tok.stats[0] - tok.stats[1]);   // rows (non-'#'; includes CR LF gaps)
    stats[1] = (int64_t)tok.stats[2];                    // kept
    stats[2] = (int64_t)tok.stats[3];                    // removed
    stats[3] = (int64_t)tok.stats[1];                    // '#' lines
  }
  const uint32_t n = tok.n_lines;
  if (!n) { *out = (char*)calloc(1, 1); return JX_OK; }
  JX_ALLOC(len, uint32_t, n, false);
  JX_ALLOC(pos, uint64_t, n, false);
  JX_LAUNCH(k_kept_len, nblk(n), 256, 0, tok.recs, n, len);
  if ((rc = jx_exclusive_sum64(ctx, arena, len, pos, n))) return rc;
  uint64_t hp = 0; uint32_t hl = 0;
  JX_CK(cudaMemcpyAsync(&hp, pos + n - 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&hl, len + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const uint64_t total = hp + hl;
  if (!total) { *out = (char*)calloc(1, 1); return JX_OK; }
  JX_ALLOC(packed, uint8_t, (size_t)total, false);
  JX_LAUNCH(k_copy_kept, nblk((size_t)n * 32), 256, 0, tok.d_text, tok.recs, pos, n, packed);
  char* host = (char*)malloc((size_t)total + 1);
  if (!host) { ctx->err = "out of host memory"; return JX_EINTERNAL; }
  if ((rc = jx_d2h_big(ctx, host, packed, (size_t)total))) { free(host); return rc; }
  host[total] = 0;
  *out = host; *out_len = (int64_t)total;
  return JX_OK;
}

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
  rc = jx_gather_lines(ctx, arena, tok.d_text, tok.nbytes, tok.li, tok.recs, rep_line, g, lines, lines_len);
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

// ---- the whole of `jcvi.formats.blast cscore` (src/jcvi/formats/blast.py:1098-1189) in one call -------
// GPU: both passes over the table (distinct names; per-name best score + float64 ratio predicate).
// Host, O(distinct names) + O(surviving lines): the name list, the reference's `pairs` dictionary
// (largest C-score per (query, subject), first line on ties), sorted(), the "{:.2f}" / "{:.1f}" writer
// and str(BlastLine) -- with the same glibc calls cblast makes.
namespace {
struct PtrBytes { const unsigned char* p; inline uint32_t operator()(int i) const { return p[i]; } };
inline bool c_isspace(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
// gene_name (cbook.py:308-329) through the tokenizer's own strip_isoform
inline std::string host_gene_name(const char* s, size_t n) {
  PtrBytes rd{(const unsigned char*)s};
  int b = (int)n;
  jx::strip_isoform(rd, 0, b);
  return std::string(s, (size_t)b);
}
}  // namespace

int jx_blast_cscore(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, double cutoff,
                    int want_pct, int want_blast, char** out, int64_t* out_len, char** blast_out, int64_t* blast_len,
                    int64_t stats[4]) {
  if (!ctx || !out || !out_len) return JX_EARG;
  *out = nullptr; *out_len = 0;
  if (blast_out) *blast_out = nullptr;
  if (blast_len) *blast_len = 0;
  cudaSetDevice(ctx->device);
  int rc;
  uint64_t dptr = 0;
  if (!text_on_device) {
    if ((rc = jx_text_upload(ctx, text, nbytes, &dptr))) return rc;
  } else dptr = (uint64_t)(uintptr_t)text;
  const char* dtext = (const char*)(uintptr_t)dptr;

  std::vector<std::string> names;
  bool separated = false;
  for (uint32_t seed = 0; seed < 4 && !separated; ++seed) {
    char* lines = nullptr; int64_t lines_len = 0; uint8_t* side = nullptr; int64_t n = 0; int64_t st[4] = {0, 0, 0, 0};
    if ((rc = jx_distinct_names(ctx, dtext, nbytes, 1, strip_names, seed, &lines, &lines_len, &side, &n, st))) return rc;
    if (stats) memcpy(stats, st, sizeof st);
    if (st[2]) {
      free(lines); free(side);
      ctx->err = std::to_string(st[2]) + " line(s) with fewer than two fields (a BlastLine with an empty name: the "
                 "reference divides by best_score[\"\"] == 0.0)";
      return JX_EUNSUPPORTED;
    }
    names.clear();
    names.reserve((size_t)n);
    const char* p = lines; const char* end = lines + lines_len;
    for (int64_t i = 0; i < n && p < end; ++i) {
      const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
      if (!eol) eol = end;
      const char* q = p;
      for (int col = 0;; ++col) {                       // sscanf("%s %s"): whitespace-delimited tokens
        while (q < eol && c_isspace((unsigned char)*q)) ++q;
        const char* t0 = q;
        while (q < eol && !c_isspace((unsigned char)*q)) ++q;
        if (col == (int)side[i]) { names.push_back(strip_names ? host_gene_name(t0, (size_t)(q - t0)) : std::string(t0, q)); break; }
      }
      p = eol + 1;
    }
    free(lines); free(side);
    std::sort(names.begin(), names.end());
    names.erase(std::unique(names.begin(), names.end()), names.end());
    // the names become the name table of both sides: rank == name id == position in sorted order
    const size_t G = names.size();
    std::string blob; std::vector<uint32_t> off(G + 1, 0);
    for (size_t i = 0; i < G; ++i) { blob += names[i]; off[i + 1] = (uint32_t)blob.size(); }
    std::vector<uint8_t> ones(G + 1, 1);
    std::vector<int32_t> zeros(G + 1, 0), ends(G + 1, (int32_t)G), same(G + 1, -1);
    std::vector<uint32_t> ids(G + 1);
    for (size_t i = 0; i <= G; ++i) ids[i] = (uint32_t)i;
    jx_bed b;
    b.n_genes = (int64_t)G; b.names = blob.data(); b.name_off = off.data(); b.indexable = ones.data();
    b.chr_lex = zeros.data(); b.chr_begin = zeros.data(); b.chr_end = ends.data(); b.length = zeros.data();
    b.accn_lexrank = ids.data(); b.name_id = ids.data(); b.n_chr = 1;
    if ((rc = jx_bed_load(ctx, 0, &b)) || (rc = jx_bed_load(ctx, 1, &b))) return rc;
    if ((rc = jx_pair_setup(ctx, 0, same.data(), (uint32_t)G))) return rc;
    uint64_t best_ptr = 0; uint32_t n_ids = 0;
    if ((rc = jx_cscore_begin(ctx, dtext, nbytes, 1, strip_names, nullptr, 0, &best_ptr, &n_ids))) return rc;
    int64_t cnt[3];
    if ((rc = jx_session_counts(ctx, cnt))) return rc;
    separated = cnt[2] == 0;        // an unmapped line = a name that lost its table entry to a 64-bit hash collision
  }
  if (!separated) { ctx->err = "could not separate the names of the table"; return JX_EINTERNAL; }
  const size_t G = names.size();
  std::vector<float> best(G + 1, 0.f);
  if (G && (rc = jx_best_table_copy(ctx, best.data(), (uint32_t)G))) return rc;
  char* kept = nullptr; int64_t kept_len = 0, n_kept = 0;
  if ((rc = jx_cscore_finish(ctx, cutoff, &kept, &kept_len, &n_kept))) return rc;

  struct Pick { double s; float pctid; std::string str; };
  std::map<std::pair<std::string, std::string>, Pick> pairs;      // iterates like sorted(pairs.items())
  {
    const char* p = kept; const char* end = kept + kept_len;
    std::string row;
    char q128[128], s128[128], sbuf[1024];
    for (int64_t i = 0; i < n_kept && p < end; ++i) {
      const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
      if (!eol) eol = end;
      row.assign(p, eol);
      float pct = 0.f, sc = 0.f;
      jx_host_blastline(row.c_str(), q128, s128, &pct, &sc, want_blast ? sbuf : nullptr, (int)sizeof sbuf);
      std::string qn = strip_names ? host_gene_name(q128, strlen(q128)) : std::string(q128);
      std::string sn = strip_names ? host_gene_name(s128, strlen(s128)) : std::string(s128);
      const auto iq = std::lower_bound(names.begin(), names.end(), qn), is = std::lower_bound(names.begin(), names.end(), sn);
      if (iq == names.end() || *iq != qn || is == names.end() || *is != sn) { free(kept); ctx->err = "name table lost a name"; return JX_EINTERNAL; }
      const double den = std::max((double)best[(size_t)(iq - names.begin())], (double)best[(size_t)(is - names.begin())]);
      const double c = (double)sc / den;                 // den > 0: the device predicate already divided by it
      if (c > cutoff) {
        auto key = std::make_pair(std::move(qn), std::move(sn));
        auto it = pairs.find(key);
        if (it == pairs.end()) pairs.emplace(std::move(key), Pick{c, pct, want_blast ? std::string(sbuf) : std::string()});
        else if (c > it->second.s) it->second = Pick{c, pct, want_blast ? std::string(sbuf) : std::string()};
      }
      p = eol + 1;
    }
  }
  free(kept);
  std::string o, ob;
  char num[64];
  for (const auto& kv : pairs) {
    o += kv.first.first; o += '\t'; o += kv.first.second;
    snprintf(num, sizeof num, "\t%.2f", kv.second.s); o += num;
    if (want_pct) { snprintf(num, sizeof num, "\t%.1f", (double)kv.second.pctid); o += num; }
    o += '\n';
    if (want_blast) { ob += kv.second.str; ob += '\n'; }
  }
  *out = (char*)malloc(o.size() + 1); memcpy(*out, o.data(), o.size()); (*out)[o.size()] = 0; *out_len = (int64_t)o.size();
  if (want_blast && blast_out && blast_len) {
    *blast_out = (char*)malloc(ob.size() + 1); memcpy(*blast_out, ob.data(), ob.size()); (*blast_out)[ob.size()] = 0;
    *blast_len = (int64_t)ob.size();
  }
  if (!text_on_device) jx_text_release(ctx);
  return JX_OK;
}

}  // extern "C"

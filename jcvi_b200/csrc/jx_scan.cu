// Stage 2: synteny scan on the GPU (K9/K10 of SURVEY.md section 2).
//
// Replaces src/jcvi/compara/synteny.py:
//   255-297  read_blast   file-order first-seen de-duplication        -> sort by (qi,si), min line per run
//   239-252  group_hits / 437-452 batch_scan  per chromosome pair      -> stable sort by chromosome-pair id
//   402-434  synteny_scan single linkage (Grouper.join)                -> window neighbour search + lock-free union-find
//   214-219  _score        min(#distinct qi, #distinct si) >= N        -> per-component counters
//   utils/grouper.py:73-81 iteration order (dict insertion)            -> "birth" = smallest sorted index that has an
//                                                                        earlier neighbour; components ordered by birth
#include <algorithm>
#include "jx_internal.cuh"

// host half of a scan result: the pinned block {q, s, score, index, head flags} x m -> result arrays + block starts
void scan_finish_host(jx_scan_result* out, const char* stage, uint32_t m) {
  memcpy(out->q_rank, stage, (size_t)m * 4);
  memcpy(out->s_rank, stage + (size_t)m * 4, (size_t)m * 4);
  memcpy(out->score, stage + (size_t)m * 8, (size_t)m * 4);
  const uint32_t* hidx = (const uint32_t*)(stage + (size_t)m * 12);
  const uint32_t* hhead = (const uint32_t*)(stage + (size_t)m * 16);
  int64_t nb = 0;
  for (uint32_t t = 0; t < m; ++t) nb += hhead[t];
  free(out->block_start);
  out->block_start = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb + 1));
  int64_t b = 0;
  for (uint32_t t = 0; t < m; ++t) {
    if (hhead[t]) out->block_start[b++] = t;
    out->index[t] = hidx[t];
  }
  out->block_start[nb] = m;
  out->n_blocks = nb;
}

namespace {

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

__global__ void k_select_valid(const uint4* __restrict__ recs, uint32_t n, uint4* out, uint32_t* out_idx, uint32_t* counter) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t nround = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    uint4 r = make_uint4(JX_NOREC, 0, 0, 0);
    if (i < n) r = __ldcs(recs + i);
    bool keep = r.x != JX_NOREC;
    uint32_t slot = warp_append(keep, counter);
    if (keep) { out[slot] = r; out_idx[slot] = i; }
  }
}

__global__ void k_make_keys(const uint4* __restrict__ C, uint32_t n, int sb, uint64_t* keys, uint32_t* vals) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint4 r = C[j];
  keys[j] = ((uint64_t)r.x << sb) | r.y;
  vals[j] = j;
}

// first occurrence in file order within each run of equal (qi,si)
__global__ void k_run_first(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                            const uint32_t* __restrict__ Cidx, uint32_t n, uint32_t* flag, uint32_t* pick) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t k = keys[t];
  const bool head = (t == 0) || keys[t - 1] != k;
  flag[t] = head;
  if (!head) return;
  uint32_t bj = vals[t], bi = Cidx[bj];
  for (uint32_t u = t + 1; u < n && keys[u] == k; ++u) {
    uint32_t j = vals[u], i = Cidx[j];
    if (i < bi) { bj = j; bi = i; }
  }
  pick[t] = bj;
}

__global__ void k_scatter_pts(const uint64_t* keys, const uint32_t* pick, const uint32_t* flag, const uint32_t* pos,
                              uint32_t n, int sb, const uint4* C, const int32_t* qchr, const int32_t* schr, int64_t nschr,
                              int32_t* px, int32_t* py, float* psc, uint64_t* cpkey, uint32_t* ident) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || !flag[t]) return;
  uint32_t o = pos[t];
  uint64_t k = keys[t];
  int32_t x = (int32_t)(k >> sb), y = (int32_t)(k & ((1ull << sb) - 1));
  px[o] = x; py[o] = y;
  psc[o] = __uint_as_float(C[pick[t]].z);
  cpkey[o] = (uint64_t)qchr[x] * (uint64_t)nschr + (uint64_t)schr[y];
  ident[o] = o;
}

template <class T>
__global__ void k_permute(const T* in, const uint32_t* perm, uint32_t n, T* out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[perm[t]];
}

__global__ void k_iota(uint32_t* p, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// records known to be valid and pair-unique (blastfilter's own survivors): one sort on
// (chromosome pair, qi, si) replaces select + pair de-duplication + the stable pair sort
__global__ void k_point_keys(const uint4* __restrict__ recs, uint32_t n, int sb, int qb, const int32_t* __restrict__ qchr,
                             const int32_t* __restrict__ schr, int64_t nschr, uint64_t* keys, uint32_t* vals) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint4 r = recs[t];
  const uint64_t cp = (uint64_t)qchr[r.x] * (uint64_t)nschr + (uint64_t)schr[r.y];
  keys[t] = (cp << (sb + qb)) | ((uint64_t)r.x << sb) | (uint64_t)r.y;
  vals[t] = t;
}
__global__ void k_points_from_keys(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                   const uint4* __restrict__ recs, uint32_t n, int sb, int qb, int32_t* px, int32_t* py,
                                   float* psc, uint64_t* cpkey) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t k = keys[t];
  px[t] = (int32_t)((k >> sb) & ((1ull << qb) - 1));
  py[t] = (int32_t)(k & ((1ull << sb) - 1));
  cpkey[t] = k >> (sb + qb);
  psc[t] = __uint_as_float(recs[vals[t]].z);
}

// synteny_scan inner loops (synteny.py:402-434): for point i walk back while x_i - x_j <= xdist (same chromosome pair).
// Measured on a dense set (9.2 M points, dist 40, ~800 earlier points per window, tests/perf/scan_dense.py): a variant
// that seeks past the parts of a row outside [y_i - ydist, y_i + ydist] (galloping + binary search) took 56 ms against
// 61 ms for this walk and was 1.7x slower on the usual sparse set -- the dense case is bound by the ~20 unions per point
// into a few giant components, not by the walk -- so the plain walk stays.
__global__ void k_link(const int32_t* __restrict__ px, const int32_t* __restrict__ py, const uint64_t* __restrict__ cp,
                       uint32_t n, int xdist, int ydist, int is_self, int intrabound, uint32_t* parent, uint8_t* hasprev) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t xi = px[i], yi = py[i];
  const uint64_t ci = cp[i];
  const int di = abs(xi - yi);
  bool any = false;
  for (int64_t j = (int64_t)i - 1; j >= 0; --j) {
    if (cp[j] != ci) break;
    const int32_t xj = px[j];
    if (xi - xj > xdist) break;
    const int32_t yj = py[j];
    if (abs(yi - yj) > ydist) continue;
    if (is_self) {
      int dj = abs(xj - yj);
      if ((di < dj ? di : dj) < intrabound) continue;
    }
    uf_union(parent, i, (uint32_t)j);
    any = true;
  }
  hasprev[i] = any;
}


__global__ void k_label(uint32_t* parent, const uint8_t* hasprev, uint32_t n, uint32_t* label, uint32_t* birth) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r = uf_find(parent, i);
  label[i] = r;
  if (hasprev[i]) atomicMin(birth + r, i);
}

// #distinct qi and #distinct si per component: (label, x) and (label, y) pairs through ONE hash set (x entries carry
// bit 31); the thread that inserts a new pair counts it
__device__ __forceinline__ void count_distinct(unsigned long long* K, uint32_t mask, uint32_t l, uint32_t v, uint32_t* cnt) {
  const unsigned long long key = ((unsigned long long)l << 32) | v;
  uint32_t h = l * 0x9E3779B1u ^ v * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  uint32_t slot = h & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(K + slot, 0xFFFFFFFFFFFFFFFFull, key);
    if (prev == 0xFFFFFFFFFFFFFFFFull) { atomicAdd(cnt + l, 1u); return; }
    if (prev == key) return;
    slot = (slot + 1u) & mask;
  }
}
__global__ void k_count_xy(const uint32_t* __restrict__ label, const int32_t* __restrict__ px, const int32_t* __restrict__ py,
                           uint32_t n, unsigned long long* K, uint32_t mask, uint32_t* nx, uint32_t* ny) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t l = label[i];
  count_distinct(K, mask, l, (uint32_t)px[i] | 0x80000000u, nx);
  count_distinct(K, mask, l, (uint32_t)py[i], ny);
}

// flag = the point's cluster passes min(#distinct qi, #distinct si) >= N; output key = (birth of the cluster, index):
// clusters in Grouper's iteration order, members ascending.  Points that do not pass get a key above all others.
__global__ void k_pass_keys(const uint32_t* label, const uint32_t* birth, const uint32_t* nx, const uint32_t* ny,
                            uint32_t n, int N, int ib, uint64_t* keys, uint32_t* vals, uint32_t* counts /* 0 passing, 1 clustered */) {
  const uint32_t nround = (n + 31u) & ~31u;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nround) return;
  bool pass = false, clustered = false;
  if (i < n) {
    const uint32_t l = label[i];
    clustered = birth[l] != 0xFFFFFFFFu;            // isolated points never enter the Grouper
    const uint32_t s = nx[l] < ny[l] ? nx[l] : ny[l];
    pass = clustered && (int)s >= N;
    keys[i] = pass ? (((uint64_t)birth[l] << ib) | i) : (1ull << (2 * ib));
    vals[i] = i;
  }
  const unsigned mp = __ballot_sync(0xFFFFFFFFu, pass), mc = __ballot_sync(0xFFFFFFFFu, clustered);
  if ((threadIdx.x & 31) == 0) {
    if (mp) atomicAdd(counts, (uint32_t)__popc(mp));
    if (mc) atomicAdd(counts + 1, (uint32_t)__popc(mc));
  }
}
// the first *m sorted entries are the anchors
__global__ void k_emit(const uint64_t* keys, const uint32_t* vals, const uint32_t* m, int ib, const int32_t* px, const int32_t* py,
                       const float* psc, const uint32_t* pidx, int32_t* oq, int32_t* os, float* osc, uint32_t* oidx,
                       uint32_t* head) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *m) return;
  uint32_t i = vals[t];
  oq[t] = px[i]; os[t] = py[i]; osc[t] = psc[i];
  if (oidx) oidx[t] = pidx ? pidx[i] : i;
  head[t] = (t == 0) || (keys[t] >> ib) != (keys[t - 1] >> ib);
}
__global__ void k_block_ids(const uint32_t* head, const uint32_t* excl, const uint32_t* m, uint32_t* block) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < *m) block[t] = excl[t] + head[t] - 1u;
}

// Core: points already sorted by (chromosome pair, qi, si) on the device.  One host round trip (the result copy).
// `dev` (optional) receives the anchors as device arrays, valid as long as the caller's Arena.
int scan_sorted_points(jx_ctx* ctx, Arena& arena, const int32_t* px, const int32_t* py, const float* psc,
                       const uint64_t* cp, const uint32_t* pidx, uint32_t n, int is_self, const jx_scan_opts* o,
                       jx_scan_result* out, ScanDev* dev, const uint32_t* extra_flag) {
  out->n_anchors = 0; out->n_blocks = 0;
  out->block_start = (int64_t*)calloc(1, sizeof(int64_t));
  if (dev) { const bool defer = dev->defer_copy, nh = dev->no_host; *dev = ScanDev(); dev->defer_copy = defer; dev->no_host = nh; }
  if (!n) {
    out->q_rank = (int32_t*)malloc(4); out->s_rank = (int32_t*)malloc(4); out->score = (float*)malloc(4);
    out->index = (int64_t*)malloc(8);
    if (extra_flag) {
      uint32_t hb = 0;
      JX_CK(cudaMemcpyAsync(&hb, extra_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaStreamSynchronize(ctx->stream));
      if (hb) { ctx->err = "score outside the exactly handled %.3g range"; return JX_EUNSUPPORTED; }
    }
    return JX_OK;
  }
  JX_ALLOC(parent, uint32_t, n, false);
  JX_ALLOC(hasprev, uint8_t, n, false);
  JX_ALLOC(label, uint32_t, n, false);
  JX_ALLOC(birth, uint32_t, n, false);
  JX_ALLOC(nx, uint32_t, n, true);
  JX_ALLOC(ny, uint32_t, n, true);
  JX_ALLOC(d_cnt, uint32_t, 4, true);
  JX_CK(cudaMemsetAsync(birth, 0xFF, (size_t)n * 4, ctx->stream));
  JX_TRACE("scan: points ready");
  JX_LAUNCH(k_iota, nblk(n), 256, 0, parent, n);
  JX_LAUNCH(k_link, nblk(n, 128), 128, 0, px, py, cp, n, o->xdist, o->ydist, is_self, o->intrabound, parent, hasprev);
  JX_LAUNCH(k_label, nblk(n), 256, 0, parent, hasprev, n, label, birth);
  JX_TRACE("scan: link+label");
  {
    uint32_t hcap = 1024;
    while ((uint64_t)hcap < 4ull * n + 16u && hcap < (1u << 31)) hcap <<= 1;
    JX_ALLOC(HK, unsigned long long, hcap, false);
    JX_CK(cudaMemsetAsync(HK, 0xFF, (size_t)hcap * 8, ctx->stream));
    JX_LAUNCH(k_count_xy, nblk(n), 256, 0, label, px, py, n, HK, hcap - 1u, nx, ny);
  }
  JX_TRACE("scan: counts");
  const int ib = jx_bits((uint64_t)n);
  JX_ALLOC(okeys, uint64_t, n, false);
  JX_ALLOC(ovals, uint32_t, n, false);
  JX_LAUNCH(k_pass_keys, nblk(n), 256, 0, label, birth, nx, ny, n, o->min_size, ib, okeys, ovals, d_cnt);
  uint64_t* ok = okeys; uint32_t* ov = ovals;
  int rc;
  if ((rc = jx_sort_pairs(ctx, arena, ok, ov, n, 0, 2 * ib + 1))) return rc;
  JX_TRACE("scan: out sort");
  // the five result arrays are one device block: ONE copy into pinned staging
  JX_ALLOC(oblk, uint32_t, (size_t)n * 5 + 4, false);
  int32_t* oq = (int32_t*)oblk; int32_t* os = (int32_t*)(oblk + n); float* osc = (float*)(oblk + 2 * (size_t)n);
  uint32_t* oidx = oblk + 3 * (size_t)n; uint32_t* head = oblk + 4 * (size_t)n;
  JX_LAUNCH(k_emit, nblk(n), 256, 0, ok, ov, d_cnt, ib, px, py, psc, pidx, oq, os, osc, oidx, head);
  if (dev) {                                      // block id of every anchor, for liftover on the device
    JX_ALLOC(excl, uint32_t, n, false);
    JX_ALLOC(blk, uint32_t, n, false);
    if ((rc = jx_exclusive_sum(ctx, arena, head, excl, n))) return rc;     // (entries beyond *m are never read)
    JX_LAUNCH(k_block_ids, nblk(n), 256, 0, head, excl, d_cnt, blk);
    dev->q = oq; dev->s = os; dev->score = osc; dev->block = blk; dev->count = d_cnt;
  }
  // counts first (8 bytes), then exactly the anchors
  uint32_t hc[2] = {0, 0}, hbad = 0;
  JX_CK(cudaMemcpyAsync(hc, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (extra_flag) JX_CK(cudaMemcpyAsync(&hbad, extra_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (hbad) { ctx->err = "score outside the exactly handled %.3g range"; return JX_EUNSUPPORTED; }
  const uint32_t m = hc[0];
  if (dev) dev->n = m;
  out->stats[2] = hc[1];
  out->q_rank = (int32_t*)malloc(4 * ((size_t)m + 1));
  out->s_rank = (int32_t*)malloc(4 * ((size_t)m + 1));
  out->score = (float*)malloc(4 * ((size_t)m + 1));
  out->index = (int64_t*)malloc(8 * ((size_t)m + 1));
  out->n_anchors = m;
  if (dev && dev->no_host) return JX_OK;
  if (m) {
    char* stage = nullptr;
    if ((rc = jx_pin_scan(ctx, (size_t)m * 20, &stage))) return rc;
    if (dev && dev->defer_copy) {
      // the caller overlaps this copy with its next stage (copy stream; everything above has completed: the
      // count read-back synchronised) and finishes the result itself
      for (int k = 0; k < 5; ++k)
        JX_CK(cudaMemcpyAsync(stage + (size_t)m * 4 * k, oblk + (size_t)n * k, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
      dev->stage = stage;
      return JX_OK;
    }
    for (int k = 0; k < 5; ++k)
      JX_CK(cudaMemcpyAsync(stage + (size_t)m * 4 * k, oblk + (size_t)n * k, (size_t)m * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    scan_finish_host(out, stage, m);
    JX_TRACE("scan: emit + D2H");
  }
  if (dev) dev->nb = (uint32_t)out->n_blocks;
  JX_CK(cudaGetLastError());
  return JX_OK;
}

int compact_positions(jx_ctx* ctx, Arena& arena, const uint32_t* flag, size_t n, uint32_t*& pos, uint32_t& count) {
  count = 0; pos = nullptr;
  if (!n) return JX_OK;
  JX_ALLOC(p, uint32_t, n, false);
  int rc = jx_exclusive_sum(ctx, arena, flag, p, n);
  if (rc) return rc;
  uint32_t last[2];
  JX_CK(cudaMemcpyAsync(&last[0], p + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&last[1], flag + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  count = last[0] + last[1];
  pos = p;
  return JX_OK;
}

// records (file order, holes marked JX_NOREC) -> scan
int scan_records(jx_ctx* ctx, Arena& arena, const uint4* recs, uint32_t nrec, size_t nvalid_hint, const jx_scan_opts* o,
                 jx_scan_result* out, bool valid_unique = false, ScanDev* dev = nullptr, const uint32_t* extra_flag = nullptr) {
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  int rc;
  {
    const int sb0 = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)), qb0 = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
    const uint64_t maxcp0 = (uint64_t)std::max<int64_t>(Q.n_chr, 1) * (uint64_t)std::max<int64_t>(S.n_chr, 1);
    if (valid_unique && nrec && jx_bits(maxcp0) + sb0 + qb0 <= 64) {
      const uint32_t n = nrec;
      JX_ALLOC(keys, uint64_t, n, false);
      JX_ALLOC(vals, uint32_t, n, false);
      JX_LAUNCH(k_point_keys, nblk(n), 256, 0, recs, n, sb0, qb0, Q.chr_lex, S.chr_lex, (int64_t)S.n_chr, keys, vals);
      uint64_t* ks = keys; uint32_t* vs = vals;
      if ((rc = jx_sort_pairs(ctx, arena, ks, vs, n, 0, jx_bits(maxcp0) + sb0 + qb0))) return rc;
      JX_ALLOC(x1, int32_t, n, false);
      JX_ALLOC(y1, int32_t, n, false);
      JX_ALLOC(s1, float, n, false);
      JX_ALLOC(c1, uint64_t, n, false);
      JX_LAUNCH(k_points_from_keys, nblk(n), 256, 0, ks, vs, recs, n, sb0, qb0, x1, y1, s1, c1);
      JX_TRACE("scan: single sort (valid, unique input)");
      out->stats[1] = n;
      return scan_sorted_points(ctx, arena, x1, y1, s1, c1, nullptr, n, ctx->is_self ? 1 : 0, o, out, dev, extra_flag);
    }
  }
  JX_ALLOC(C, uint4, nvalid_hint + 1, false);
  JX_ALLOC(Cidx, uint32_t, nvalid_hint + 1, false);
  JX_ALLOC(d_cnt, uint32_t, 1, true);
  if (nrec) {
    unsigned grid = std::min<unsigned>(nblk(nrec), 148u * 16u);
    JX_LAUNCH(k_select_valid, grid, 256, 0, recs, nrec, C, Cidx, d_cnt);
  }
  uint32_t nC = 0;
  JX_CK(cudaMemcpyAsync(&nC, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const int sb = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)), qb = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
  uint32_t n = 0;
  int32_t *px = nullptr, *py = nullptr; float* psc = nullptr; uint64_t* cp = nullptr;
  if (nC) {
    JX_ALLOC(keys, uint64_t, nC, false);
    JX_ALLOC(vals, uint32_t, nC, false);
  JX_TRACE("scan: select_valid");
    JX_LAUNCH(k_make_keys, nblk(nC), 256, 0, C, nC, sb, keys, vals);
    uint64_t* ks = keys; uint32_t* vs = vals;
    if ((rc = jx_sort_pairs(ctx, arena, ks, vs, nC, 0, sb + qb))) return rc;
    JX_ALLOC(flag, uint32_t, nC, false);
    JX_ALLOC(pick, uint32_t, nC, false);
    JX_LAUNCH(k_run_first, nblk(nC), 256, 0, ks, vs, Cidx, nC, flag, pick);
    uint32_t* pos;
    if ((rc = compact_positions(ctx, arena, flag, nC, pos, n))) return rc;
    JX_ALLOC(x0, int32_t, n, false);
    JX_ALLOC(y0, int32_t, n, false);
    JX_ALLOC(s0, float, n, false);
    JX_ALLOC(c0, uint64_t, n, false);
    JX_ALLOC(id0, uint32_t, n, false);
    JX_LAUNCH(k_scatter_pts, nblk(nC), 256, 0, ks, pick, flag, pos, nC, sb, C, Q.chr_lex, S.chr_lex, (int64_t)S.n_chr, x0, y0,
              s0, c0, id0);
    // stable sort by chromosome-pair id keeps (qi, si) order inside each pair
    uint64_t* cs = c0; uint32_t* perm = id0;
    const uint64_t maxcp = (uint64_t)std::max<int64_t>(Q.n_chr, 1) * (uint64_t)std::max<int64_t>(S.n_chr, 1);
    if ((rc = jx_sort_pairs(ctx, arena, cs, perm, n, 0, jx_bits(maxcp)))) return rc;
    JX_ALLOC(x1, int32_t, n, false);
    JX_ALLOC(y1, int32_t, n, false);
    JX_ALLOC(s1, float, n, false);
  JX_TRACE("scan: dedupe+sorts");
    JX_LAUNCH(k_permute<int32_t>, nblk(n), 256, 0, x0, perm, n, x1);
    JX_LAUNCH(k_permute<int32_t>, nblk(n), 256, 0, y0, perm, n, y1);
    JX_LAUNCH(k_permute<float>, nblk(n), 256, 0, s0, perm, n, s1);
    px = x1; py = y1; psc = s1; cp = cs;
  }
  out->stats[1] = n;
  return scan_sorted_points(ctx, arena, px, py, psc, cp, nullptr, n, ctx->is_self ? 1 : 0, o, out, dev, extra_flag);
}

__global__ void k_filtered_to_recs(const int32_t* q, const int32_t* s, const float* score, uint32_t n, int is_self,
                                   const int32_t* qchr, const int32_t* schr, uint4* recs, uint32_t* bad) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int32_t qi = q[t], si = s[t];
  float r;
  if (!jx::round3g_f32(score[t], r)) { atomicExch(bad, 1u); r = score[t]; }   // "%.3g" text round trip
  uint4 rec = make_uint4((uint32_t)qi, (uint32_t)si, __float_as_uint(r), 0u);
  if (is_self) {                                   // read_blast, synteny.py:274-283
    if (qi > si) { rec.x = (uint32_t)si; rec.y = (uint32_t)qi; int32_t tmp = qi; qi = si; si = tmp; }
    if (qchr[qi] == schr[si] && si - qi < 40) rec.x = JX_NOREC;
  }
  recs[t] = rec;
}

}  // namespace

extern "C" {

void jx_scan_result_free(jx_scan_result* r) {
  if (!r) return;
  free(r->q_rank); free(r->s_rank); free(r->score); free(r->block_start); free(r->index);
  memset(r, 0, sizeof *r);
}

int jx_scan_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                 const jx_scan_opts* opts, jx_scan_result* out) {
  if (!ctx || !opts || !out) return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  TokMode mode;
  mode.strip = strip_names;
  mode.neardiag = ctx->is_self ? 1 : 0;
  TokOut tok;
  int rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok);
  if (rc) return rc;
  out->stats[0] = (int64_t)(tok.stats[0] - tok.stats[1]);
  out->stats[3] = (int64_t)tok.stats[1];
  return scan_records(ctx, arena, tok.recs, tok.n_lines, (size_t)tok.stats[2], opts, out);
}

int jx_scan_filtered(jx_ctx* ctx, const jx_filter_result* f, const jx_scan_opts* opts, jx_scan_result* out) {
  if (!ctx || !f || !opts || !out) return JX_EARG;
  memset(out, 0, sizeof *out);
  if (!ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t n = (uint32_t)f->n;
  if (n && f->device_token && f->device_token == ctx->lastf.token && (int64_t)n == ctx->lastf.n)
    return jx_scan_survivors(ctx, arena, opts, out, nullptr);          // still resident: no round trip
  JX_ALLOC(dq, int32_t, n, false);
  JX_ALLOC(ds, int32_t, n, false);
  JX_ALLOC(dsc, float, n, false);
  JX_ALLOC(recs, uint4, n, false);
  JX_ALLOC(bad, uint32_t, 1, true);
  if (n) {
    if (!f->q_rank || !f->s_rank || !f->score) { ctx->err = "filter result has no host arrays and is no longer resident"; return JX_EARG; }
    JX_CK(cudaMemcpyAsync(dq, f->q_rank, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(ds, f->s_rank, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dsc, f->score, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_LAUNCH(k_filtered_to_recs, nblk(n), 256, 0, dq, ds, dsc, n, ctx->is_self ? 1 : 0, ctx->side[0].chr_lex,
              ctx->side[1].chr_lex, recs, bad);
  }
  out->stats[0] = n;
  // host arrays may be anything: the general route (select valid, pair de-duplication, stable pair sort)
  return scan_records(ctx, arena, recs, n, n, opts, out, false, nullptr, bad);
}

int jx_scan_points(jx_ctx* ctx, const int32_t* qi, const int32_t* si, const float* score, const int64_t* group,
                   int64_t n64, int is_self, const jx_scan_opts* opts, jx_scan_result* out) {
  if (!ctx || !opts || !out || n64 < 0 || n64 >= (1ll << 31)) return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t n = (uint32_t)n64;
  int rc;
  // host-side key build is avoided: sort by (qi,si) then stably by group on the device
  std::vector<uint64_t> hk(n), hg(n);
  int64_t maxq = 0, maxs = 0, maxg = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (qi[i] < 0 || si[i] < 0 || (group && group[i] < 0)) { ctx->err = "negative coordinate or group"; return JX_EARG; }
    maxq = std::max<int64_t>(maxq, qi[i]); maxs = std::max<int64_t>(maxs, si[i]);
    if (group) maxg = std::max<int64_t>(maxg, group[i]);
  }
  const int sb = jx_bits((uint64_t)maxs + 1), qb = jx_bits((uint64_t)maxq + 1);
  for (uint32_t i = 0; i < n; ++i) { hk[i] = ((uint64_t)qi[i] << sb) | (uint32_t)si[i]; hg[i] = group ? (uint64_t)group[i] : 0; }
  JX_ALLOC(keys, uint64_t, n, false);
  JX_ALLOC(vals, uint32_t, n, false);
  JX_ALLOC(g0, uint64_t, n, false);
  JX_ALLOC(dx, int32_t, n, false);
  JX_ALLOC(dy, int32_t, n, false);
  JX_ALLOC(dsc, float, n, false);
  int32_t *px = nullptr, *py = nullptr; float* psc = nullptr; uint64_t* cp = nullptr; uint32_t* pidx = nullptr;
  if (n) {
    JX_CK(cudaMemcpyAsync(keys, hk.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(g0, hg.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dx, qi, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dy, si, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dsc, score, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_LAUNCH(k_iota, nblk(n), 256, 0, vals, n);
    uint64_t* ks = keys; uint32_t* vs = vals;
    if ((rc = jx_sort_pairs(ctx, arena, ks, vs, n, 0, sb + qb))) return rc;
    JX_ALLOC(g1, uint64_t, n, false);
    JX_LAUNCH(k_permute<uint64_t>, nblk(n), 256, 0, g0, vs, n, g1);
    uint64_t* gs = g1; uint32_t* perm = vs;
    if ((rc = jx_sort_pairs(ctx, arena, gs, perm, n, 0, jx_bits((uint64_t)maxg + 1)))) return rc;
    JX_ALLOC(x1, int32_t, n, false);
    JX_ALLOC(y1, int32_t, n, false);
    JX_ALLOC(s1, float, n, false);
    JX_LAUNCH(k_permute<int32_t>, nblk(n), 256, 0, dx, perm, n, x1);
    JX_LAUNCH(k_permute<int32_t>, nblk(n), 256, 0, dy, perm, n, y1);
    JX_LAUNCH(k_permute<float>, nblk(n), 256, 0, dsc, perm, n, s1);
    px = x1; py = y1; psc = s1; cp = gs; pidx = perm;
  }
  out->stats[1] = n;
  return scan_sorted_points(ctx, arena, px, py, psc, cp, pidx, n, is_self, opts, out, nullptr, nullptr);
}

}  // extern "C"

namespace {
__global__ void k_cands_to_recs(const uint4* __restrict__ F, uint32_t n, int is_self, const int32_t* qchr, const int32_t* schr,
                                uint4* recs, uint32_t* bad) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint4 c = F[t];
  int32_t qi = (int32_t)c.x, si = (int32_t)c.y;
  float r;
  if (!jx::round3g_f32(__uint_as_float(c.z), r)) { atomicExch(bad, 1u); r = __uint_as_float(c.z); }   // "%.3g" text round trip
  uint4 rec = make_uint4((uint32_t)qi, (uint32_t)si, __float_as_uint(r), 0u);
  if (is_self) {                                   // read_blast, synteny.py:274-283
    if (qi > si) { rec.x = (uint32_t)si; rec.y = (uint32_t)qi; int32_t tmp = qi; qi = si; si = tmp; }
    if (qchr[qi] == schr[si] && si - qi < 40) rec.x = JX_NOREC;
  }
  recs[t] = rec;
}
}  // namespace

// scan of a set of survivors given as device candidates {qi, si, score bits, ..} (a rank's share of a sharded pair, or
// all of them): the same hand-over as jx_scan_filtered
namespace {
__global__ void k_pack_anchor_cols(const int32_t* q, const int32_t* s, const float* sc, const uint32_t* blk, uint32_t m, int32_t* out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  out[4 * (size_t)t] = q[t]; out[4 * (size_t)t + 1] = s[t]; out[4 * (size_t)t + 2] = __float_as_int(sc[t]); out[4 * (size_t)t + 3] = (int32_t)blk[t];
}
}  // namespace
// dev_rows != NULL: the anchors are ALSO left on the device as int32 rows {qi, si, score bits, block} (session memory,
// see jx_shard_reset); host_result = 0 then skips the host arrays of `out` (n_anchors is still set).
extern "C" int jx_scan_cands(jx_ctx* ctx, uint64_t surv_dev, int64_t n64, const jx_scan_opts* opts, jx_scan_result* out, uint64_t* dev_rows,
                             int host_result) {
  if (!ctx || !opts || !out || n64 < 0 || n64 >= (1ll << 31)) return JX_EARG;
  memset(out, 0, sizeof *out);
  if (dev_rows) *dev_rows = 0;
  if (!ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t n = (uint32_t)n64;
  JX_ALLOC(recs, uint4, (size_t)n + 1, false);
  JX_ALLOC(bad, uint32_t, 1, true);
  if (n) JX_LAUNCH(k_cands_to_recs, nblk(n), 256, 0, (const uint4*)(uintptr_t)surv_dev, n, ctx->is_self ? 1 : 0, ctx->side[0].chr_lex,
                   ctx->side[1].chr_lex, recs, bad);
  out->stats[0] = n;
  ScanDev dev;
  dev.no_host = dev_rows && !host_result;
  int rc = scan_records(ctx, arena, recs, n, n, opts, out, !ctx->is_self, dev_rows ? &dev : nullptr, bad);
  if (rc) return rc;
  if (dev_rows && dev.n) {
    JX_SALLOC(rows, int32_t, (size_t)dev.n * 4, false);
    JX_LAUNCH(k_pack_anchor_cols, nblk(dev.n), 256, 0, dev.q, dev.s, dev.score, dev.block, dev.n, rows);
    JX_CK(cudaStreamSynchronize(ctx->stream));
    *dev_rows = (uint64_t)(uintptr_t)rows;
  }
  return JX_OK;
}

// the in-memory hand-over of the LAST jx_blastfilter's survivors (still on the device) to scan: equivalent to writing
// `.filtered` with "%.3g" scores (cblast.pyx:17) and reading it back with read_blast
int jx_scan_survivors(jx_ctx* ctx, Arena& arena, const jx_scan_opts* opts, jx_scan_result* out, ScanDev* dev) {
  const uint32_t n = (uint32_t)ctx->lastf.n;
  JX_ALLOC(recs, uint4, (size_t)n + 1, false);
  JX_ALLOC(bad, uint32_t, 1, true);
  if (n) JX_LAUNCH(k_filtered_to_recs, nblk(n), 256, 0, ctx->lastf.q, ctx->lastf.s, ctx->lastf.score, n, ctx->is_self ? 1 : 0,
                   ctx->side[0].chr_lex, ctx->side[1].chr_lex, recs, bad);
  out->stats[0] = n;
  // the library's own survivors are valid and pair-unique (blastfilter de-duplicates twice); a self comparison
  // re-orients pairs (qi <= si) and can make duplicates
  return scan_records(ctx, arena, recs, n, n, opts, out, !ctx->is_self, dev, bad);
}

// =====================================================================================================
// synteny mcscan (src/jcvi/compara/synteny.py:1538-1652): stack the anchor blocks on a
// reference BED.  Chaining = range_chain (src/jcvi/utils/range.py:412-463: end points sorted by
// (position, LEFT before RIGHT, index, score), one DP cell per end point, strict `>`), repeated on the blocks
// that are left (synteny.py:1601-1620) -- sequential by nature and O(blocks) per chain: host.  The
// per-gene track table (synteny.py:1622-1637: for every gene and track the partner in the FIRST block of
// the track, in chain order, that holds the gene) is one atomicMin per (block, gene) entry on the device.
// =====================================================================================================
namespace {
__global__ void k_track_fill(const int32_t* __restrict__ pair_gene, const int32_t* __restrict__ pair_block, uint32_t n_pairs,
                             const int32_t* __restrict__ track_of, const uint32_t* __restrict__ pos_of, uint32_t n_tracks,
                             unsigned long long* cell) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int32_t t = track_of[pair_block[p]];
  if (t < 0) return;
  const unsigned long long key = ((unsigned long long)pos_of[pair_block[p]] << 32) | p;   // first block of the track wins
  atomicMin(cell + (size_t)pair_gene[p] * n_tracks + (uint32_t)t, key);
}
__global__ void k_track_partner(const unsigned long long* __restrict__ cell, size_t n, const int32_t* __restrict__ pair_partner,
                                int32_t* table) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long c = cell[i];
  table[i] = c == ~0ull ? -1 : pair_partner[(uint32_t)c];
}

struct McEnd { double pos; int lr; int i; long long score; };
// range_chain on `cur` (indices into the caller's ranges); chain = positions in `cur`, in chain order
long long mc_chain(const std::vector<int>& cur, const double* start, const double* end, const int64_t* score, std::vector<int>& chain) {
  std::vector<McEnd> ends;
  ends.reserve(2 * cur.size());
  for (size_t i = 0; i < cur.size(); ++i) {
    ends.push_back({start[cur[i]], 0, (int)i, (long long)score[cur[i]]});
    ends.push_back({end[cur[i]], 1, (int)i, (long long)score[cur[i]]});
  }
  std::sort(ends.begin(), ends.end(), [](const McEnd& a, const McEnd& b) {
    if (a.pos != b.pos) return a.pos < b.pos;
    if (a.lr != b.lr) return a.lr < b.lr;
    if (a.i != b.i) return a.i < b.i;
    return a.score < b.score;
  });
  struct Cell { long long score; int from, which; };
  std::vector<Cell> sc;
  sc.reserve(ends.size());
  std::vector<int> left(cur.size(), -1);
  for (size_t i = 0; i < ends.size(); ++i) {
    Cell c = i == 0 ? Cell{0, -1, -1} : sc.back();
    const McEnd& e = ends[i];
    if (e.lr == 0) left[(size_t)e.i] = (int)i;
    else {
      const int lj = left[(size_t)e.i];
      const long long cs = sc[(size_t)lj].score + e.score;
      if (cs > c.score) c = Cell{cs, lj, e.i};
    }
    sc.push_back(c);
  }
  chain.clear();
  int last = sc.back().from, which = sc.back().which;
  while (last != -1) {
    if (which != -1) chain.push_back(which);
    which = sc[(size_t)last].which;
    last = sc[(size_t)last].from;
  }
  std::reverse(chain.begin(), chain.end());
  return sc.back().score;
}
}  // namespace

extern "C" int jx_mcscan(jx_ctx* ctx, const double* start, const double* end, const int64_t* score, const int32_t* block_id,
                         int64_t n_ranges, int32_t max_iter, const int32_t* pair_gene, const int32_t* pair_partner,
                         const int32_t* pair_block, int64_t n_pairs, int32_t n_genes, int32_t n_blocks, int32_t* n_tracks,
                         int32_t** track_len, int32_t** track_blocks, int64_t** track_score, int32_t** table) {
  if (!ctx || !n_tracks || !track_len || !track_blocks || !track_score || !table || n_ranges < 0 || n_pairs < 0 ||
      n_pairs >= (1ll << 32) || n_genes < 0 || n_blocks < 0) return JX_EARG;
  *n_tracks = 0; *track_len = nullptr; *track_blocks = nullptr; *track_score = nullptr; *table = nullptr;
  for (int64_t i = 0; i < n_ranges; ++i)
    if (block_id[i] < 0 || block_id[i] >= n_blocks) { ctx->err = "jx_mcscan: block id out of range"; return JX_EARG; }
  for (int64_t i = 0; i < n_pairs; ++i)
    if (pair_gene[i] < 0 || pair_gene[i] >= n_genes || pair_block[i] < 0 || pair_block[i] >= n_blocks) {
      ctx->err = "jx_mcscan: pair entry out of range"; return JX_EARG;
    }
  cudaSetDevice(ctx->device);
  // ---- chains (host) -----------------------------------------------------------------------------------------
  std::vector<int> cur((size_t)n_ranges);
  for (int64_t i = 0; i < n_ranges; ++i) cur[(size_t)i] = (int)i;
  std::vector<int32_t> tlen, tblocks;
  std::vector<int64_t> tscore;
  std::vector<int32_t> track_of((size_t)std::max(n_blocks, 1), -1);
  std::vector<uint32_t> pos_of((size_t)std::max(n_blocks, 1), 0u);
  std::vector<char> taken((size_t)std::max(n_blocks, 1), 0);
  int iteration = 0;
  while (!cur.empty() && iteration < max_iter) {
    std::vector<int> chain;
    const long long sc = mc_chain(cur, start, end, score, chain);
    uint32_t pos = 0;
    for (int x : chain) {
      const int32_t b = block_id[cur[(size_t)x]];
      tblocks.push_back(b);
      if (!taken[(size_t)b]) { taken[(size_t)b] = 1; track_of[(size_t)b] = iteration; pos_of[(size_t)b] = pos; }
      ++pos;
    }
    tlen.push_back((int32_t)chain.size());
    tscore.push_back(sc);
    std::vector<int> rest;
    rest.reserve(cur.size());
    for (int r : cur) if (!taken[(size_t)block_id[r]]) rest.push_back(r);
    cur.swap(rest);
    ++iteration;
  }
  const uint32_t T = (uint32_t)iteration;
  *n_tracks = (int32_t)T;
  auto dup = [](const void* src, size_t bytes) { void* p = malloc(bytes ? bytes : 1); if (p && bytes) memcpy(p, src, bytes); return p; };
  *track_len = (int32_t*)dup(tlen.data(), tlen.size() * 4);
  *track_blocks = (int32_t*)dup(tblocks.data(), tblocks.size() * 4);
  *track_score = (int64_t*)dup(tscore.data(), tscore.size() * 8);
  const size_t cells = (size_t)n_genes * T;
  *table = (int32_t*)malloc(cells ? cells * 4 : 4);
  if (!*track_len || !*track_blocks || !*track_score || !*table) { ctx->err = "out of host memory"; return JX_EINTERNAL; }
  if (!cells) return JX_OK;
  // ---- track table (device) ------------------------------------------------------------------------------------
  Arena arena(ctx);
  JX_ALLOC(d_cell, unsigned long long, cells, false);
  JX_CK(cudaMemsetAsync(d_cell, 0xFF, cells * 8, ctx->stream));
  JX_ALLOC(d_table, int32_t, cells, false);
  JX_ALLOC(d_pg, int32_t, (size_t)std::max<int64_t>(n_pairs, 1), false);
  JX_ALLOC(d_pp, int32_t, (size_t)std::max<int64_t>(n_pairs, 1), false);
  JX_ALLOC(d_pb, int32_t, (size_t)std::max<int64_t>(n_pairs, 1), false);
  JX_ALLOC(d_to, int32_t, track_of.size(), false);
  JX_ALLOC(d_po, uint32_t, pos_of.size(), false);
  if (n_pairs) {
    JX_CK(cudaMemcpyAsync(d_pg, pair_gene, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(d_pp, pair_partner, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(d_pb, pair_block, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  JX_CK(cudaMemcpyAsync(d_to, track_of.data(), track_of.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  JX_CK(cudaMemcpyAsync(d_po, pos_of.data(), pos_of.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n_pairs) JX_LAUNCH(k_track_fill, nblk((size_t)n_pairs), 256, 0, d_pg, d_pb, (uint32_t)n_pairs, d_to, d_po, T, d_cell);
  JX_LAUNCH(k_track_partner, nblk(cells), 256, 0, d_cell, cells, d_pp, d_table);
  JX_CK(cudaMemcpyAsync(*table, d_table, cells * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  return JX_OK;
}

// =====================================================================================================
// synteny simple (src/jcvi/compara/synteny.py:1137-1316): per block of an anchors file the
// first / last gene rank on both sides, the number of distinct genes, the size, and the sums the
// orientation needs (get_orientation, synteny.py:222-236: the sign of the least-squares slope = the sign
// of the covariance n*sum(xy) - sum(x)*sum(y), accumulated exactly on ranks centred at the block minimum).
// One segmented reduction over the anchors: atomics for min / max / count / sums, and a sort of
// (block, rank) keys per side for the distinct counts.
// =====================================================================================================
namespace {
__global__ void k_block_minmax(const int32_t* __restrict__ blk, const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                               uint32_t n, int32_t* xmin, int32_t* xmax, int32_t* ymin, int32_t* ymax, uint32_t* size) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t b = blk[i];
  atomicMin(xmin + b, x[i]); atomicMax(xmax + b, x[i]);
  atomicMin(ymin + b, y[i]); atomicMax(ymax + b, y[i]);
  atomicAdd(size + b, 1u);
}
__global__ void k_block_sums(const int32_t* __restrict__ blk, const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                             uint32_t n, const int32_t* __restrict__ xmin, const int32_t* __restrict__ ymin,
                             unsigned long long* sx, unsigned long long* sy, unsigned long long* sxy) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t b = blk[i];
  const unsigned long long cx = (unsigned long long)(x[i] - xmin[b]), cy = (unsigned long long)(y[i] - ymin[b]);
  atomicAdd(sx + b, cx); atomicAdd(sy + b, cy); atomicAdd(sxy + b, cx * cy);
}
__global__ void k_block_rank_keys(const int32_t* __restrict__ blk, const int32_t* __restrict__ r, uint32_t n, uint64_t* keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = ((uint64_t)(uint32_t)blk[i] << 32) | (uint32_t)r[i];
}
__global__ void k_block_distinct(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* distinct) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || keys[i - 1] != keys[i]) atomicAdd(distinct + (uint32_t)(keys[i] >> 32), 1u);
}
}  // namespace

extern "C" int jx_block_stats(jx_ctx* ctx, const int32_t* block, const int32_t* qrank, const int32_t* srank, int64_t n,
                              int32_t n_blocks, int32_t* qmin, int32_t* qmax, int32_t* smin, int32_t* smax, uint32_t* size,
                              uint32_t* distinct_q, uint32_t* distinct_s, uint64_t* sum_q, uint64_t* sum_s, uint64_t* sum_qs) {
  if (!ctx || n < 0 || n >= (1ll << 31) || n_blocks < 0 || (n && (!block || !qrank || !srank)) || !qmin || !qmax || !smin ||
      !smax || !size || !distinct_q || !distinct_s || !sum_q || !sum_s || !sum_qs) return JX_EARG;
  for (int64_t i = 0; i < n; ++i)
    if (block[i] < 0 || block[i] >= n_blocks || qrank[i] < 0 || srank[i] < 0) { ctx->err = "jx_block_stats: entry out of range"; return JX_EARG; }
  if (!n_blocks) return JX_OK;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  int rc;
  const size_t nb = (size_t)n_blocks, nn = (size_t)std::max<int64_t>(n, 1);
  JX_ALLOC(d_b, int32_t, nn, false);
  JX_ALLOC(d_x, int32_t, nn, false);
  JX_ALLOC(d_y, int32_t, nn, false);
  JX_ALLOC(d_i32, int32_t, 4 * nb, false);          // xmin xmax ymin ymax
  JX_ALLOC(d_u32, uint32_t, 3 * nb, true);          // size distinct_q distinct_s
  JX_ALLOC(d_u64, unsigned long long, 3 * nb, true);
  int32_t *xmin = d_i32, *xmax = d_i32 + nb, *ymin = d_i32 + 2 * nb, *ymax = d_i32 + 3 * nb;
  JX_CK(cudaMemsetAsync(xmin, 0x7F, nb * 4, ctx->stream));     // 0x7F7F7F7F: above every rank (< 2^31 - 2 genes per BED)
  JX_CK(cudaMemsetAsync(ymin, 0x7F, nb * 4, ctx->stream));
  JX_CK(cudaMemsetAsync(xmax, 0xFF, nb * 4, ctx->stream));     // -1
  JX_CK(cudaMemsetAsync(ymax, 0xFF, nb * 4, ctx->stream));
  if (n) {
    JX_CK(cudaMemcpyAsync(d_b, block, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(d_x, qrank, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(d_y, srank, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t un = (uint32_t)n;
    JX_LAUNCH(k_block_minmax, nblk(un), 256, 0, d_b, d_x, d_y, un, xmin, xmax, ymin, ymax, d_u32);
    JX_LAUNCH(k_block_sums, nblk(un), 256, 0, d_b, d_x, d_y, un, xmin, ymin, d_u64, d_u64 + nb, d_u64 + 2 * nb);
    for (int side = 0; side < 2; ++side) {
      JX_ALLOC(keys, uint64_t, nn, false);
      JX_LAUNCH(k_block_rank_keys, nblk(un), 256, 0, d_b, side ? d_y : d_x, un, keys);
      uint64_t* ks = keys;
      if ((rc = jx_sort_keys(ctx, arena, ks, (size_t)n, 0, 64))) return rc;
      JX_LAUNCH(k_block_distinct, nblk(un), 256, 0, ks, un, d_u32 + (size_t)(1 + side) * nb);
    }
  }
  JX_CK(cudaMemcpyAsync(qmin, xmin, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(qmax, xmax, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(smin, ymin, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(smax, ymax, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(size, d_u32, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(distinct_q, d_u32 + nb, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(distinct_s, d_u32 + 2 * nb, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(sum_q, d_u64, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(sum_s, d_u64 + nb, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(sum_qs, d_u64 + 2 * nb, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  return JX_OK;
}


// Stage 3: synteny liftover on the GPU (K11 of SURVEY.md section 2).
//
// Replaces src/jcvi/compara/synteny.py:
//   300-374  process_blast_for_liftover  file-order first-seen de-duplication of the raw hits
//   377-399  read_anchors                (qi,si) -> block, duplicate check
//   455-473  synteny_liftover            scipy cKDTree(anchors, leafsize=16).query(p=1, distance_upper_bound=dist)
//   1946-1962 append lifted pairs to the nearest anchor's block, per sorted chromosome pair, in file order
// and the lookup half of AnchorFile.filter_blocks (compara/base.py:52-83).
//
// Nearest anchor.  Anchors are sorted by (qi, si) with a row pointer per query rank, so the
// candidates of a hit are one contiguous slice (rows qi-dist+1 .. qi+dist-1, clipped to the
// hit's chromosome); the kernel takes the exact minimum L1 distance.  Which BLOCK the hit joins
// only depends on cKDTree's traversal order when several anchors at that minimum distance
// belong to different blocks; exactly those hits are re-run through a bit-exact emulation of
// scipy's tree (host build with std::nth_element -- its permutation defines the leaf order --
// and a device walk with scipy's near-first descent and binary-heap tie rules).
#include <algorithm>
#include <map>

#include <thread>
#include "jx_internal.cuh"

namespace {

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }



__global__ void k_rowptr(const uint64_t* akey, uint32_t na, int sb, uint32_t G, uint32_t* rowptr) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  uint64_t key = (uint64_t)g << sb;
  uint32_t lo = 0, hi = na;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (akey[mid] < key) lo = mid + 1; else hi = mid; }
  rowptr[g] = lo;
}

// occupancy bitmap over (q cell, s cell): a set bit means "an anchor may be within dist"
__global__ void k_mark_cells(const uint64_t* akey, uint32_t na, int sb, int dist, int cshift, uint32_t nsc, uint32_t nqc,
                             uint32_t* bitmap) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na) return;
  const uint64_t a = akey[t];
  const int32_t x = (int32_t)(a >> sb), y = (int32_t)(a & ((1ull << sb) - 1));
  const int32_t x0 = max(x - dist, 0) >> cshift, x1 = min((uint32_t)((x + dist) >> cshift), nqc - 1);
  const int32_t y0 = max(y - dist, 0) >> cshift, y1 = min((uint32_t)((y + dist) >> cshift), nsc - 1);
  for (int32_t cx = x0; cx <= x1; ++cx)
    for (int32_t cy = y0; cy <= y1; ++cy) {
      const uint64_t bit = (uint64_t)cx * nsc + (uint64_t)cy;
      atomicOr(bitmap + (bit >> 5), 1u << (bit & 31));
    }
}

// is some anchor of the same chromosome pair at L1 distance < dist (0 included)?
__device__ __forceinline__ bool anchor_within(int32_t x, int32_t y, int sb, const uint64_t* __restrict__ akey,
                                              const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ qbeg,
                                              const int32_t* __restrict__ qend, const int32_t* __restrict__ schr, int dist) {
  const uint64_t smask = (1ull << sb) - 1;
  const int32_t lo = max(x - (dist - 1), qbeg[x]), hi = min(x + dist - 1, qend[x] - 1);
  if (hi < lo) return false;
  const uint32_t t0 = rowptr[lo], t1 = rowptr[hi + 1];
  const int32_t cy = schr[y];
  for (uint32_t u = t0; u < t1; ++u) {
    const uint64_t a = akey[u];
    const int32_t ax = (int32_t)(a >> sb), ay = (int32_t)(a & smask);
    if (abs(ax - x) + abs(ay - y) < dist && schr[ay] == cy) return true;
  }
  return false;
}

// raw records -> the few that can matter to liftover: hits within dist of an anchor (the
// anchors themselves included: AnchorFile.filter_blocks needs their raw scores).
// Pass 1 (k_select_cell) only tests the occupancy bitmap -- uniform work, 4 records in flight per
// thread; pass 2 (k_keep_near) runs the exact window test on the compacted candidates so that its
// loop executes with full warps (one fused kernel ran with 5 of 32 lanes active).
__global__ void __launch_bounds__(256) k_select_cell(const uint4* __restrict__ recs, uint32_t n, const uint32_t* __restrict__ bitmap,
                                                     int cshift, uint32_t nsc, uint4* out, uint32_t* out_idx, uint32_t* counter) {
#ifndef JX_SEL_U
#define JX_SEL_U 4
#endif
  constexpr int U = JX_SEL_U;
  __shared__ uint32_t s_cnt[256 / 32 + 1];
  const uint32_t stride = gridDim.x * blockDim.x * U;
  const uint32_t lane = threadIdx.x & 31u, woff = (threadIdx.x & ~31u) * U;
  // the trip count is uniform over the CTA (block_append_multi synchronises)
  for (uint64_t c0 = (uint64_t)blockIdx.x * blockDim.x * U; c0 < n; c0 += stride) {
    const uint32_t i0 = (uint32_t)c0 + woff;
    uint4 r[U]; bool keep[U]; uint32_t w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      r[u] = i < n ? __ldcs(recs + i) : make_uint4(JX_NOREC, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      keep[u] = r[u].x != JX_NOREC;
      const uint64_t bit = (uint64_t)(r[u].x >> cshift) * nsc + (uint64_t)(r[u].y >> cshift);
      w[u] = keep[u] ? __ldg(bitmap + (bit >> 5)) >> (bit & 31) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) keep[u] = keep[u] && (w[u] & 1u);
    uint32_t slot[U];
    block_append_multi<U>(keep, counter, slot, s_cnt);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (keep[u]) { out[slot[u]] = r[u]; out_idx[slot[u]] = i0 + u * 32u + lane; }
  }
}
__global__ void k_keep_near(const uint4* __restrict__ cand, const uint32_t* __restrict__ cand_idx, uint32_t n, int sb,
                            const uint64_t* __restrict__ akey, const uint32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ qbeg, const int32_t* __restrict__ qend,
                            const int32_t* __restrict__ schr, int dist, uint4* out, uint32_t* out_idx, uint32_t* counter) {
  const uint32_t nround = (n + 31u) & ~31u;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nround) return;
  uint4 r = make_uint4(JX_NOREC, 0, 0, 0);
  bool keep = false;
  if (t < n) {
    r = cand[t];
    keep = anchor_within((int32_t)r.x, (int32_t)r.y, sb, akey, rowptr, qbeg, qend, schr, dist);
  }
  const uint32_t slot = warp_append(keep, counter);
  if (keep) { out[slot] = r; out_idx[slot] = cand_idx[t]; }
}


// ---- scipy cKDTree emulation ------------------------------------------------------------------
struct KNode { long long split; int split_dim, lesser, greater, start, end, pad; };

struct HostTrees {
  std::vector<KNode> nodes;
  std::vector<int> indices;           // per tree: permutation of local point ids, global offset applied
  std::vector<long long> px, py;      // concatenated points (tree-local order = input order)
  std::vector<int> root, ptbase;
  std::vector<long long> bounds;      // mins0, mins1, maxes0, maxes1 per tree

  int build(int base, int start, int end) {      // [start, end) are global positions in `indices`
    int me = (int)nodes.size();
    nodes.push_back(KNode{0, -1, -1, -1, start, end, 0});
    if (end - start <= 16) return me;
    auto cx = [&](int id, int d) { return d == 0 ? px[(size_t)base + id] : py[(size_t)base + id]; };
    long long mx[2], mn[2];
    for (int d = 0; d < 2; ++d) mx[d] = mn[d] = cx(indices[(size_t)start], d);
    for (int i = start + 1; i < end; ++i)
      for (int d = 0; d < 2; ++d) { long long v = cx(indices[(size_t)i], d); if (v > mx[d]) mx[d] = v; if (v < mn[d]) mn[d] = v; }
    int d = 0; long long size = 0;
    for (int i = 0; i < 2; ++i) if (mx[i] - mn[i] > size) { d = i; size = mx[i] - mn[i]; }
    if (mx[d] == mn[d]) return me;
    const int npts = end - start;
    std::nth_element(indices.begin() + start, indices.begin() + start + npts / 2, indices.begin() + end,
                     [&](int a, int b) { return cx(a, d) < cx(b, d); });
    long long split = cx(indices[(size_t)(start + npts / 2)], d);
    int p = start, q = end - 1;
    while (p <= q) {
      if (cx(indices[(size_t)p], d) < split) ++p;
      else if (cx(indices[(size_t)q], d) >= split) --q;
      else { std::swap(indices[(size_t)p], indices[(size_t)q]); ++p; --q; }
    }
    if (p == start) {
      int j = start; split = cx(indices[(size_t)j], d);
      for (int i = start + 1; i < end; ++i) if (cx(indices[(size_t)i], d) < split) { j = i; split = cx(indices[(size_t)j], d); }
      std::swap(indices[(size_t)start], indices[(size_t)j]);
      p = start + 1;
    } else if (p == end) {
      int j = end - 1; split = cx(indices[(size_t)j], d);
      for (int i = start; i < end - 1; ++i) if (cx(indices[(size_t)i], d) > split) { j = i; split = cx(indices[(size_t)j], d); }
      std::swap(indices[(size_t)(end - 1)], indices[(size_t)j]);
      p = end - 1;
    }
    nodes[(size_t)me].split_dim = d;
    nodes[(size_t)me].split = split;
    int l = build(base, start, p);
    int g = build(base, p, end);
    nodes[(size_t)me].lesser = l;
    nodes[(size_t)me].greater = g;
    return me;
  }
  // points of one tree in input order; returns the tree id
  int add_tree(const std::vector<std::pair<long long, long long>>& pts) {
    const int base = (int)px.size(), n = (int)pts.size();
    long long mn0 = 0, mn1 = 0, mx0 = 0, mx1 = 0;
    for (int i = 0; i < n; ++i) {
      px.push_back(pts[(size_t)i].first); py.push_back(pts[(size_t)i].second);
      indices.push_back(i);
      if (i == 0) { mn0 = mx0 = pts[0].first; mn1 = mx1 = pts[0].second; }
      mn0 = std::min(mn0, pts[(size_t)i].first); mx0 = std::max(mx0, pts[(size_t)i].first);
      mn1 = std::min(mn1, pts[(size_t)i].second); mx1 = std::max(mx1, pts[(size_t)i].second);
    }
    ptbase.push_back(base);
    bounds.push_back(mn0); bounds.push_back(mn1); bounds.push_back(mx0); bounds.push_back(mx1);
    root.push_back(n ? build(base, base, base + n) : -1);
    return (int)root.size() - 1;
  }
};

constexpr int HEAP_CAP = 96;

// k = 1, p = 1, eps = 0 query of one point against its tree; all quantities are integers, so
// int64 arithmetic reproduces scipy's float64 arithmetic exactly.
__global__ void k_kdwalk(const long long* __restrict__ qx, const long long* __restrict__ qy, const int* __restrict__ qtree,
                         uint32_t n, const KNode* __restrict__ nodes, const int* __restrict__ indices,
                         const long long* __restrict__ px, const long long* __restrict__ py, const int* __restrict__ root,
                         const int* __restrict__ ptbase, const long long* __restrict__ bounds, long long bound,
                         int* out_local, long long* out_dist, uint32_t* overflow) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int tr = qtree[t];
  const long long x[2] = {qx[t], qy[t]};
  long long best = bound; int besti = -1;
  int cur = root[tr];
  if (cur < 0) { out_local[t] = -1; out_dist[t] = -1; return; }
  const int base = ptbase[tr];
  long long hp[HEAP_CAP], hs0[HEAP_CAP], hs1[HEAP_CAP]; int hn[HEAP_CAP];
  int hsize = 0;
  long long sd[2], prio = 0;
  for (int d = 0; d < 2; ++d) {
    long long mn = bounds[tr * 4 + d], mx = bounds[tr * 4 + 2 + d];
    long long s = max(0ll, max(mn - x[d], x[d] - mx));
    sd[d] = s; prio += s;
  }
  for (;;) {
    const KNode nd = nodes[cur];
    if (nd.split_dim == -1) {
      for (int i = nd.start; i < nd.end; ++i) {
        int id = indices[i];
        long long dd = llabs(px[base + id] - x[0]) + llabs(py[base + id] - x[1]);
        if (dd < best) { best = dd; besti = id; }
      }
      if (hsize == 0) break;
      // pop: move last to the root, sift down (swap only when the parent is strictly larger)
      cur = hn[0]; prio = hp[0]; sd[0] = hs0[0]; sd[1] = hs1[0];
      --hsize;
      hp[0] = hp[hsize]; hn[0] = hn[hsize]; hs0[0] = hs0[hsize]; hs1[0] = hs1[hsize];
      int i = 0;
      for (;;) {
        int j = 2 * i + 1, k = 2 * i + 2, l = i;
        if (j < hsize && hp[i] > hp[j]) l = j;
        if (k < hsize && hp[l] > hp[k]) l = k;
        if (l == i) break;
        long long tp = hp[i]; hp[i] = hp[l]; hp[l] = tp;
        int tn = hn[i]; hn[i] = hn[l]; hn[l] = tn;
        long long a0 = hs0[i]; hs0[i] = hs0[l]; hs0[l] = a0;
        long long a1 = hs1[i]; hs1[i] = hs1[l]; hs1[l] = a1;
        i = l;
      }
    } else {
      if (prio > best) break;
      const int d = nd.split_dim;
      int nearn, farn;
      if (x[d] < nd.split) { nearn = nd.lesser; farn = nd.greater; } else { nearn = nd.greater; farn = nd.lesser; }
      const long long nsd = llabs(nd.split - x[d]);
      const long long fprio = prio + nsd - sd[d];
      if (fprio <= best) {
        if (hsize >= HEAP_CAP) { atomicExch(overflow, 1u); }
        else {
          int i = hsize++;
          hp[i] = fprio; hn[i] = farn; hs0[i] = d == 0 ? nsd : sd[0]; hs1[i] = d == 1 ? nsd : sd[1];
          while (i > 0) {
            int par = (i - 1) / 2;
            if (hp[i] < hp[par]) {
              long long tp = hp[i]; hp[i] = hp[par]; hp[par] = tp;
              int tn = hn[i]; hn[i] = hn[par]; hn[par] = tn;
              long long a0 = hs0[i]; hs0[i] = hs0[par]; hs0[par] = a0;
              long long a1 = hs1[i]; hs1[i] = hs1[par]; hs1[par] = a1;
              i = par;
            } else break;
          }
        }
      }
      cur = nearn;
    }
  }
  out_local[t] = besti;
  out_dist[t] = besti < 0 ? -1 : best;
}

int run_walk(jx_ctx* ctx, Arena& arena, const HostTrees& T, const std::vector<long long>& qx, const std::vector<long long>& qy,
             const std::vector<int>& qtree, long long bound, std::vector<int>& local, std::vector<long long>& dist) {
  const uint32_t n = (uint32_t)qx.size();
  local.assign(n, -1); dist.assign(n, -1);
  if (!n) return JX_OK;
  JX_ALLOC(dqx, long long, n, false);
  JX_ALLOC(dqy, long long, n, false);
  JX_ALLOC(dqt, int, n, false);
  JX_ALLOC(dnodes, KNode, T.nodes.size(), false);
  JX_ALLOC(dind, int, T.indices.size(), false);
  JX_ALLOC(dpx, long long, T.px.size(), false);
  JX_ALLOC(dpy, long long, T.py.size(), false);
  JX_ALLOC(droot, int, T.root.size(), false);
  JX_ALLOC(dbase, int, T.ptbase.size(), false);
  JX_ALLOC(dbounds, long long, T.bounds.size(), false);
  JX_ALLOC(dol, int, n, false);
  JX_ALLOC(dod, long long, n, false);
  JX_ALLOC(dovf, uint32_t, 1, true);
#define UP(dst, src) if (!(src).empty()) JX_CK(cudaMemcpyAsync(dst, (src).data(), (src).size() * sizeof((src)[0]), cudaMemcpyHostToDevice, ctx->stream))
  UP(dqx, qx); UP(dqy, qy); UP(dqt, qtree); UP(dnodes, T.nodes); UP(dind, T.indices); UP(dpx, T.px); UP(dpy, T.py);
  UP(droot, T.root); UP(dbase, T.ptbase); UP(dbounds, T.bounds);
#undef UP
  JX_LAUNCH(k_kdwalk, nblk(n, 64), 64, 0, dqx, dqy, dqt, n, dnodes, dind, dpx, dpy, droot, dbase, dbounds, bound, dol, dod, dovf);
  uint32_t ovf = 0;
  JX_CK(cudaMemcpyAsync(local.data(), dol, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(dist.data(), dod, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&ovf, dovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (ovf) { ctx->err = "KD-tree walk heap overflow"; return JX_EINTERNAL; }
  return JX_OK;
}

__global__ void k_patch_amb(const uint32_t* amb_hit, const uint32_t* amb_block, uint32_t n, uint32_t* hit_block) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) hit_block[amb_hit[t]] = amb_block[t];
}
__global__ void k_emit_lifted(const uint32_t* v, uint32_t n, const uint64_t* hkey, const float* hsc, const uint32_t* hit_block,
                              int sb, int32_t* oq, int32_t* os, float* osc, int32_t* ob) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t h = v[t];
  uint64_t k = hkey[h];
  oq[t] = (int32_t)(k >> sb); os[t] = (int32_t)(k & ((1ull << sb) - 1));
  osc[t] = hsc[h]; ob[t] = (int32_t)hit_block[h];
}


}  // namespace

extern "C" {

void jx_liftover_result_free(jx_liftover_result* r) {
  if (!r) return;
  free(r->q_rank); free(r->s_rank); free(r->score); free(r->block); free(r->accepted); free(r->raw_score);
  memset(r, 0, sizeof *r);
}

int jx_liftover_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                     const int32_t* a_qi, const int32_t* a_si, const int32_t* a_block, int64_t n_lines64,
                     jx_liftover_result* out) {
  if (!ctx || !out || n_lines64 < 0 || n_lines64 >= (1ll << 31) || dist < 0) return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t nal = (uint32_t)n_lines64;
  int rc;
  JX_TRACE("lo: enter");
  // ---- raw records: reuse the ones jx_blastfilter just produced from the same held text (the
  // reference parses the raw table twice: blastfilter.py:48 and synteny.py:1939), else tokenize
  const uint4* recs = nullptr; uint32_t n_lines = 0;
  unsigned long long tstats[12];
  {
    auto& c = ctx->cache;
    if (c.valid && c.plain && text_on_device && (const uint8_t*)text == c.text && nbytes == c.nbytes && (strip_names != 0) == (c.strip != 0) &&
        !ctx->is_self) {
      recs = c.recs; n_lines = c.n_lines;
      memcpy(tstats, c.stats, sizeof tstats);
      c.valid = false;                      // one-shot: the next pipeline pass tokenizes again
    } else {
      TokMode mode;
      mode.strip = strip_names;
      mode.neardiag = ctx->is_self ? 1 : 0;
      TokOut tok;
      if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
      recs = tok.recs; n_lines = tok.n_lines;
      memcpy(tstats, tok.stats, sizeof tstats);
    }
  }
  JX_ALLOC(daq, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(das, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(dab, uint32_t, (size_t)nal + 1, false);
  if (nal) {
    JX_CK(cudaMemcpyAsync(daq, a_qi, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(das, a_si, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dab, a_block, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  JX_TRACE("lo: records ready");
  return jx_liftover_core(ctx, arena, recs, n_lines, tstats, dist, daq, das, dab, nal, a_qi, a_si, a_block, out);
}

}  // extern "C"

namespace {
// anchor lines -> sort keys: (qi << sb | si) for anchors in the BED, one key above all others for the rest
__global__ void k_anchor_keys(const int32_t* __restrict__ aq, const int32_t* __restrict__ as_, uint32_t nal, int sb, int kb,
                              uint32_t Gq, uint32_t Gs, uint64_t* keys, uint32_t* order, uint32_t* counts /* 0 valid, 1 out of range */) {
  const uint32_t nround = (nal + 31u) & ~31u;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nround) return;
  bool valid = false;
  if (t < nal) {
    const int32_t q = aq[t], s = as_[t];
    valid = q >= 0 && s >= 0;
    if (valid && ((uint32_t)q >= Gq || (uint32_t)s >= Gs)) { atomicExch(counts + 1, 1u); valid = false; }
    keys[t] = valid ? (((uint64_t)(uint32_t)q << sb) | (uint32_t)s) : (1ull << kb);
    order[t] = t;
  }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, valid);
  if (m && (threadIdx.x & 31) == 0) atomicAdd(counts, (uint32_t)__popc(m));
}
__global__ void k_anchor_blocks2(const uint64_t* keys, const uint32_t* order, const uint32_t* block_in, const uint32_t* na,
                                 uint32_t* block_out, uint32_t* dup) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *na) return;
  block_out[i] = block_in[order[i]];
  if (i && keys[i] == keys[i - 1]) atomicExch(dup, 1u);
}
__global__ void k_rowptr2(const uint64_t* akey, const uint32_t* na_p, int sb, uint32_t G, uint32_t* rowptr) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  const uint32_t na = *na_p;
  uint64_t key = (uint64_t)g << sb;
  uint32_t lo = 0, hi = na;
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (akey[mid] < key) lo = mid + 1; else hi = mid; }
  rowptr[g] = lo;
}
__global__ void k_mark_cells2(const uint64_t* akey, const uint32_t* na, int sb, int dist, int cshift, uint32_t nsc, uint32_t nqc,
                              uint32_t* bitmap) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *na) return;
  const uint64_t a = akey[t];
  const int32_t x = (int32_t)(a >> sb), y = (int32_t)(a & ((1ull << sb) - 1));
  const int32_t x0 = max(x - dist, 0) >> cshift, x1 = min((uint32_t)((x + dist) >> cshift), nqc - 1);
  const int32_t y0 = max(y - dist, 0) >> cshift, y1 = min((uint32_t)((y + dist) >> cshift), nsc - 1);
  for (int32_t cx = x0; cx <= x1; ++cx)
    for (int32_t cy = y0; cy <= y1; ++cy) {
      const uint64_t bit = (uint64_t)cx * nsc + (uint64_t)cy;
      atomicOr(bitmap + (bit >> 5), 1u << (bit & 31));
    }
}
// pass 1 as k_select_cell, candidates in the compact form {qi, si, score bits, record index}
__global__ void __launch_bounds__(256) k_select_cell2(const uint4* __restrict__ recs, uint32_t n, const uint32_t* __restrict__ bitmap,
                                                      int cshift, uint32_t nsc, uint4* out, uint32_t* counter) {
  constexpr int U = 4;
  __shared__ uint32_t s_cnt[256 / 32 + 1];
  const uint32_t stride = gridDim.x * blockDim.x * U;
  const uint32_t lane = threadIdx.x & 31u, woff = (threadIdx.x & ~31u) * U;
  for (uint64_t c0 = (uint64_t)blockIdx.x * blockDim.x * U; c0 < n; c0 += stride) {
    const uint32_t i0 = (uint32_t)c0 + woff;
    uint4 r[U]; bool keep[U]; uint32_t w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      r[u] = i < n ? __ldcs(recs + i) : make_uint4(JX_NOREC, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      keep[u] = r[u].x != JX_NOREC;
      const uint64_t bit = (uint64_t)(r[u].x >> cshift) * nsc + (uint64_t)(r[u].y >> cshift);
      w[u] = keep[u] ? __ldg(bitmap + (bit >> 5)) >> (bit & 31) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) keep[u] = keep[u] && (w[u] & 1u);
    uint32_t slot[U];
    block_append_multi<U>(keep, counter, slot, s_cnt);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (keep[u]) { r[u].w = i0 + u * 32u + lane; out[slot[u]] = r[u]; }
  }
}
__device__ __forceinline__ uint32_t pair_hash2(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  return h;
}
// pass 2: exact window test; the candidates that are near an anchor enter the pair table, which keeps the FIRST
// occurrence in file order per (qi, si) (process_blast_for_liftover's `seen`, synteny.py:339-360): value = index << 32 | j
__global__ void k_near_insert(const uint4* __restrict__ cand, const uint32_t* __restrict__ n_p, int sb, const uint64_t* __restrict__ akey,
                              const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ qbeg, const int32_t* __restrict__ qend,
                              const int32_t* __restrict__ schr, int dist, unsigned long long* K, unsigned long long* V, uint32_t mask,
                              uint32_t* Cslot) {
  const uint32_t n = *n_p;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const uint4 r = cand[t];
    uint32_t slot = 0xFFFFFFFFu;
    if (anchor_within((int32_t)r.x, (int32_t)r.y, sb, akey, rowptr, qbeg, qend, schr, dist)) {
      const unsigned long long key = ((unsigned long long)r.x << 32) | r.y;
      slot = pair_hash2(r.x, r.y) & mask;
      for (;;) {
        const unsigned long long prev = atomicCAS(K + slot, 0xFFFFFFFFFFFFFFFFull, key);
        if (prev == 0xFFFFFFFFFFFFFFFFull || prev == key) break;
        slot = (slot + 1u) & mask;
      }
      atomicMin(V + slot, ((unsigned long long)r.w << 32) | t);
    }
    Cslot[t] = slot;
  }
}
// the unique hits (first occurrences), in any order
__global__ void k_hits_emit(const uint4* __restrict__ cand, const uint32_t* __restrict__ n_p, const uint32_t* __restrict__ Cslot,
                            const unsigned long long* __restrict__ V, int sb, uint64_t* hkey, float* hsc, uint32_t* hline, uint32_t* nh) {
  const uint32_t n = *n_p;
  const uint32_t nround = (n + 31u) & ~31u;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
    bool w = false;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (t < n) {
      const uint32_t slot = Cslot[t];
      if (slot != 0xFFFFFFFFu) { w = (uint32_t)V[slot] == t; r = cand[t]; }
    }
    const uint32_t pos = warp_append(w, nh);
    if (w) { hkey[pos] = ((uint64_t)r.x << sb) | r.y; hsc[pos] = __uint_as_float(r.z); hline[pos] = r.w; }
  }
}
// accepted[(query, subject)] lookup for every anchor line: a probe of the pair table
__global__ void k_lookup2(const unsigned long long* __restrict__ K, const unsigned long long* __restrict__ V, uint32_t mask,
                          const uint4* __restrict__ cand, const int32_t* aq, const int32_t* as_, uint32_t na, uint8_t* acc, float* raw) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na) return;
  acc[t] = 0; raw[t] = 0.f;
  if (aq[t] < 0 || as_[t] < 0) return;
  const uint32_t a = (uint32_t)aq[t], b = (uint32_t)as_[t];
  const unsigned long long key = ((unsigned long long)a << 32) | b;
  uint32_t slot = pair_hash2(a, b) & mask;
  for (;;) {
    const unsigned long long k = K[slot];
    if (k == key) { acc[t] = 1; raw[t] = __uint_as_float(cand[(uint32_t)V[slot]].z); return; }
    if (k == 0xFFFFFFFFFFFFFFFFull) return;
    slot = (slot + 1u) & mask;
  }
}
__global__ void k_nearest2(const uint64_t* __restrict__ hkey, const uint32_t* __restrict__ nh_p, int sb, const uint64_t* __restrict__ akey,
                           const uint32_t* __restrict__ ablock, const uint32_t* __restrict__ rowptr,
                           const int32_t* __restrict__ qbeg, const int32_t* __restrict__ qend,
                           const int32_t* __restrict__ schr, int dist, uint32_t* hit_block, uint8_t* state, uint32_t* counts /* 0 lifted, 1 ambiguous */) {
  const uint32_t nh = *nh_p;
  const uint32_t nround = (nh + 31u) & ~31u;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
    uint8_t st = 0;
    if (t < nh) {
      const uint64_t k = hkey[t];
      const uint64_t smask = (1ull << sb) - 1;
      const int32_t x = (int32_t)(k >> sb), y = (int32_t)(k & smask);
      const int32_t lo = max(x - (dist - 1), qbeg[x]), hi = min(x + dist - 1, qend[x] - 1);
      uint32_t bt = 0xFFFFFFFFu, bblk = 0; bool amb = false;
      int best = dist;
      if (hi >= lo) {
        const uint32_t t0 = rowptr[lo], t1 = rowptr[hi + 1];
        const int32_t cy = schr[y];
        for (uint32_t u = t0; u < t1; ++u) {
          const uint64_t a = akey[u];
          const int32_t ax = (int32_t)(a >> sb), ay = (int32_t)(a & smask);
          const int d = abs(ax - x) + abs(ay - y);
          if (d > best) continue;
          if (schr[ay] != cy) continue;
          if (d < best) { best = d; bt = u; bblk = ablock[u]; amb = false; }
          else if (bt != 0xFFFFFFFFu && ablock[u] != bblk) amb = true;
        }
      }
      // state: 0 none within dist / is an anchor itself, 1 lifted (unambiguous), 2 needs the tree walk
      if (bt != 0xFFFFFFFFu && best > 0) st = amb ? 2 : 1;
      state[t] = st;
      hit_block[t] = st == 1 ? bblk : 0u;
    }
    const unsigned ml = __ballot_sync(0xFFFFFFFFu, st != 0), ma = __ballot_sync(0xFFFFFFFFu, st == 2);
    if ((threadIdx.x & 31) == 0) {
      if (ml) atomicAdd(counts, (uint32_t)__popc(ml));
      if (ma) atomicAdd(counts + 1, (uint32_t)__popc(ma));
    }
  }
}
__global__ void k_gather_amb2(const uint8_t* state, const uint32_t* nh_p, uint32_t* amb_hit, uint32_t* namb) {
  const uint32_t nh = *nh_p;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nh && state[t] == 2) amb_hit[atomicAdd(namb, 1u)] = t;
}
// lifted hits (any order) with the key of the reference's append order: (block, sorted chromosome pair, file order)
__global__ void k_lift_emit(const uint8_t* state, const uint32_t* nh_p, const uint64_t* hkey, const uint32_t* hline, const uint32_t* hit_block,
                            int sb, const int32_t* qchr, const int32_t* schr, uint64_t nschr, int ob, int cb, uint64_t* keys,
                            uint32_t* vals, uint32_t* npos) {
  const uint32_t nh = *nh_p;
  const uint32_t nround = (nh + 31u) & ~31u;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
    const bool w = t < nh && state[t] != 0;
    const uint32_t pos = warp_append(w, npos);
    if (w) {
      const uint64_t k = hkey[t];
      const int32_t x = (int32_t)(k >> sb), y = (int32_t)(k & ((1ull << sb) - 1));
      const uint64_t cp = (uint64_t)qchr[x] * nschr + (uint64_t)schr[y];
      keys[pos] = ((uint64_t)hit_block[t] << (ob + cb)) | (cp << ob) | hline[t];
      vals[pos] = t;
    }
  }
}
}  // namespace

__global__ void k_pack_accepted(const uint8_t* acc, const float* raw, uint32_t n, long long* out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = ((long long)acc[t] << 32) | (long long)__float_as_uint(raw[t]);
}
// the anchor side of a liftover: anchors sorted by (qi, si) with a row pointer per query rank, the occupancy bitmap of
// the cells within `dist` of an anchor, device counters (0 anchors in the BED, 1 rank out of range, 2 duplicate,
// 3 candidates, 4 unique hits, 5 lifted, 6 ambiguous, 7 emitted)
struct LoAnchors {
  int sb = 0, qb = 0, cshift = 4, dist_pre = 1;
  uint32_t nqc = 0, nsc = 0, na = 0;
  uint64_t* aks = nullptr; uint32_t* dablock = nullptr; uint32_t* rowptr = nullptr; uint32_t* bitmap = nullptr; uint32_t* d_cnt = nullptr;
};
static int lo_anchors(jx_ctx* ctx, Arena& arena, int dist, const int32_t* daq, const int32_t* das, const uint32_t* dab, uint32_t nal,
                      LoAnchors& L) {
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  int rc;
  L.sb = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)); L.qb = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
  const int sb = L.sb, kb = L.sb + L.qb;
  JX_ALLOC(d_cnt, uint32_t, 12, true);
  JX_ALLOC(akey, uint64_t, (size_t)nal + 1, false);
  JX_ALLOC(aord, uint32_t, (size_t)nal + 1, false);
  JX_LAUNCH(k_anchor_keys, nblk(std::max(nal, 1u)), 256, 0, daq, das, nal, sb, kb, (uint32_t)Q.G, (uint32_t)S.G, akey, aord, d_cnt);
  uint64_t* aks = akey; uint32_t* aos = aord;
  if (nal && (rc = jx_sort_pairs(ctx, arena, aks, aos, nal, 0, kb + 1))) return rc;
  JX_ALLOC(dablock, uint32_t, (size_t)nal + 1, false);
  JX_LAUNCH(k_anchor_blocks2, nblk(std::max(nal, 1u)), 256, 0, aks, aos, dab, d_cnt, dablock, d_cnt + 2);
  JX_ALLOC(rowptr, uint32_t, (size_t)Q.G + 2, false);
  JX_LAUNCH(k_rowptr2, nblk((size_t)Q.G + 1), 256, 0, aks, d_cnt, sb, (uint32_t)Q.G, rowptr);
  int cshift = 4;                                          // 16 x 16 rank cells, coarser for huge genomes
  while ((((uint64_t)Q.G >> cshift) + 1) * (((uint64_t)S.G >> cshift) + 1) > (1ull << 31)) ++cshift;
  L.cshift = cshift;
  L.nqc = (uint32_t)((uint64_t)Q.G >> cshift) + 1; L.nsc = (uint32_t)((uint64_t)S.G >> cshift) + 1;
  const size_t nwords = (size_t)(((uint64_t)L.nqc * L.nsc + 31) >> 5);
  JX_ALLOC(bitmap, uint32_t, nwords, true);
  // the prefilter keeps records at distance 0 from an anchor even when dist == 0 (`scan --dist=1 --liftover`
  // gives liftover_dist 0): AnchorFile.filter_blocks needs the anchors' own raw pairs; lifting stays `< dist`
  L.dist_pre = std::max(dist, 1);
  JX_LAUNCH(k_mark_cells2, nblk(std::max(nal, 1u)), 256, 0, aks, d_cnt, sb, L.dist_pre, cshift, L.nsc, L.nqc, bitmap);
  L.aks = aks; L.dablock = dablock; L.rowptr = rowptr; L.bitmap = bitmap; L.d_cnt = d_cnt;
  return JX_OK;
}
static int lo_back(jx_ctx* ctx, Arena& arena, LoAnchors& L, const uint4* C0, uint32_t nC0, uint32_t n_lines, int dist,
                   const int32_t* daq, const int32_t* das, const uint32_t* dab, uint32_t nal,
                   const int32_t* h_aq, const int32_t* h_as, const int32_t* h_ablock, jx_liftover_result* out, uint64_t* dev_out = nullptr);

// liftover on device-resident records and anchors.  d_aq / d_as / d_ablock: per anchor LINE (file order; -1 = name not
// in the BED).  h_*: the same on the host when the caller has them (only the rare tie walk needs them; null: copied back).
int jx_liftover_core(jx_ctx* ctx, Arena& arena, const uint4* recs, uint32_t n_lines, const unsigned long long* tstats, int dist,
                     const int32_t* daq, const int32_t* das, const uint32_t* dab, uint32_t nal,
                     const int32_t* h_aq, const int32_t* h_as, const int32_t* h_ablock, jx_liftover_result* out) {
  int rc;
  out->stats[0] = (int64_t)(tstats[0] - tstats[1]);
  out->stats[5] = (int64_t)tstats[1];
  out->stats[1] = (int64_t)tstats[2];        // mapped raw records (before pair de-duplication)
  out->accepted = (uint8_t*)calloc((size_t)nal + 1, 1);
  out->raw_score = (float*)calloc((size_t)nal + 1, sizeof(float));
  auto alloc_out = [&](uint32_t n) {
    out->q_rank = (int32_t*)malloc(4 * ((size_t)n + 1)); out->s_rank = (int32_t*)malloc(4 * ((size_t)n + 1));
    out->score = (float*)malloc(4 * ((size_t)n + 1)); out->block = (int32_t*)malloc(4 * ((size_t)n + 1));
    out->n_lifted = n;
  };
  if (!nal || !tstats[2]) { alloc_out(0); JX_CK(cudaStreamSynchronize(ctx->stream)); return JX_OK; }
  LoAnchors L;
  if ((rc = lo_anchors(ctx, arena, dist, daq, das, dab, nal, L))) return rc;
  JX_TRACE("lo: anchors sorted");
  // ---- raw records -> candidates near an anchor -> unique (qi,si), first occurrence ------------------------
  const size_t capC = (size_t)tstats[2] + 1;
  JX_ALLOC(C0, uint4, capC, false);
  if (n_lines) {
    unsigned grid = std::min<unsigned>(nblk((n_lines + 3) / 4), 148u * 8u);
    JX_LAUNCH(k_select_cell2, grid, 256, 0, recs, n_lines, L.bitmap, L.cshift, L.nsc, C0, L.d_cnt + 3);
  }
  uint32_t h4[4] = {0, 0, 0, 0};
  JX_CK(cudaMemcpyAsync(h4, L.d_cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (h4[1]) { ctx->err = "anchor rank out of range"; return JX_EARG; }
  if (h4[2]) { ctx->err = "duplicate anchors"; return JX_EDUPANCHOR; }
  L.na = h4[0];
  out->stats[2] = L.na;
  JX_TRACE("lo: select_cell");
  if (!L.na || !h4[3]) { alloc_out(0); return JX_OK; }
  return lo_back(ctx, arena, L, C0, h4[3], n_lines, dist, daq, das, dab, nal, h_aq, h_as, h_ablock, out);
}

// candidates C0[0 .. nC0) = {qi, si, score bits, order} (order: file-order key, < n_lines) -> the liftover result
// dev_out != NULL (sharded run): the result stays on the device in session memory -- dev_out[0] = int32 columns
// {q | s | score bits | block} x n_lifted, dev_out[1] = int64 per anchor line: accepted << 32 | raw score bits -- and the
// host arrays of `out` are not filled (counts and stats are)
static int lo_back(jx_ctx* ctx, Arena& arena, LoAnchors& L, const uint4* C0, uint32_t nC0, uint32_t n_lines, int dist,
                   const int32_t* daq, const int32_t* das, const uint32_t* dab, uint32_t nal,
                   const int32_t* h_aq, const int32_t* h_as, const int32_t* h_ablock, jx_liftover_result* out, uint64_t* dev_out) {
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  int rc;
  const int sb = L.sb;
  uint64_t* const aks = L.aks; uint32_t* const dablock = L.dablock; uint32_t* const rowptr = L.rowptr; uint32_t* const d_cnt = L.d_cnt;
  const int dist_pre = L.dist_pre;
  auto alloc_out = [&](uint32_t n) {
    out->q_rank = (int32_t*)malloc(4 * ((size_t)n + 1)); out->s_rank = (int32_t*)malloc(4 * ((size_t)n + 1));
    out->score = (float*)malloc(4 * ((size_t)n + 1)); out->block = (int32_t*)malloc(4 * ((size_t)n + 1));
    out->n_lifted = n;
  };
  JX_CK(cudaMemcpyAsync(d_cnt + 3, &nC0, 4, cudaMemcpyHostToDevice, ctx->stream));
  uint32_t hcap = 1024;
  while (hcap < 2u * nC0 + 16u && hcap < (1u << 31)) hcap <<= 1;
  const uint32_t hmask = hcap - 1u;
  const unsigned g1 = std::max(1u, std::min<unsigned>(nblk(nC0), 148u * 8u));
  JX_ALLOC(HK, unsigned long long, hcap, false);
  JX_ALLOC(HV, unsigned long long, hcap, false);
  JX_ALLOC(Cslot, uint32_t, (size_t)nC0 + 1, false);
  JX_CK(cudaMemsetAsync(HK, 0xFF, (size_t)hcap * 8, ctx->stream));
  JX_CK(cudaMemsetAsync(HV, 0xFF, (size_t)hcap * 8, ctx->stream));
  JX_LAUNCH(k_near_insert, g1, 256, 0, C0, d_cnt + 3, sb, aks, rowptr, Q.chr_begin, Q.chr_end, S.chr_lex, dist_pre, HK, HV, hmask, Cslot);
  JX_ALLOC(hkey, uint64_t, (size_t)nC0 + 1, false);
  JX_ALLOC(hsc, float, (size_t)nC0 + 1, false);
  JX_ALLOC(hline, uint32_t, (size_t)nC0 + 1, false);
  JX_LAUNCH(k_hits_emit, g1, 256, 0, C0, d_cnt + 3, Cslot, HV, sb, hkey, hsc, hline, d_cnt + 4);
  // accepted[(query, subject)] for every anchor line
  JX_ALLOC(dacc, uint8_t, (size_t)nal + 1, false);
  JX_ALLOC(draw, float, (size_t)nal + 1, false);
  JX_LAUNCH(k_lookup2, nblk(nal), 256, 0, HK, HV, hmask, C0, daq, das, nal, dacc, draw);
  if (dev_out) {
    JX_SALLOC(packed, long long, (size_t)nal + 1, false);
    JX_LAUNCH(k_pack_accepted, nblk(nal), 256, 0, dacc, draw, nal, packed);
    dev_out[1] = (uint64_t)(uintptr_t)packed;
  } else {
    JX_CK(cudaMemcpyAsync(out->accepted, dacc, nal, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaMemcpyAsync(out->raw_score, draw, (size_t)nal * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  JX_TRACE("lo: near+dedupe+lookup");

  // ---- nearest anchor per hit -------------------------------------------------------------------------------
  JX_ALLOC(hit_block, uint32_t, (size_t)nC0 + 1, false);
  JX_ALLOC(state, uint8_t, (size_t)nC0 + 1, false);
  JX_LAUNCH(k_nearest2, g1, 128, 0, hkey, d_cnt + 4, sb, aks, dablock, rowptr, Q.chr_begin, Q.chr_end, S.chr_lex, dist, hit_block, state,
            d_cnt + 5);
  uint32_t h3[3] = {0, 0, 0};
  JX_CK(cudaMemcpyAsync(h3, d_cnt + 4, 12, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const uint32_t nh = h3[0], nlift = h3[1], namb = h3[2];
  out->stats[3] = nlift;
  out->stats[4] = namb;

  // ---- cross-block ties: exact cKDTree emulation ---------------------------------------------------------------
  if (namb) {
    JX_ALLOC(amb_hit, uint32_t, namb, false);
    JX_LAUNCH(k_gather_amb2, nblk(nh), 256, 0, state, d_cnt + 4, amb_hit, d_cnt + 8);
    std::vector<uint32_t> amb_h(namb);
    std::vector<uint64_t> hk_h(nh);
    std::vector<int32_t> caq, cas, cab;
    if (!h_aq) {                                   // anchors that never left the device (fused scan + liftover)
      caq.resize(nal); cas.resize(nal); cab.resize(nal);
      JX_CK(cudaMemcpyAsync(caq.data(), daq, (size_t)nal * 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(cas.data(), das, (size_t)nal * 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(cab.data(), dab, (size_t)nal * 4, cudaMemcpyDeviceToHost, ctx->stream));
      h_aq = caq.data(); h_as = cas.data(); h_ablock = cab.data();
    }
    JX_CK(cudaMemcpyAsync(amb_h.data(), amb_hit, (size_t)namb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaMemcpyAsync(hk_h.data(), hkey, (size_t)nh * 8, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    // anchors in the BED, in file order (all_anchors[chr_pair], synteny.py:392)
    std::vector<uint64_t> akeys_h; std::vector<uint32_t> ablock_h;
    akeys_h.reserve(nal); ablock_h.reserve(nal);
    for (uint32_t t = 0; t < nal; ++t) {
      if (h_aq[t] < 0 || h_as[t] < 0) continue;
      akeys_h.push_back(((uint64_t)(uint32_t)h_aq[t] << sb) | (uint32_t)h_as[t]);
      ablock_h.push_back((uint32_t)h_ablock[t]);
    }
    const uint64_t smask = (1ull << sb) - 1;
    auto cp_of = [&](uint64_t key) {
      return (uint64_t)Q.h_chr_lex[(size_t)(key >> sb)] * (uint64_t)S.n_chr + (uint64_t)S.h_chr_lex[(size_t)(key & smask)];
    };
    std::map<uint64_t, int> tree_of;          // chromosome pair -> tree id
    std::vector<long long> qx(namb), qy(namb); std::vector<int> qtree(namb);
    for (uint32_t i = 0; i < namb; ++i) {
      uint64_t key = hk_h[amb_h[i]];
      qx[i] = (long long)(key >> sb); qy[i] = (long long)(key & smask);
      tree_of.emplace(cp_of(key), -1);
    }
    std::map<uint64_t, std::vector<uint32_t>> members;
    for (uint32_t i = 0; i < (uint32_t)akeys_h.size(); ++i) {
      uint64_t cp = cp_of(akeys_h[i]);
      if (tree_of.count(cp)) members[cp].push_back(i);
    }
    HostTrees T;
    std::vector<const std::vector<uint32_t>*> tree_members;
    for (auto& kv : members) {
      std::vector<std::pair<long long, long long>> pts;
      pts.reserve(kv.second.size());
      for (uint32_t i : kv.second) pts.emplace_back((long long)(akeys_h[i] >> sb), (long long)(akeys_h[i] & smask));
      tree_of[kv.first] = T.add_tree(pts);
      tree_members.push_back(&kv.second);
    }
    for (uint32_t i = 0; i < namb; ++i) qtree[i] = tree_of[cp_of(hk_h[amb_h[i]])];
    std::vector<int> local; std::vector<long long> dd;
    if ((rc = run_walk(ctx, arena, T, qx, qy, qtree, (long long)dist, local, dd))) return rc;
    std::vector<uint32_t> amb_block(namb);
    for (uint32_t i = 0; i < namb; ++i) {
      if (local[i] < 0) { ctx->err = "KD-tree walk found no anchor for a flagged hit"; return JX_EINTERNAL; }
      uint32_t ai = (*tree_members[(size_t)qtree[i]])[(size_t)local[i]];
      amb_block[i] = ablock_h[ai];
    }
    JX_ALLOC(d_ab, uint32_t, namb, false);
    JX_CK(cudaMemcpyAsync(d_ab, amb_block.data(), (size_t)namb * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_LAUNCH(k_patch_amb, nblk(namb), 256, 0, amb_hit, d_ab, namb, hit_block);
    JX_CK(cudaStreamSynchronize(ctx->stream));
  }
  JX_TRACE("lo: nearest+ambiguous");

  // ---- order: (block, sorted chromosome pair, file order) in ONE key ------------------------------------------------------------
  if (dev_out) { alloc_out(0); out->n_lifted = nlift; } else alloc_out(nlift);
  if (nlift) {
    const uint64_t maxcp = (uint64_t)std::max<int64_t>(Q.n_chr, 1) * (uint64_t)std::max<int64_t>(S.n_chr, 1);
    const int ob = jx_bits((uint64_t)std::max<uint32_t>(n_lines, 1u)), cb = jx_bits(maxcp), bb = jx_bits((uint64_t)nal + 1);
    if (ob + cb + bb > 64) { ctx->err = "too many chromosome pairs / blocks / lines for one liftover sort key"; return JX_EUNSUPPORTED; }
    JX_ALLOC(k1, uint64_t, nlift, false);
    JX_ALLOC(v1, uint32_t, nlift, false);
    JX_LAUNCH(k_lift_emit, std::max(1u, std::min<unsigned>(nblk(nh), 148u * 8u)), 256, 0, state, d_cnt + 4, hkey, hline, hit_block, sb,
              Q.chr_lex, S.chr_lex, (uint64_t)S.n_chr, ob, cb, k1, v1, d_cnt + 7);
    uint64_t* k1s = k1; uint32_t* v1s = v1;
    if ((rc = jx_sort_pairs(ctx, arena, k1s, v1s, nlift, 0, ob + cb + bb))) return rc;
    // one device block, one copy into pinned staging (instead of four synchronous pageable copies)
    JX_ALLOC(oblk, uint32_t, (size_t)nlift * 4, false);
    int32_t* oq = (int32_t*)oblk; int32_t* os = (int32_t*)(oblk + nlift); float* osc = (float*)(oblk + 2 * (size_t)nlift);
    int32_t* ob_ = (int32_t*)(oblk + 3 * (size_t)nlift);
    JX_LAUNCH(k_emit_lifted, nblk(nlift), 256, 0, v1s, nlift, hkey, hsc, hit_block, sb, oq, os, osc, ob_);
    if (dev_out) {
      JX_SALLOC(keep, uint32_t, (size_t)nlift * 4, false);
      JX_CK(cudaMemcpyAsync(keep, oblk, (size_t)nlift * 16, cudaMemcpyDeviceToDevice, ctx->stream));
      JX_CK(cudaStreamSynchronize(ctx->stream));
      dev_out[0] = (uint64_t)(uintptr_t)keep;
      return JX_OK;
    }
    char* stage = nullptr;
    if ((rc = jx_pin_small(ctx, (size_t)nlift * 16, &stage))) return rc;
    JX_CK(cudaMemcpyAsync(stage, oblk, (size_t)nlift * 16, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    void* dst[4] = {out->q_rank, out->s_rank, out->score, out->block};
    if (nlift < (1u << 18)) {
      for (int k = 0; k < 4; ++k) memcpy(dst[k], stage + (size_t)nlift * 4 * k, (size_t)nlift * 4);
    } else {                                       // fresh pages: fault and fill the four arrays side by side
      std::vector<std::thread> th;
      int started = 0;
      try {
        for (; started < 3; ++started) {
          const int k = started + 1;
          th.emplace_back([=]() { memcpy(dst[k], stage + (size_t)nlift * 4 * k, (size_t)nlift * 4); });
        }
      } catch (...) {}                               // no more threads: this one copies the rest
      memcpy(dst[0], stage, (size_t)nlift * 4);
      for (int k = started + 1; k < 4; ++k) memcpy(dst[k], stage + (size_t)nlift * 4 * k, (size_t)nlift * 4);
      for (auto& x : th) x.join();
    }
  }
  JX_CK(cudaStreamSynchronize(ctx->stream));
  JX_TRACE("lo: order+emit");
  JX_CK(cudaGetLastError());
  return JX_OK;
}

extern "C" {

namespace {
__global__ void k_widen_idx(const uint32_t* in, uint32_t n, uint64_t* out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = in[t];
}
__global__ void k_narrow_idx(const uint64_t* in, uint32_t n, uint32_t* out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (uint32_t)in[t];
}
}  // namespace

static int near_lines_impl(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                           const int32_t* a_qi, const int32_t* a_si, int64_t n_lines64, char** lines, int64_t* lines_len,
                           int64_t* n_out, uint64_t* dev_out);
int jx_near_lines(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                  const int32_t* a_qi, const int32_t* a_si, int64_t n_lines64, char** lines, int64_t* lines_len,
                  int64_t* n_out) {
  if (!lines) return JX_EARG;
  return near_lines_impl(ctx, text, nbytes, text_on_device, strip_names, dist, a_qi, a_si, n_lines64, lines, lines_len, n_out, nullptr);
}
int jx_near_lines_device(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                         const int32_t* a_qi, const int32_t* a_si, int64_t n_lines64, uint64_t* lines_dev, int64_t* lines_len,
                         int64_t* n_out) {
  if (!lines_dev) return JX_EARG;
  char* unused = nullptr;
  return near_lines_impl(ctx, text, nbytes, text_on_device, strip_names, dist, a_qi, a_si, n_lines64, &unused, lines_len, n_out, lines_dev);
}
static int near_lines_impl(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist,
                           const int32_t* a_qi, const int32_t* a_si, int64_t n_lines64, char** lines, int64_t* lines_len,
                           int64_t* n_out, uint64_t* dev_out) {
  if (!ctx || !lines || !lines_len || !n_out || n_lines64 < 0 || n_lines64 >= (1ll << 31) || dist < 0) return JX_EARG;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  *lines = nullptr; *lines_len = 0; *n_out = 0;
  int rc;
  const int sb = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)), qb = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
  if (dev_out) *dev_out = 0;
  TokOut tok;
  {
    auto& c = ctx->cache;                   // records of the same held slice (jx_cscore_begin / jx_blastfilter), one-shot
    if (c.valid && c.plain && text_on_device && (const uint8_t*)text == c.text && nbytes == c.nbytes &&
        (strip_names != 0) == (c.strip != 0) && !ctx->is_self) {
      tok.recs = c.recs; tok.desc = c.desc; tok.ntiles = c.ntiles; tok.n_lines = c.n_lines; tok.d_text = c.text; tok.nbytes = c.nbytes;
      tok.li.desc = c.desc; tok.li.ntiles = c.ntiles; tok.li.cap = c.cap; tok.li.sub_bytes = c.sub_bytes;
      memcpy(tok.stats, c.stats, sizeof tok.stats);
      c.valid = false;
    } else {
      TokMode mode;
      mode.strip = strip_names;
      mode.neardiag = ctx->is_self ? 1 : 0;
      if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
    }
  }
  std::vector<uint64_t> akeys_h;
  for (int64_t t = 0; t < n_lines64; ++t)
    if (a_qi[t] >= 0 && a_si[t] >= 0 && a_qi[t] < Q.G && a_si[t] < S.G)
      akeys_h.push_back(((uint64_t)(uint32_t)a_qi[t] << sb) | (uint32_t)a_si[t]);
  const uint32_t na = (uint32_t)akeys_h.size();
  if (!na || !tok.stats[2]) { if (!dev_out) *lines = (char*)calloc(1, 1); return JX_OK; }
  JX_ALLOC(akey, uint64_t, na, false);
  JX_CK(cudaMemcpyAsync(akey, akeys_h.data(), (size_t)na * 8, cudaMemcpyHostToDevice, ctx->stream));
  uint64_t* aks = akey;
  if ((rc = jx_sort_keys(ctx, arena, aks, na, 0, sb + qb))) return rc;
  JX_ALLOC(rowptr, uint32_t, (size_t)Q.G + 2, false);
  JX_LAUNCH(k_rowptr, nblk((size_t)Q.G + 1), 256, 0, aks, na, sb, (uint32_t)Q.G, rowptr);
  int cshift = 4;
  while ((((uint64_t)Q.G >> cshift) + 1) * (((uint64_t)S.G >> cshift) + 1) > (1ull << 31)) ++cshift;
  const uint32_t nqc = (uint32_t)((uint64_t)Q.G >> cshift) + 1, nsc = (uint32_t)((uint64_t)S.G >> cshift) + 1;
  JX_ALLOC(bitmap, uint32_t, (size_t)(((uint64_t)nqc * nsc + 31) >> 5), true);
  const int dist_pre = std::max(dist, 1);        // records AT an anchor are kept even when dist == 0 (see jx_liftover_text)
  JX_LAUNCH(k_mark_cells, nblk(na), 256, 0, aks, na, sb, dist_pre, cshift, nsc, nqc, bitmap);
  const size_t capC = (size_t)tok.stats[2] + 1;
  JX_ALLOC(C0, uint4, capC, false);
  JX_ALLOC(C0idx, uint32_t, capC, false);
  JX_ALLOC(d_cnt, uint32_t, 2, true);
  unsigned grid = std::min<unsigned>(nblk((tok.n_lines + 3) / 4), 148u * 8u);
  JX_LAUNCH(k_select_cell, grid, 256, 0, tok.recs, tok.n_lines, bitmap, cshift, nsc, C0, C0idx, d_cnt);
  uint32_t nC0 = 0;
  JX_CK(cudaMemcpyAsync(&nC0, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  JX_ALLOC(C, uint4, (size_t)nC0 + 1, false);
  JX_ALLOC(Cidx, uint32_t, (size_t)nC0 + 1, false);
  if (nC0) JX_LAUNCH(k_keep_near, nblk(nC0), 256, 0, C0, C0idx, nC0, sb, aks, rowptr, Q.chr_begin, Q.chr_end, S.chr_lex, dist_pre, C, Cidx,
                     d_cnt + 1);
  uint32_t nC = 0;
  JX_CK(cudaMemcpyAsync(&nC, d_cnt + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  // back to file order
  JX_ALLOC(k64, uint64_t, nC, false);
  JX_ALLOC(idx, uint32_t, nC, false);
  if (nC) {
    JX_LAUNCH(k_widen_idx, nblk(nC), 256, 0, Cidx, nC, k64);
    uint64_t* ks = k64;
    if ((rc = jx_sort_keys(ctx, arena, ks, nC, 0, 32))) return rc;
    JX_LAUNCH(k_narrow_idx, nblk(nC), 256, 0, ks, nC, idx);
  }
  if ((rc = jx_gather_lines(ctx, arena, tok.d_text, tok.nbytes, tok.li, tok.recs, idx, nC, lines, lines_len, 0, dev_out))) return rc;
  *n_out = nC;
  return JX_OK;
}

int jx_nearest_anchor(jx_ctx* ctx, const int64_t* points_xy, int64_t n_points, const int64_t* anchors_xy,
                      int64_t n_anchors, int dist, int64_t* nearest, int64_t* dist_out) {
  if (!ctx || n_points < 0 || n_anchors < 0 || n_points >= (1ll << 31) || n_anchors >= (1ll << 31)) return JX_EARG;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  HostTrees T;
  std::vector<std::pair<long long, long long>> pts((size_t)n_anchors);
  for (int64_t i = 0; i < n_anchors; ++i) pts[(size_t)i] = {(long long)anchors_xy[2 * i], (long long)anchors_xy[2 * i + 1]};
  T.add_tree(pts);
  std::vector<long long> qx((size_t)n_points), qy((size_t)n_points);
  std::vector<int> qtree((size_t)n_points, 0);
  for (int64_t i = 0; i < n_points; ++i) { qx[(size_t)i] = points_xy[2 * i]; qy[(size_t)i] = points_xy[2 * i + 1]; }
  std::vector<int> local; std::vector<long long> dd;
  int rc = run_walk(ctx, arena, T, qx, qy, qtree, (long long)dist, local, dd);
  if (rc) return rc;
  for (int64_t i = 0; i < n_points; ++i) {
    nearest[i] = local[(size_t)i] < 0 ? n_anchors : local[(size_t)i];
    if (dist_out) dist_out[i] = dd[(size_t)i];
  }
  return JX_OK;
}

}  // extern "C"

// ---- liftover of one genome pair sharded over ranks (see jx_filter.cu: jx_shard_*) -------------------------------------
// Every rank tests ITS raw records against ALL anchors (the anchors are small and every rank has them), keeps the ones
// within `dist` of an anchor and sends them to the owner of their chromosome pair (jx_shard_partition + all-to-all); the
// owner runs the rest of liftover on what it receives.
namespace {
__global__ void k_near_emit(const uint4* __restrict__ cand, const uint32_t* __restrict__ n_p, int sb, const uint64_t* __restrict__ akey,
                            const uint32_t* __restrict__ rowptr, const int32_t* __restrict__ qbeg, const int32_t* __restrict__ qend,
                            const int32_t* __restrict__ schr, int dist, uint32_t ord_base, uint4* out, uint32_t* counter) {
  const uint32_t n = *n_p;
  const uint32_t nround = (n + 31u) & ~31u;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
    uint4 r = make_uint4(0, 0, 0, 0);
    bool keep = false;
    if (t < n) {
      r = cand[t];
      keep = anchor_within((int32_t)r.x, (int32_t)r.y, sb, akey, rowptr, qbeg, qend, schr, dist);
    }
    const uint32_t slot = warp_append(keep, counter);
    if (keep) { r.w += ord_base; out[slot] = r; }
  }
}
}  // namespace

extern "C" {

// this rank's raw records (of the slice jx_cscore_begin tokenized, or `text` when given) within dist of an anchor:
// device array of {qi, si, score bits, ord_base + record index} + count
int jx_shard_near(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int dist, const int32_t* a_qi,
                  const int32_t* a_si, int64_t nal64, int anchors_on_device, uint32_t ord_base, uint64_t* cand_dev, int64_t* n_out) {
  if (!ctx || !cand_dev || !n_out || nal64 < 0 || nal64 >= (1ll << 31) || dist < 0) return JX_EARG;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  const uint32_t nal = (uint32_t)nal64;
  int rc;
  *cand_dev = 0; *n_out = 0;
  const uint4* recs = nullptr; uint32_t n_lines = 0;
  unsigned long long tstats[12];
  {
    auto& c = ctx->cache;                   // records of the same held slice (jx_cscore_begin), one-shot
    if (c.valid && c.plain && (!text || (text_on_device && (const uint8_t*)text == c.text && nbytes == c.nbytes)) &&
        (strip_names != 0) == (c.strip != 0) && !ctx->is_self) {
      recs = c.recs; n_lines = c.n_lines;
      memcpy(tstats, c.stats, sizeof tstats);
      c.valid = false;
    } else {
      if (!text) { ctx->err = "jx_shard_near: no retained records and no text"; return JX_EARG; }
      TokMode mode;
      mode.strip = strip_names;
      mode.neardiag = ctx->is_self ? 1 : 0;
      TokOut tok;
      if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
      recs = tok.recs; n_lines = tok.n_lines;
      memcpy(tstats, tok.stats, sizeof tstats);
    }
  }
  if ((uint64_t)ord_base + n_lines >= (1ull << 31)) { ctx->err = "more than 2^31 record slots over all ranks"; return JX_EUNSUPPORTED; }
  if (!nal || !tstats[2]) return JX_OK;
  JX_ALLOC(daq0, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(das0, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(dab, uint32_t, (size_t)nal + 1, true);
  const int32_t* daq = daq0; const int32_t* das = das0;
  if (anchors_on_device) { daq = a_qi; das = a_si; }
  else {
    JX_CK(cudaMemcpyAsync(daq0, a_qi, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(das0, a_si, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  LoAnchors L;
  if ((rc = lo_anchors(ctx, arena, dist, daq, das, dab, nal, L))) return rc;
  const size_t capC = (size_t)tstats[2] + 1;
  JX_ALLOC(C0, uint4, capC, false);
  if (n_lines) {
    unsigned grid = std::min<unsigned>(nblk((n_lines + 3) / 4), 148u * 8u);
    JX_LAUNCH(k_select_cell2, grid, 256, 0, recs, n_lines, L.bitmap, L.cshift, L.nsc, C0, L.d_cnt + 3);
  }
  uint32_t h4[4] = {0, 0, 0, 0};
  JX_CK(cudaMemcpyAsync(h4, L.d_cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (h4[1]) { ctx->err = "anchor rank out of range"; return JX_EARG; }
  if (!h4[0] || !h4[3]) return JX_OK;
  JX_SALLOC(out, uint4, (size_t)h4[3] + 1, false);
  JX_LAUNCH(k_near_emit, std::max(1u, std::min<unsigned>(nblk(h4[3]), 148u * 8u)), 256, 0, C0, L.d_cnt + 3, L.sb, L.aks, L.rowptr,
            Q.chr_begin, Q.chr_end, S.chr_lex, L.dist_pre, ord_base, out, L.d_cnt + 7);
  uint32_t nn = 0;
  JX_CK(cudaMemcpyAsync(&nn, L.d_cnt + 7, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  *cand_dev = (uint64_t)(uintptr_t)out; *n_out = nn;
  return JX_OK;
}

// owner side: the near candidates received for this rank's chromosome pairs + all anchor lines (ranks and GLOBAL block
// ids) -> lifted pairs ordered by (block, chromosome pair, global order), accepted / raw score for the anchor lines whose
// pair this rank received.  n_slots_total = upper bound of the global order values.
int jx_shard_lift(jx_ctx* ctx, uint64_t recv_dev, int64_t n, int64_t n_slots_total, int dist, const int32_t* a_qi, const int32_t* a_si,
                  const int32_t* a_block, int64_t nal64, int anchors_on_device, jx_liftover_result* out, uint64_t* dev_out) {
  if (!ctx || !out || n < 0 || n >= (1ll << 31) || nal64 < 0 || nal64 >= (1ll << 31) || n_slots_total < 0 || n_slots_total >= (1ll << 31))
    return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t nal = (uint32_t)nal64;
  int rc;
  out->accepted = (uint8_t*)calloc((size_t)nal + 1, 1);
  out->raw_score = (float*)calloc((size_t)nal + 1, sizeof(float));
  auto alloc_out = [&](uint32_t k) {
    out->q_rank = (int32_t*)malloc(4 * ((size_t)k + 1)); out->s_rank = (int32_t*)malloc(4 * ((size_t)k + 1));
    out->score = (float*)malloc(4 * ((size_t)k + 1)); out->block = (int32_t*)malloc(4 * ((size_t)k + 1));
    out->n_lifted = k;
  };
  if (dev_out) { dev_out[0] = 0; dev_out[1] = 0; }
  if (!nal) { alloc_out(0); return JX_OK; }
  if (!n) {
    alloc_out(0);
    if (dev_out) {                                 // nothing received: every anchor line "not accepted" here
      JX_SALLOC(packed, long long, (size_t)nal + 1, true);
      JX_CK(cudaStreamSynchronize(ctx->stream));
      dev_out[1] = (uint64_t)(uintptr_t)packed;
    }
    return JX_OK;
  }
  JX_ALLOC(daq0, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(das0, int32_t, (size_t)nal + 1, false);
  JX_ALLOC(dab0, uint32_t, (size_t)nal + 1, false);
  const int32_t* daq = daq0; const int32_t* das = das0; const uint32_t* dab = dab0;
  if (anchors_on_device) { daq = a_qi; das = a_si; dab = (const uint32_t*)a_block; }
  else {
    JX_CK(cudaMemcpyAsync(daq0, a_qi, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(das0, a_si, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
    JX_CK(cudaMemcpyAsync(dab0, a_block, (size_t)nal * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  LoAnchors L;
  if ((rc = lo_anchors(ctx, arena, dist, daq, das, dab, nal, L))) return rc;
  uint32_t h3[3] = {0, 0, 0};
  JX_CK(cudaMemcpyAsync(h3, L.d_cnt, 12, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (h3[1]) { ctx->err = "anchor rank out of range"; return JX_EARG; }
  if (h3[2]) { ctx->err = "duplicate anchors"; return JX_EDUPANCHOR; }
  L.na = h3[0];
  out->stats[2] = L.na;
  if (!L.na) { alloc_out(0); return JX_OK; }
  rc = lo_back(ctx, arena, L, (const uint4*)(uintptr_t)recv_dev, (uint32_t)n, (uint32_t)n_slots_total, dist, daq, das, dab, nal,
               anchors_on_device ? nullptr : a_qi, anchors_on_device ? nullptr : a_si, anchors_on_device ? nullptr : a_block, out, dev_out);
  if (!rc && dev_out && !dev_out[1]) {             // (early exits of the core: no candidate survived)
    JX_SALLOC(packed, long long, (size_t)nal + 1, true);
    dev_out[1] = (uint64_t)(uintptr_t)packed;
  }
  if (!rc) JX_CK(cudaStreamSynchronize(ctx->stream));
  return rc;
}

}  // extern "C"

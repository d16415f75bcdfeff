// jcvi.compara.synfind (src/jcvi/compara/synfind.py:45-240) on the device.
//
// For every gene r of the query BED the reference takes the hits of the rows [rmin, rmax) around it (same seqid,
// synfind.py:155-160), sorts them by subject rank, single-links consecutive hits that are < window // 2 ranks apart on
// one subject seqid (78-86), and for every linked group decides the closest flanker (45-65), the longer of the
// LIS / LDS over the subject ranks with the patience-sort back-links of src/jcvi/algorithms/lis.py:81-116, the score
// min(#distinct x, #distinct y) of that track (- 1 without a syntelog) and the track's first / last subject rank
// (93-124); regions with score >= int(cutoff * window) are kept, ordered by decreasing score (127).
//
// Here: the distinct (qi, si) pairs of read_blast (synteny.py:255-297) come from the path's tokenizer, one radix sort
// puts them in (x, y) order, so the window of a gene is ONE contiguous slice of that list.  One CTA per query gene:
//   1. the slice is sorted by (y, position in the slice) in shared memory (bitonic network) -- the position in the
//      slice is the (x, y) order, so ties in y are in x order like the reference's stable sort;
//   2. link flags of neighbours, group ids by a block scan; groups are the runs of linked neighbours;
//   3. a second in-memory sort by (group, position in the slice) leaves every group contiguous and in (x, y) order,
//      which is the order get_flanker's group.sort() produces and the order the LIS runs over;
//   4. one thread per group: flanker by binary search, patience sort forwards (LIS: ties in y extend, because the
//      reference compares (y, i) tuples with growing i) and backwards (LDS: strictly decreasing), back-links to the
//      element that was on top of the previous pile, walk from the FIRST element of the last pile (lis.py:110-114).
// Genes are dealt to size classes (slice <= 1024 / 4096 / 16384 points in shared memory; larger slices use a per-CTA
// scratch in global memory with 64-bit keys) so that small windows run many CTAs per SM.  Regions are appended to one
// list and ordered at the end by (gene, -score, first element of the group) with one radix sort: `sorted(g)` orders
// the groups by their first element in (y, x) order compared as (x, y) tuples, and the final sort by score is stable.
#include "jx_internal.cuh"

namespace {

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

constexpr int SF_CLASSES = 4;                                  // 3 shared-memory classes + the global-scratch class
constexpr uint32_t SF_CAP[SF_CLASSES] = {1024u, 4096u, 16384u, 0xFFFFFFFFu};
constexpr int SF_THREADS[SF_CLASSES] = {128, 256, 512, 512};

struct SfSide {
  const int32_t* xbegin;      // per query rank: first rank of its seqid run
  const int32_t* xend;        // one past the last
  const int32_t* ychr;        // seqid id per subject rank
  uint32_t Gx;
};

struct SfArgs {
  const uint32_t* py;         // subject rank of every point; points sorted by (x, y)
  const uint32_t* rowptr;     // [Gx + 1]
  SfSide side;
  int w, cutoff;
  int sb, xb, scb;            // bits of a subject rank / of a row offset inside a window / of a score
  uint32_t smax;
  const uint32_t* genes;      // the class's gene list
  const uint32_t* ngenes;
  uint32_t* ticket;
  uint32_t* counter;          // regions appended (keeps counting past cap)
  uint32_t cap;
  uint4* out;                 // {gene, syntelog, left, right}
  uint64_t* okey;
  uint32_t* ometa;            // score << 3 | gray << 1 | minus
  unsigned char* scratch;     // global-scratch class: per CTA region
  size_t scratch_per_cta;
  uint32_t p2cap;
};

__global__ void k_sf_keys(const uint4* __restrict__ recs, uint32_t n, int sb, uint64_t* keys, uint32_t* counter) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t nround = (n + 31u) & ~31u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    uint4 r = make_uint4(JX_NOREC, 0, 0, 0);
    if (i < n) r = __ldcs(recs + i);
    const bool keep = r.x != JX_NOREC;
    const uint32_t slot = warp_append(keep, counter);
    if (keep) keys[slot] = ((uint64_t)r.x << sb) | r.y;
  }
}
__global__ void k_sf_heads(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* flag) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) flag[t] = (t == 0 || keys[t - 1] != keys[t]) ? 1u : 0u;
}
__global__ void k_sf_compact(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ pos,
                             uint32_t n, uint64_t* out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && flag[t]) out[pos[t]] = keys[t];
}
// (x << sb | y) -> (y << qb | x)
__global__ void k_sf_transpose(const uint64_t* __restrict__ keys, uint32_t n, int sb, int qb, uint64_t* out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint64_t k = keys[t];
  out[t] = ((k & ((1ull << sb) - 1)) << qb) | (k >> sb);
}
__global__ void k_sf_split(const uint64_t* __restrict__ keys, uint32_t n, int sb, uint32_t* py) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) py[t] = (uint32_t)(keys[t] & ((1ull << sb) - 1));
}
__global__ void k_sf_rowptr(const uint64_t* __restrict__ keys, uint32_t n, int sb, uint32_t Gx, uint32_t* rowptr) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x > Gx) return;
  const uint64_t want = (uint64_t)x << sb;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < want) lo = mid + 1; else hi = mid;
  }
  rowptr[x] = lo;
}

__device__ __forceinline__ void sf_window(const SfArgs& a, uint32_t r, int& rmin, int& rmax) {
  const int first = a.side.xbegin[r], last = a.side.xend[r] - 1;
  rmin = max((int)r - a.w, first);
  rmax = min((int)r + a.w + 1, last);                      // synfind.py:156-157 (the run's last rank is clipped away)
}

// genes -> size classes; cls[c * Gx ...] lists, cnt[c] counts, cnt[SF_CLASSES] = largest slice
__global__ void k_sf_classify(SfArgs a, uint32_t* cls, uint32_t* cnt) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.side.Gx) return;
  int rmin, rmax;
  sf_window(a, r, rmin, rmax);
  if (rmax <= rmin) return;
  const uint32_t n = a.rowptr[rmax] - a.rowptr[rmin];
  if (n < 2) return;
  const int c = n <= 1024u ? 0 : n <= 4096u ? 1 : n <= 16384u ? 2 : 3;        // SF_CAP
  cls[(size_t)c * a.side.Gx + atomicAdd(cnt + c, 1u)] = r;
  atomicMax(cnt + SF_CLASSES, n);
}

template <typename K>
__device__ void sf_bitonic(K* a, uint32_t P2) {
  for (uint32_t k = 2; k <= P2; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < (P2 >> 1); i += blockDim.x) {
        const uint32_t lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const uint32_t hi = lo | j;
        const bool up = (lo & k) == 0;
        const K x = a[lo], y = a[hi];
        if ((x > y) == up) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

// exclusive scan of one 64-bit value per thread over the CTA; total in *tot
__device__ unsigned long long sf_block_scan(unsigned long long v, unsigned long long* s_w, unsigned long long* tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long s = 0;
    for (int i = 0; i < nwarp; ++i) { const unsigned long long c = s_w[i]; s_w[i] = s; s += c; }
    s_w[nwarp] = s;
  }
  __syncthreads();
  const unsigned long long r = s_w[warp] + inc - v;
  *tot = s_w[nwarp];
  __syncthreads();
  return r;
}

template <typename K, typename I>
__global__ void k_synfind(SfArgs a) {
  extern __shared__ __align__(16) unsigned char sf_smem[];
  __shared__ unsigned long long s_w[20];
  __shared__ uint32_t s_gene;
  int* rows = (int*)sf_smem;                                            // [2w + 2] row starts relative to the slice
  const size_t rows_bytes = (((size_t)(2 * a.w + 3) * 4) + 15) & ~(size_t)15;
  unsigned char* buf = a.scratch ? a.scratch + (size_t)blockIdx.x * a.scratch_per_cta : sf_smem + rows_bytes;
  K* A = (K*)buf;
  K* B = A + a.p2cap;
  I* goff = (I*)(B + a.p2cap);                                          // [p2cap / 2 + 2]
  const K PAD = ~(K)0;
  const uint32_t T = blockDim.x;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_gene = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t gi = s_gene;
    if (gi >= *a.ngenes) return;
    const uint32_t r = a.genes[gi];
    int rmin, rmax;
    sf_window(a, r, rmin, rmax);
    const uint32_t base = a.rowptr[rmin];
    const uint32_t n = a.rowptr[rmax] - base;
    const int nrows = rmax - rmin;
    for (int i = threadIdx.x; i <= nrows; i += T) rows[i] = (int)(a.rowptr[rmin + i] - base);
    uint32_t P2 = 2;
    int jb = 1;
    while (P2 < n) { P2 <<= 1; ++jb; }
    const K jmask = ((K)1 << jb) - 1;
    // 1. (y, position) keys, sorted
    for (uint32_t j = threadIdx.x; j < P2; j += T) A[j] = j < n ? (((K)a.py[base + j] << jb) | (K)j) : PAD;
    __syncthreads();
    sf_bitonic(A, P2);
    // 2. link flags, group ids
    const uint32_t chunk = (n + T - 1) / T;
    const uint32_t k0 = min(n, threadIdx.x * chunk), k1 = min(n, k0 + chunk);
    auto linked = [&](uint32_t k) -> bool {                             // k and k + 1 are joined (synfind.py:85)
      if (k + 1 >= n) return false;
      const uint32_t ya = (uint32_t)(A[k] >> jb), yb = (uint32_t)(A[k + 1] >> jb);
      return (int)(yb - ya) < a.w && a.side.ychr[ya] == a.side.ychr[yb];
    };
    uint32_t nstart = 0, ning = 0;
    {
      bool prev = k0 > 0 && k0 < n ? linked(k0 - 1) : false;
      for (uint32_t k = k0; k < k1; ++k) {
        const bool cur = linked(k);
        nstart += cur && !prev;
        ning += cur || prev;
        prev = cur;
      }
    }
    unsigned long long tot;
    const unsigned long long ex = sf_block_scan(((unsigned long long)nstart << 32) | ning, s_w, &tot);
    const uint32_t ng = (uint32_t)(tot >> 32), M = (uint32_t)tot;
    {
      uint32_t gs = (uint32_t)(ex >> 32), gr = (uint32_t)ex;           // starts / group members before this chunk
      bool prev = k0 > 0 && k0 < n ? linked(k0 - 1) : false;
      for (uint32_t k = k0; k < k1; ++k) {
        const bool cur = linked(k);
        const bool st = cur && !prev, in = cur || prev;
        if (st) { goff[gs] = (I)gr; ++gs; }
        B[k] = in ? (((K)(gs - 1) << jb) | (A[k] & jmask)) : PAD;
        gr += in;
        prev = cur;
      }
    }
    for (uint32_t k = n + threadIdx.x; k < P2; k += T) B[k] = PAD;
    if (threadIdx.x == 0) goff[ng] = (I)M;
    __syncthreads();
    if (ng == 0) continue;
    // 3. groups contiguous, each in (x, y) order; then (row offset, y) per member
    sf_bitonic(B, P2);
    for (uint32_t p = threadIdx.x; p < M; p += T) {
      const uint32_t j = (uint32_t)(B[p] & jmask);
      int lo = 0, hi = nrows;                                           // last row whose start is <= j
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint32_t)rows[mid] <= j) lo = mid; else hi = mid; }
      B[p] = ((K)(uint32_t)lo << a.sb) | (K)a.py[base + j];
    }
    __syncthreads();
    // 4. one thread per group
    const K ymask = ((K)1 << a.sb) - 1;
    const uint32_t qx = r - (uint32_t)rmin;
    for (uint32_t g = threadIdx.x; g < ng; g += T) {
      const uint32_t o = (uint32_t)goff[g], m = (uint32_t)goff[g + 1] - o;
      if ((int)m < a.cutoff) continue;                                  // score <= m
      const K* E = B + o;
      I* tops = (I*)A + 2 * (size_t)o;
      I* back = tops + m;
      auto Y = [&](uint32_t t) -> uint32_t { return (uint32_t)(E[t] & ymask); };
      auto X = [&](uint32_t t) -> uint32_t { return (uint32_t)(E[t] >> a.sb); };
      // first element of the group in (y, x) order; flanker (synfind.py:45-65)
      uint64_t fmin = ~0ull;
      for (uint32_t t = 0; t < m; ++t) { const uint64_t f = ((uint64_t)Y(t) << a.xb) | X(t); fmin = f < fmin ? f : fmin; }
      uint32_t pos;
      { uint32_t lo = 0, hi = m; while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (X(mid) < qx) lo = mid + 1; else hi = mid; } pos = lo; }
      const uint32_t lf = pos == 0 ? 0 : pos - 1, rf = pos == m ? m - 1 : pos;
      const int dl = abs((int)qx - (int)X(lf)), dr = abs((int)qx - (int)X(rf));
      const uint32_t fl = dl < dr ? lf : rf;
      const bool flanked = !(pos == 0 || pos == m);
      // LIS (ties in y extend)
      uint32_t np = 0, first_last = 0;
      for (uint32_t t = 0; t < m; ++t) {
        const uint32_t y = Y(t);
        uint32_t lo = 0, hi = np;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (Y(tops[mid]) > y) hi = mid; else lo = mid + 1; }
        back[t] = lo ? tops[lo - 1] : (I)0;
        tops[lo] = (I)t;
        if (lo == np) { ++np; first_last = t; }
      }
      const uint32_t len_i = np;
      uint32_t dx_i = 1, dy_i = 1, front_i, back_i;
      {
        uint32_t e = first_last, px = X(e), pyv = Y(e);
        back_i = pyv;
        for (uint32_t i = 1; i < np; ++i) {
          e = back[e];
          const uint32_t x = X(e), y = Y(e);
          dx_i += x != px; dy_i += y != pyv; px = x; pyv = y;
        }
        front_i = pyv;
      }
      // LDS = reversed LIS of the reversed sequence (strictly decreasing)
      np = 0; first_last = 0;
      for (uint32_t t = m; t-- > 0;) {
        const uint32_t y = Y(t);
        uint32_t lo = 0, hi = np;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (Y(tops[mid]) >= y) hi = mid; else lo = mid + 1; }
        back[t] = lo ? tops[lo - 1] : (I)0;
        tops[lo] = (I)t;
        if (lo == np) { ++np; first_last = t; }
      }
      uint32_t len = len_i, dx = dx_i, dy = dy_i, front = front_i, backy = back_i, minus = 0;
      if (len_i < np) {
        minus = 1; len = np; dy = np;
        uint32_t e = first_last, px = X(e);
        front = Y(e); dx = 1;
        for (uint32_t i = 1; i < np; ++i) { e = back[e]; const uint32_t x = X(e); dx += x != px; px = x; }
        backy = Y(e);
      }
      int score = (int)min(dx, dy);
      uint32_t gray = 0;                                                // 0 S, 1 G, 2 F
      if (X(fl) != qx) { gray = flanked ? 2u : 1u; --score; }
      if (score < a.cutoff) continue;
      const uint32_t slot = atomicAdd(a.counter, 1u);
      if (slot < a.cap) {
        a.out[slot] = make_uint4(r, Y(fl), front, backy);
        a.ometa[slot] = ((uint32_t)score << 3) | (gray << 1) | minus;
        const uint64_t fx = fmin & ((1ull << a.xb) - 1), fy = fmin >> a.xb;
        a.okey[slot] = ((((uint64_t)r << a.scb) | (uint64_t)(a.smax - (uint32_t)score)) << (a.xb + a.sb)) | (fx << a.sb) | fy;
      }
    }
  }
}

__global__ void k_sf_gather(const uint32_t* __restrict__ perm, uint32_t n, const uint4* __restrict__ in, const uint32_t* __restrict__ meta,
                            int32_t* out) {           // columns {gene | syntelog | left | right | meta} x n
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t s = perm[t];
  const uint4 v = in[s];
  out[t] = (int32_t)v.x; out[(size_t)n + t] = (int32_t)v.y; out[2 * (size_t)n + t] = (int32_t)v.z; out[3 * (size_t)n + t] = (int32_t)v.w;
  out[4 * (size_t)n + t] = (int32_t)meta[s];
}
__global__ void k_sf_iota(uint32_t* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

template <typename K, typename I>
int sf_launch(jx_ctx* ctx, const SfArgs& a, int threads, unsigned grid, size_t smem) {
  if (smem > 48 * 1024) {
    if (cudaFuncSetAttribute(k_synfind<K, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      ctx->err = "synfind: shared memory request refused"; return JX_ECUDA;
    }
  }
  JX_LAUNCH((k_synfind<K, I>), grid, threads, smem, a);
  return JX_OK;
}

struct SfOut { std::vector<int32_t> cols; uint32_t n = 0; };

// one pass (forward, or the mirror with the sides swapped): keys = distinct (x << sb | y), sorted
int sf_pass(jx_ctx* ctx, Arena& arena, const uint64_t* keys, uint32_t n, int sb, const SfSide& side, int w, int cutoff, SfOut& res) {
  res.n = 0; res.cols.clear();
  const uint32_t Gx = side.Gx;
  if (!n || !Gx) return JX_OK;
  int rc;
  SfArgs a;
  memset(&a, 0, sizeof a);
  a.side = side; a.w = w; a.cutoff = cutoff; a.sb = sb;
  a.xb = jx_bits((uint64_t)(2 * w + 1));
  a.smax = (uint32_t)(2 * w + 2);
  a.scb = jx_bits((uint64_t)a.smax);
  const int qb = jx_bits((uint64_t)Gx);
  if (qb + a.scb + a.xb + sb > 64) { ctx->err = "synfind: window too large for the 64-bit region key"; return JX_EUNSUPPORTED; }
  const bool small_keys = a.xb + sb <= 32;
  JX_ALLOC(py, uint32_t, n, false);
  JX_ALLOC(rowptr, uint32_t, (size_t)Gx + 1, false);
  JX_LAUNCH(k_sf_split, nblk(n), 256, 0, keys, n, sb, py);
  JX_LAUNCH(k_sf_rowptr, nblk((size_t)Gx + 1), 256, 0, keys, n, sb, Gx, rowptr);
  a.py = py; a.rowptr = rowptr;
  JX_ALLOC(cls, uint32_t, (size_t)SF_CLASSES * Gx, false);
  JX_ALLOC(cnt, uint32_t, SF_CLASSES + 1 + SF_CLASSES + 1, true);       // class counts, largest slice, tickets, region counter
  JX_LAUNCH(k_sf_classify, nblk(Gx), 256, 0, a, cls, cnt);
  uint32_t h[SF_CLASSES + 1];
  JX_CK(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  uint32_t cap = std::max<uint32_t>(1u << 16, 4u * Gx);
  for (int attempt = 0; attempt < 2; ++attempt) {
    JX_ALLOC(out, uint4, cap, false);
    JX_ALLOC(okey, uint64_t, cap, false);
    JX_ALLOC(ometa, uint32_t, cap, false);
    JX_CK(cudaMemsetAsync(cnt + SF_CLASSES + 1, 0, (SF_CLASSES + 1) * 4, ctx->stream));
    a.out = out; a.okey = okey; a.ometa = ometa; a.cap = cap; a.counter = cnt + 2 * SF_CLASSES + 1;
    for (int c = 0; c < SF_CLASSES; ++c) {
      uint32_t ngenes = h[c];
      if (!ngenes) continue;
      a.genes = cls + (size_t)c * Gx; a.ngenes = cnt + c; a.ticket = cnt + SF_CLASSES + 1 + c;
      const size_t rows_bytes = (((size_t)(2 * w + 3) * 4) + 15) & ~(size_t)15;
      // keys (y << jb | j) and (group << jb | j) need sb + jb bits; the shared-memory classes use 32-bit keys when they fit
      int jbmax = 1; while ((1u << jbmax) < SF_CAP[c] && jbmax < 31) ++jbmax;
      const bool smem_class = c < SF_CLASSES - 1 && small_keys && sb + jbmax <= 32 && rows_bytes < 8192;
      if (smem_class) {
        a.scratch = nullptr; a.scratch_per_cta = 0; a.p2cap = SF_CAP[c];
        const size_t smem = rows_bytes + (size_t)a.p2cap * 8 + ((size_t)a.p2cap / 2 + 2) * 2;
        const unsigned per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(16, (200u << 10) / smem));
        const unsigned grid = std::min<unsigned>(ngenes, 148u * per_sm);
        if ((rc = sf_launch<uint32_t, uint16_t>(ctx, a, SF_THREADS[c], grid, smem))) return rc;
      } else {
        uint32_t p2 = 2; while (p2 < std::min<uint32_t>(h[SF_CLASSES], c < SF_CLASSES - 1 ? SF_CAP[c] : 0xFFFFFFFFu)) p2 <<= 1;
        a.p2cap = p2;
        const size_t per = (((size_t)p2 * 16 + ((size_t)p2 / 2 + 2) * 4) + 255) & ~(size_t)255;
        const unsigned grid = std::min<unsigned>(ngenes, 148u * 2u);
        JX_ALLOC(scratch, unsigned char, per * grid, false);
        a.scratch = scratch; a.scratch_per_cta = per;
        if ((rc = sf_launch<uint64_t, uint32_t>(ctx, a, SF_THREADS[c], grid, rows_bytes))) return rc;
      }
    }
    uint32_t produced = 0;
    JX_CK(cudaMemcpyAsync(&produced, a.counter, 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    JX_CK(cudaGetLastError());
    if (produced > cap) { cap = produced; continue; }                   // the list was too short: once more with the exact size
    if (!produced) return JX_OK;
    JX_ALLOC(perm, uint32_t, produced, false);
    JX_LAUNCH(k_sf_iota, nblk(produced), 256, 0, perm, produced);
    uint64_t* ks = okey; uint32_t* vs = perm;
    if ((rc = jx_sort_pairs(ctx, arena, ks, vs, produced, 0, qb + a.scb + a.xb + sb))) return rc;
    JX_ALLOC(cols, int32_t, (size_t)produced * 5, false);
    JX_LAUNCH(k_sf_gather, nblk(produced), 256, 0, vs, produced, out, ometa, cols);
    res.cols.resize((size_t)produced * 5);
    JX_CK(cudaMemcpyAsync(res.cols.data(), cols, (size_t)produced * 20, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    res.n = produced;
    return JX_OK;
  }
  ctx->err = "synfind: region list overflow"; return JX_EINTERNAL;
}

// distinct sorted keys (x << sb | y) -> both passes -> result
int sf_run(jx_ctx* ctx, Arena& arena, uint64_t* keys, uint32_t n, int sb, int qb, int window_half, int cutoff, int passes,
           jx_synfind_result* out) {
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  int rc;
  SfOut fwd, mir;
  SfSide sf{Q.chr_begin, Q.chr_end, S.chr_lex, (uint32_t)Q.G};
  if ((passes & 1) && (rc = sf_pass(ctx, arena, keys, n, sb, sf, window_half, cutoff, fwd))) return rc;
  if ((passes & 2) && n) {
    JX_ALLOC(tk, uint64_t, n, false);
    JX_LAUNCH(k_sf_transpose, nblk(n), 256, 0, keys, n, sb, qb, tk);
    uint64_t* ks = tk;
    if ((rc = jx_sort_keys(ctx, arena, ks, n, 0, sb + qb))) return rc;
    SfSide sm{S.chr_begin, S.chr_end, Q.chr_lex, (uint32_t)S.G};
    if ((rc = sf_pass(ctx, arena, ks, n, qb, sm, window_half, cutoff, mir))) return rc;
  }
  const size_t tot = (size_t)fwd.n + mir.n;
  out->n_forward = fwd.n; out->n_mirror = mir.n;
  out->n_points = n;
  if (!tot) return JX_OK;
  int32_t** dst[5] = {&out->query, &out->syntelog, &out->left, &out->right, nullptr};
  for (int c = 0; c < 4; ++c) {
    int32_t* p = (int32_t*)malloc(tot * 4);
    if (!p) { ctx->err = "out of host memory"; return JX_EINTERNAL; }
    if (fwd.n) memcpy(p, fwd.cols.data() + (size_t)c * fwd.n, (size_t)fwd.n * 4);
    if (mir.n) memcpy(p + fwd.n, mir.cols.data() + (size_t)c * mir.n, (size_t)mir.n * 4);
    *dst[c] = p;
  }
  out->score = (int32_t*)malloc(tot * 4);
  out->gray = (uint8_t*)malloc(tot);
  out->minus = (uint8_t*)malloc(tot);
  if (!out->score || !out->gray || !out->minus) { ctx->err = "out of host memory"; return JX_EINTERNAL; }
  for (size_t i = 0; i < tot; ++i) {
    const uint32_t m = (uint32_t)(i < fwd.n ? fwd.cols[4 * (size_t)fwd.n + i] : mir.cols[4 * (size_t)mir.n + (i - fwd.n)]);
    out->score[i] = (int32_t)(m >> 3); out->gray[i] = (uint8_t)((m >> 1) & 3u); out->minus[i] = (uint8_t)(m & 1u);
  }
  return JX_OK;
}

int sf_unique(jx_ctx* ctx, Arena& arena, uint64_t*& keys, uint32_t& n, int bits) {
  if (!n) return JX_OK;
  int rc;
  if ((rc = jx_sort_keys(ctx, arena, keys, n, 0, bits))) return rc;
  JX_ALLOC(flag, uint32_t, n, false);
  JX_ALLOC(pos, uint32_t, n, false);
  JX_LAUNCH(k_sf_heads, nblk(n), 256, 0, keys, n, flag);
  if ((rc = jx_exclusive_sum(ctx, arena, flag, pos, n))) return rc;
  uint32_t last[2];
  JX_CK(cudaMemcpyAsync(&last[0], pos + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&last[1], flag + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const uint32_t nu = last[0] + last[1];
  JX_ALLOC(uk, uint64_t, nu, false);
  JX_LAUNCH(k_sf_compact, nblk(n), 256, 0, keys, flag, pos, n, uk);
  keys = uk; n = nu;
  return JX_OK;
}

}  // namespace

extern "C" {

void jx_synfind_result_free(jx_synfind_result* r) {
  if (!r) return;
  free(r->query); free(r->syntelog); free(r->left); free(r->right); free(r->score); free(r->gray); free(r->minus);
  memset(r, 0, sizeof *r);
}

int jx_synfind_text(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names, int window_half,
                    int cutoff, jx_synfind_result* out) {
  if (!ctx || !out || window_half < 0) return JX_EARG;
  memset(out, 0, sizeof *out);
  if (!ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  TokMode mode;
  mode.strip = strip_names;
  mode.neardiag = ctx->is_self ? 1 : 0;
  TokOut tok;
  int rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok);
  if (rc) return rc;
  out->stats[0] = (int64_t)(tok.stats[0] - tok.stats[1]);
  out->stats[2] = (int64_t)tok.stats[1];
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  const int sb = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)), qb = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
  const size_t nvalid = (size_t)tok.stats[2];
  JX_ALLOC(keys, uint64_t, nvalid + 1, false);
  JX_ALLOC(d_cnt, uint32_t, 1, true);
  if (tok.n_lines) {
    const unsigned grid = std::min<unsigned>(nblk(tok.n_lines), 148u * 16u);
    JX_LAUNCH(k_sf_keys, grid, 256, 0, tok.recs, tok.n_lines, sb, keys, d_cnt);
  }
  uint32_t n = 0;
  JX_CK(cudaMemcpyAsync(&n, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  uint64_t* ks = keys;
  if ((rc = sf_unique(ctx, arena, ks, n, sb + qb))) return rc;
  out->stats[1] = n;
  rc = sf_run(ctx, arena, ks, n, sb, qb, window_half, cutoff, ctx->is_self ? 1 : 3, out);
  if (rc) jx_synfind_result_free(out);
  return rc;
}

int jx_synfind_points(jx_ctx* ctx, const int32_t* qi, const int32_t* si, int64_t n64, int window_half, int cutoff, int passes,
                      jx_synfind_result* out) {
  if (!ctx || !out || n64 < 0 || n64 >= (1ll << 31) || window_half < 0) return JX_EARG;
  memset(out, 0, sizeof *out);
  if (!ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  const int sb = jx_bits((uint64_t)std::max<int64_t>(S.G, 1)), qb = jx_bits((uint64_t)std::max<int64_t>(Q.G, 1));
  uint32_t n = (uint32_t)n64;
  std::vector<uint64_t> hk(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (qi[i] < 0 || si[i] < 0 || qi[i] >= Q.G || si[i] >= S.G) { ctx->err = "synfind: rank outside the BED"; return JX_EARG; }
    hk[i] = ((uint64_t)qi[i] << sb) | (uint32_t)si[i];
  }
  JX_ALLOC(keys, uint64_t, (size_t)n + 1, false);
  if (n) JX_CK(cudaMemcpyAsync(keys, hk.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  uint64_t* ks = keys;
  int rc;
  if ((rc = sf_unique(ctx, arena, ks, n, sb + qb))) return rc;
  out->stats[1] = n;
  rc = sf_run(ctx, arena, ks, n, sb, qb, window_half, cutoff, passes, out);
  if (rc) jx_synfind_result_free(out);
  return rc;
}

}  // extern "C"

// Stage 1: blastfilter on the GPU (K3-K8 of SURVEY.md section 2).
//
// Replaces src/jcvi/compara/blastfilter.py:
//   49,86-89   stable score-descending sort + first-seen set  -> per (qi,si) run: max score, min line
//   225-237    filter_cscore (float64 division, strict >)      -> k_select (fused with compaction)
//   262-284    tandem_grouper                                  -> sort + adjacent-rank unions (lock-free union-find)
//   164-185    write_localdups (mother = longest, then accn)   -> atomicMax of (length, ~lexrank) per root
//   240-259    filter_tandem                                   -> mother substitution + second run-best
//   158-161    write_new_blast                                 -> order by (score desc, line asc), host "%.3g" writer
// The C-score predicate is applied BEFORE the pair de-duplication: the reference keeps the
// best-scoring record of each pair and the predicate is monotone in the score for fixed genes,
// so the surviving record of every pair is the same, while the sort that de-duplicates only
// sees the ~5-10 % of records that pass.
#include <thread>

#include "jx_format.cuh"
#include "jx_internal.cuh"

namespace {

// 4 records per thread and iteration: the four 16-byte loads and the eight table gathers are all
// in flight before anything is consumed (the first version was latency bound: long-scoreboard
// stalls 84 warps per issue, 11 % issue utilisation)
// best score per gene rank (the table is keyed by name id: one gather level less in k_select)
__global__ void k_gene_best(const uint32_t* __restrict__ best, const uint32_t* __restrict__ nid, uint32_t G, uint32_t* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) out[i] = best[nid[i]];
}

#ifndef JX_SEL_MINB
#define JX_SEL_MINB 4
#endif
#ifndef JX_SEL_GRID
#define JX_SEL_GRID 8
#endif
__global__ void __launch_bounds__(256, JX_SEL_MINB) k_select(const uint4* __restrict__ recs, uint32_t n, const uint32_t* __restrict__ bestq,
                         const uint32_t* __restrict__ bests, double cscore,
                         int use_cscore, uint32_t ord_base, uint4* __restrict__ out, uint32_t* counter,
                         uint32_t* errflag) {
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
    uint4 r[U]; bool keep[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * 32u + lane;
      r[u] = i < n ? __ldcs(recs + i) : make_uint4(JX_NOREC, 0, 0, 0);
    }
    uint32_t bq[U], bs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      keep[u] = r[u].x != JX_NOREC;
      bq[u] = keep[u] && use_cscore ? __ldg(bestq + r[u].x) : 0u;
      bs[u] = keep[u] && use_cscore ? __ldg(bests + r[u].y) : 0u;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (keep[u] && use_cscore) {
        const double a = (double)__uint_as_float(bq[u]), b = (double)__uint_as_float(bs[u]);
        const double den = a > b ? a : b;                     // max(best_score[q], best_score[s])
        if (den == 0.0) { atomicExch(errflag, 1u); keep[u] = false; }
        else {
          // the reference's test is the IEEE quotient score / den > cscore (strict).  One multiply
          // settles all but the hair-thin band around the threshold; only there is the division done.
          const double sc = (double)__uint_as_float(r[u].z), th = den * cscore;
          if (sc < th * (1.0 - 1e-12)) keep[u] = false;
          else if (sc > th * (1.0 + 1e-12)) keep[u] = true;
          else keep[u] = (sc / den) > cscore;
        }
      }
    }
    uint32_t slot[U];
    block_append_multi<U>(keep, counter, slot, s_cnt);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (keep[u]) {     // candidate: ranks, score, (order << 1) | (evalue < 1e-10); order = record index = file order
        uint4 c = r[u];
        c.w = ((ord_base + i0 + u * 32u + lane) << 1) | (c.w & 1u);
        out[slot[u]] = c;
      }
  }
}

// ---- K3 / K7: per-pair best record through a hash table --------------------------------------------------------
// The reference keeps, per (query, subject), the first record of a stable score-descending sort (blastfilter.py:49,
// 86-89; again at 250-258 after the tandem substitution) = the record with the highest score and, among equals, the
// lowest line.  That is one 64-bit atomicMax per candidate -- value = (ordered score bits << 32) | ~order -- into an
// open-addressing table keyed by (qi << 32 | si): no sort, no run walk, and the table doubles as the pair SET the
// tandem pass probes.  `order` = the candidate's record index (file order); candidates come in any order.
#define JX_HEMPTY 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ uint32_t pair_hash(uint32_t a, uint32_t b) {
  uint32_t h = a * 0x9E3779B1u ^ b * 0x85EBCA77u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  return h;
}
__device__ __forceinline__ unsigned long long cand_value(const uint4& c) {   // higher = better: score, then earlier line
  return ((unsigned long long)score_ord(c.z) << 32) | (0xFFFFFFFFu - (c.w >> 1));
}
// find-or-insert; returns the slot
__device__ __forceinline__ uint32_t hash_slot_insert(unsigned long long* K, uint32_t mask, uint32_t a, uint32_t b) {
  const unsigned long long key = ((unsigned long long)a << 32) | b;
  uint32_t slot = pair_hash(a, b) & mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(K + slot, JX_HEMPTY, key);
    if (prev == JX_HEMPTY || prev == key) return slot;
    slot = (slot + 1u) & mask;
  }
}
__device__ __forceinline__ int64_t hash_slot_find(const unsigned long long* K, uint32_t mask, uint32_t a, uint32_t b) {
  const unsigned long long key = ((unsigned long long)a << 32) | b;
  uint32_t slot = pair_hash(a, b) & mask;
  for (;;) {
    const unsigned long long k = K[slot];
    if (k == key) return slot;
    if (k == JX_HEMPTY) return -1;
    slot = (slot + 1u) & mask;
  }
}
// candidates C[0 .. *n): insert (max value per pair); Cslot[j] = the pair's slot
__global__ void k_pair_insert(const uint4* __restrict__ C, const uint32_t* __restrict__ n, const uint8_t* __restrict__ live,
                              unsigned long long* K, unsigned long long* V, uint32_t mask, uint32_t* Cslot) {
  const uint32_t N = *n;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    if (live && !live[j]) continue;
    const uint4 c = C[j];
    const uint32_t slot = hash_slot_insert(K, mask, c.x, c.y);
    atomicMax(V + slot, cand_value(c));
    Cslot[j] = slot;
  }
}
// win[j] = 1 iff candidate j is its pair's record; W[slot] = j for the winner
__global__ void k_pair_winner(const uint4* __restrict__ C, const uint32_t* __restrict__ n, const uint8_t* __restrict__ live,
                              const uint32_t* __restrict__ Cslot, const unsigned long long* __restrict__ V, uint32_t* W, uint8_t* win,
                              uint32_t* nwin) {
  const uint32_t N = *n;
  const uint32_t nround = (N + 31u) & ~31u;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += gridDim.x * blockDim.x) {
    bool w = false;
    if (j < N && (!live || live[j])) {
      const uint32_t slot = Cslot[j];
      w = V[slot] == cand_value(C[j]);
      if (w && W) W[slot] = j;
    }
    if (j < N) win[j] = w;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, w);       // distinct pairs (a statistic only)
    if (nwin && m && (threadIdx.x & 31) == 0) atomicAdd(nwin, (uint32_t)__popc(m));
  }
}
// tandem_grouper (blastfilter.py:262-284) without its sorts: among the winners with evalue < 1e-10, the hits of one
// gene sorted by partner rank are joined when consecutive ones lie on one seqid within Nmax ranks.  For the winner
// (q, s): its successor in q's list is the first (q, s + d), d = 1 .. Nmax, that is a flagged winner -- found by probing
// the pair table -- and symmetrically (q + d, s) for s's list.  One edge per side at most; the union-find then gives the
// same families (joining consecutive members or all members within Nmax has the same closure).
__global__ void k_tandem_edges(const uint4* __restrict__ C, const uint32_t* __restrict__ n, const uint8_t* __restrict__ win,
                               const unsigned long long* __restrict__ K, const uint32_t* __restrict__ W, uint32_t mask, int nmax,
                               uint32_t Gq, uint32_t Gs, const int32_t* __restrict__ qchr, const int32_t* __restrict__ schr,
                               uint2* qedges, uint2* sedges, uint32_t* nedges /* [0] q side, [1] s side */) {
  const uint32_t N = *n;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    if (!win[j]) continue;
    const uint4 c = C[j];
    if (!(c.w & 1u)) continue;
    const uint32_t q = c.x, s = c.y;
    const int32_t cs = schr[s], cq = qchr[q];
    for (int d = 1; d <= nmax; ++d) {                 // flip=False: per query, subject ranks ascending -> subject families
      const uint32_t s2 = s + (uint32_t)d;
      if (s2 >= Gs || schr[s2] != cs) break;
      const int64_t slot = hash_slot_find(K, mask, q, s2);
      if (slot >= 0 && (C[W[slot]].w & 1u)) { sedges[atomicAdd(nedges + 1, 1u)] = make_uint2(s, s2); break; }
    }
    for (int d = 1; d <= nmax; ++d) {                 // flip=True: per subject, query ranks ascending -> query families
      const uint32_t q2 = q + (uint32_t)d;
      if (q2 >= Gq || qchr[q2] != cq) break;
      const int64_t slot = hash_slot_find(K, mask, q2, s);
      if (slot >= 0 && (C[W[slot]].w & 1u)) { qedges[atomicAdd(nedges, 1u)] = make_uint2(q, q2); break; }
    }
  }
}
__global__ void k_apply_edges(const uint2* __restrict__ edges, const uint32_t* __restrict__ n, uint32_t* parent) {
  const uint32_t N = *n;
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) uf_union(parent, edges[t].x, edges[t].y);
}
// filter_tandem (blastfilter.py:240-259): winners -> (mother_q[q], mother_s[s]); `query == subject` pairs dropped; the
// substituted candidate D[j] enters the second pair table
__global__ void k_collapse_insert(const uint4* __restrict__ C, const uint32_t* __restrict__ n, const uint8_t* __restrict__ win,
                                  const int32_t* __restrict__ mq, const int32_t* __restrict__ ms, const int32_t* __restrict__ q_same_s,
                                  unsigned long long* K, unsigned long long* V, uint32_t mask, uint4* D, uint32_t* Dslot, uint8_t* live) {
  const uint32_t N = *n;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
    bool l = false;
    if (win[j]) {
      uint4 c = C[j];
      const int32_t a = mq[c.x], b = ms[c.y];
      if (q_same_s[a] != b) {                        // filter_tandem: `if b.query == b.subject: continue`
        c.x = (uint32_t)a; c.y = (uint32_t)b;
        D[j] = c;
        const uint32_t slot = hash_slot_insert(K, mask, c.x, c.y);
        atomicMax(V + slot, cand_value(c));
        Dslot[j] = slot;
        l = true;
      }
    }
    live[j] = l;
  }
}
// the survivors, in any order, with their output-order key (score desc, line asc)
__global__ void k_final_emit(const uint4* __restrict__ D, const uint32_t* __restrict__ n, const uint8_t* __restrict__ fin,
                             uint4* F, unsigned long long* keys, uint32_t* vals, uint32_t* nF) {
  const uint32_t N = *n;
  const uint32_t nround = (N + 31u) & ~31u;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nround; j += gridDim.x * blockDim.x) {
    const bool keep = j < N && fin[j];
    const uint32_t pos = warp_append(keep, nF);
    if (keep) {
      const uint4 c = D[j];
      F[pos] = c;
      keys[pos] = ((unsigned long long)(~score_ord(c.z)) << 32) | (c.w >> 1);
      vals[pos] = pos;
    }
  }
}

__global__ void k_iota(uint32_t* p, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
__global__ void k_flatten(uint32_t* parent, uint32_t n, uint32_t* root) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) root[i] = uf_find(parent, i);
}
__device__ __forceinline__ unsigned long long mother_key(int32_t len, uint32_t lexrank) {
  return ((unsigned long long)(uint32_t)len << 32) | (0xFFFFFFFFu - lexrank);
}
__global__ void k_mother1(const uint32_t* root, const int32_t* len, const uint32_t* lex, uint32_t n, unsigned long long* bestkey) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) atomicMax(bestkey + root[g], mother_key(len[g], lex[g]));
}
__global__ void k_mother2(const uint32_t* root, const int32_t* len, const uint32_t* lex, uint32_t n,
                          const unsigned long long* bestkey, uint32_t* mroot) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n && bestkey[root[g]] == mother_key(len[g], lex[g])) mroot[root[g]] = g;
}
__global__ void k_mother3(const uint32_t* root, const uint32_t* mroot, uint32_t n, int32_t* mother) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) mother[g] = (int32_t)mroot[root[g]];
}

// survivors in output order -> result columns + the byte offset of each survivor's line (from its record)
__global__ void k_gather_out(const uint32_t* __restrict__ order, uint32_t n, const uint4* __restrict__ F, const uint4* __restrict__ recs,
                             uint32_t ord_base, const LineIndex li, int32_t* oq, int32_t* os, float* oscore, uint64_t* ooff) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint4 c = F[order[t]];
  oq[t] = (int32_t)c.x;
  os[t] = (int32_t)c.y;
  oscore[t] = __uint_as_float(c.z);
  const uint32_t idx = (c.w >> 1) - ord_base;
  ooff[t] = jx_line_offset(li, idx, recs[idx].w);
}

__global__ void k_line_len(const uint8_t* text, int64_t n, const uint64_t* off, const uint32_t* idx, uint32_t cnt, uint32_t* len,
                           int rstrip = 0) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  int64_t p = (int64_t)off[idx ? idx[t] : t], e = p;
  while (e < n && !jx::is_term(text[e])) ++e;
  if (rstrip)                          // str.rstrip(): ASCII white space and the separators 0x1c-0x1f
    while (e > p && (jx::is_space(text[e - 1]) || (uint32_t)(text[e - 1] - 0x1Cu) <= 3u)) --e;
  len[t] = (uint32_t)(e - p) + 1;      // + '\n'
}
__global__ void k_copy_lines(const uint8_t* text, const uint64_t* off, const uint32_t* idx, const uint32_t* len,
                             const uint64_t* pos, uint32_t cnt, uint8_t* out) {
  uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= cnt) return;
  const uint8_t* src = text + off[idx ? idx[w] : w];
  uint8_t* dst = out + pos[w];
  uint32_t L = len[w] - 1;
  for (uint32_t i = lane; i < L; i += 32) dst[i] = src[i];
  if (lane == 0) dst[L] = '\n';
}

// ---- K12: exact `.filtered` writer on the device ------------------------------------------------
// Pass 1 (k_format_len): the block stages the first 208 bytes of its 128 lines in shared memory with coalesced 16-byte
// loads (a thread that reads its line byte by byte from global memory waits ~70 dependent round trips), every thread
// scans and formats the numeric tail of its line ONCE, keeps the tail (<= FMT_TAIL bytes) in a scratch row and returns
// the line's length.  Pass 2 (k_format_write), after the 64-bit prefix sum: names + kept tail into a shared staging
// area at the line's position inside the block's (contiguous) output range, then one coalesced copy to global memory.
constexpr int FMT_THREADS = 128;
constexpr int FMT_STAGE = 208;          // bytes staged per line (13 x 16); longer lines continue in global memory
constexpr int FMT_TAIL = 96;            // kept tail bytes per line; a longer tail is declined (host glibc formats the line)
constexpr int FMT_OUT = FMT_THREADS * 136;   // output staging per block; a block that needs more writes directly

struct StagedReader {          // bytes of one line: staged part in shared memory, the rest in global memory
  const uint8_t* s;            // shared copy of text[off ..]
  int avail;
  const uint8_t* g;            // text + off
  int64_t remain;
  __device__ __forceinline__ uint32_t operator()(int p) const {
    if ((int64_t)p >= remain) return (uint32_t)'\n';
    return p < avail ? s[p] : g[p];
  }
};

// formats the numeric tail of the line in `rd` into `tail` (<= 160 bytes); returns false -> glibc
template <class R>
__device__ bool format_tail(R& rd, char* tail, int& tail_len) {
  int e = 0;
  while (!jx::is_term(rd(e))) ++e;
  jx::LineFields f;
  jx::scan_fields(rd, 0, e, f);
  return jx::fmt_numeric_tail(tail, tail_len, f);
}

__global__ void __launch_bounds__(FMT_THREADS) k_format_len(const uint8_t* text, int64_t n, const uint64_t* off, const int32_t* q,
                                                            const int32_t* s, const uint32_t* qmeta, const uint32_t* smeta, uint32_t cnt,
                                                            uint32_t* len, uint8_t* tails, uint32_t* fallback) {
  __shared__ __align__(16) uint8_t s_in[FMT_THREADS * FMT_STAGE];
  const uint32_t t0 = blockIdx.x * FMT_THREADS;
  for (int j = threadIdx.x; j < FMT_THREADS * (FMT_STAGE / 16); j += FMT_THREADS) {
    const int line = j / (FMT_STAGE / 16), part = j - line * (FMT_STAGE / 16);
    const uint32_t t = t0 + (uint32_t)line;
    uint4 v = make_uint4(0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au);
    if (t < cnt) {
      const uint64_t a = (off[t] & ~15ull) + (uint64_t)part * 16u;      // the text buffer is 16-byte aligned
      if ((int64_t)a + 16 <= n) v = __ldg(reinterpret_cast<const uint4*>(text + a));
      else if ((int64_t)a < n) {                                        // the last, partial piece of the text
        uint32_t w[4] = {0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au, 0x0A0A0A0Au};
        for (int i = 0; i < 16 && (int64_t)a + i < n; ++i) w[i >> 2] = (w[i >> 2] & ~(0xFFu << (8 * (i & 3)))) | ((uint32_t)text[a + i] << (8 * (i & 3)));
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    reinterpret_cast<uint4*>(s_in)[j] = v;
  }
  __syncthreads();
  const uint32_t t = t0 + threadIdx.x;
  if (t >= cnt) return;
  const uint64_t o = off[t];
  const int sh = (int)(o & 15u);
  StagedReader rd{s_in + threadIdx.x * FMT_STAGE + sh, FMT_STAGE - sh, text + o, n - (int64_t)o};
  char tail[160];
  int tl = 0;
  if (!format_tail(rd, tail, tl) || tl > FMT_TAIL) { atomicAdd(fallback, 1u); len[t] = 0; return; }   // declined: host formats it
  uint8_t* row = tails + (size_t)t * FMT_TAIL;
  for (int i = 0; i < tl; ++i) row[i] = (uint8_t)tail[i];
  len[t] = (qmeta[q[t]] & 127u) + 1u + (smeta[s[t]] & 127u) + (uint32_t)tl;
}

__global__ void __launch_bounds__(FMT_THREADS) k_format_write(const int32_t* q, const int32_t* s, const uint32_t* qmeta, const uint32_t* qblob,
                                                              const uint32_t* smeta, const uint32_t* sblob, uint32_t cnt, const uint32_t* len,
                                                              const uint64_t* pos, const uint8_t* tails, char* out) {
  __shared__ uint8_t s_out[FMT_OUT];
  const uint32_t t0 = blockIdx.x * FMT_THREADS, t1 = min(cnt, t0 + FMT_THREADS);
  const uint64_t base = pos[t0];
  const uint64_t total = pos[t1 - 1] + len[t1 - 1] - base;        // the block's lines are one contiguous output range
  const bool staged = total <= (uint64_t)FMT_OUT;
  const uint32_t t = t0 + threadIdx.x;
  if (t < cnt && len[t]) {
    uint8_t* o = staged ? s_out + (pos[t] - base) : reinterpret_cast<uint8_t*>(out) + pos[t];
    uint32_t m = qmeta[q[t]], l = m & 127u;
    const uint32_t* w = qblob + (m >> 7);
    for (uint32_t i = 0; i < l; ++i) *o++ = (uint8_t)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
    *o++ = (uint8_t)'\t';
    m = smeta[s[t]];
    const uint32_t l2 = m & 127u;
    w = sblob + (m >> 7);
    for (uint32_t i = 0; i < l2; ++i) *o++ = (uint8_t)((w[i >> 2] >> (8 * (i & 3))) & 0xFFu);
    const uint32_t tl = len[t] - l - 1u - l2;
    const uint8_t* row = tails + (size_t)t * FMT_TAIL;
    for (uint32_t i = 0; i < tl; ++i) *o++ = row[i];
  }
  if (!staged) return;
  __syncthreads();
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + base;
  for (uint32_t i = threadIdx.x; i < (uint32_t)total; i += FMT_THREADS) dst[i] = s_out[i];
}

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

// the reference's own formatting of one survivor: sscanf with cblast.pyx:15, start/stop
// normalisation (cblast.pyx:115-120), sprintf with cblast.pyx:17 and the substituted names
void format_range(const char* lines, const uint64_t* pos, const uint32_t* len, const int32_t* q_all, const int32_t* s_all,
                  const uint32_t* idx, size_t lo, size_t hi, const SideDev& Q, const SideDev& S, std::string& out) {
  char qn[160], sn[160], buf[1024];
  std::string line;
  for (size_t t = lo; t < hi; ++t) {
    const int32_t* q = q_all + (idx ? idx[t] : t) - t;   // q[t] / s[t] below address the survivor
    const int32_t* s = s_all + (idx ? idx[t] : t) - t;
    line.assign(lines + pos[t], len[t]);
    float pct = 0, score = 0; double ev = 0; int v[7] = {0, 0, 0, 0, 0, 0, 0};
    qn[0] = sn[0] = 0;
    sscanf(line.c_str(), "%127s\t%127s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f", qn, sn, &pct, &v[0], &v[1], &v[2],
           &v[3], &v[4], &v[5], &v[6], &ev, &score);
    if (v[3] > v[4]) std::swap(v[3], v[4]);
    if (v[5] > v[6]) std::swap(v[5], v[6]);
    const uint32_t qa = Q.h_off[(size_t)q[t]], qb = Q.h_off[(size_t)q[t] + 1];
    const uint32_t sa = S.h_off[(size_t)s[t]], sb = S.h_off[(size_t)s[t] + 1];
    out.append(Q.h_names.data() + qa, qb - qa);
    out.push_back('\t');
    out.append(S.h_names.data() + sa, sb - sa);
    int w = snprintf(buf, sizeof buf, "\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.3g\n", pct, v[0], v[1], v[2], v[3],
                     v[4], v[5], v[6], ev, score);
    out.append(buf, (size_t)w);
  }
}

}  // namespace

namespace {
__global__ void k_cscore_flags(const uint4* __restrict__ recs, uint32_t n, const uint32_t* __restrict__ best,
                               const uint32_t* __restrict__ qnid, const uint32_t* __restrict__ snid, double cscore,
                               int use_cscore, uint32_t* flags, uint32_t* errflag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = __ldcs(recs + i);
  bool keep = r.x != JX_NOREC;
  if (keep && use_cscore) {
    double bq = (double)__uint_as_float(__ldg(best + __ldg(qnid + r.x)));
    double bs = (double)__uint_as_float(__ldg(best + __ldg(snid + r.y)));
    double den = bq > bs ? bq : bs;
    if (den == 0.0) { atomicExch(errflag, 1u); keep = false; }
    else keep = ((double)__uint_as_float(r.z) / den) > cscore;
  }
  flags[i] = keep;
}
__global__ void k_compact_idx(const uint32_t* flags, const uint32_t* pos, uint32_t n, uint32_t* idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) idx[pos[i]] = i;
}
__global__ void k_line_off(const uint32_t* idx, uint32_t cnt, const uint4* recs, const LineIndex li, uint64_t* off) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const uint64_t i = idx[t];
  off[t] = jx_line_offset(li, i, recs[i].w);
}
}  // namespace

int jx_gather_lines(jx_ctx* ctx, Arena& arena, const uint8_t* d_text, int64_t nbytes, const LineIndex& li,
                    const uint4* recs, const uint32_t* rec_idx, uint32_t cnt, char** out, int64_t* out_len, int rstrip, uint64_t* dev_out) {
  if (out) *out = nullptr;
  *out_len = 0;
  if (dev_out) *dev_out = 0;
  if (!cnt) { if (!dev_out) *out = (char*)calloc(1, 1); return JX_OK; }
  int rc;
  JX_ALLOC(off, uint64_t, cnt, false);
  JX_ALLOC(len, uint32_t, cnt, false);
  JX_ALLOC(pos, uint64_t, cnt, false);
  JX_LAUNCH(k_line_off, nblk(cnt), 256, 0, rec_idx, cnt, recs, li, off);
  JX_LAUNCH(k_line_len, nblk(cnt), 256, 0, d_text, nbytes, off, (const uint32_t*)nullptr, cnt, len, rstrip);
  // byte positions are 64-bit: the packed lines of a big slice may exceed 4 GiB (C-score filter off)
  if ((rc = jx_exclusive_sum64(ctx, arena, len, pos, cnt))) return rc;
  uint64_t hpos = 0; uint32_t hlen = 0;
  JX_CK(cudaMemcpyAsync(&hpos, pos + cnt - 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(&hlen, len + cnt - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  const uint64_t total = hpos + hlen;
  if (dev_out) {                                   // context-owned device result (valid until the next one)
    if (ctx->lines_dev_cap < total + 16) {
      JX_CK(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->lines_dev); ctx->lines_dev = nullptr; ctx->lines_dev_cap = 0;
      const size_t cap = (size_t)(total + total / 4 + (1u << 20));
      JX_CK(cudaMalloc((void**)&ctx->lines_dev, cap));
      ctx->lines_dev_cap = cap;
    }
    JX_LAUNCH(k_copy_lines, nblk((size_t)cnt * 32), 256, 0, d_text, off, (const uint32_t*)nullptr, len, pos, cnt, ctx->lines_dev);
    JX_CK(cudaStreamSynchronize(ctx->stream));
    *dev_out = (uint64_t)(uintptr_t)ctx->lines_dev; *out_len = (int64_t)total;
    return JX_OK;
  }
  JX_ALLOC(packed, uint8_t, (size_t)total, false);
  JX_LAUNCH(k_copy_lines, nblk((size_t)cnt * 32), 256, 0, d_text, off, (const uint32_t*)nullptr, len, pos, cnt, packed);
  char* host = (char*)malloc((size_t)total + 1);
  if (!host) { ctx->err = "out of host memory"; return JX_EINTERNAL; }
  if ((rc = jx_d2h_big(ctx, host, packed, (size_t)total))) { free(host); return rc; }
  host[total] = 0;
  *out = host; *out_len = (int64_t)total;
  return JX_OK;
}

extern "C" {

int jx_cscore_begin(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, int strip_names,
                    const uint64_t* exclude_keys, int64_t n_exclude, uint64_t* best_ptr, uint32_t* n_ids) {
  if (!ctx || !best_ptr || !n_ids) return JX_EARG;
  cudaSetDevice(ctx->device);
  ctx->sess_open = false;
  int rc;
  // the slice must outlive this call: hold it on the device (jx_text_upload) unless it already is
  uint64_t dptr = 0;
  if (!text_on_device) {
    if ((rc = jx_text_upload(ctx, text, nbytes, &dptr))) return rc;
  } else if (!(ctx->held_text && (const uint8_t*)text == ctx->held_text && nbytes == ctx->held_n)) {
    ctx->err = "jx_cscore_begin: device text must come from jx_text_upload"; return JX_EARG;
  } else dptr = (uint64_t)(uintptr_t)text;
  Arena arena(ctx);
  if (ctx->sess_best_cap < (size_t)ctx->n_name_ids + 1) {
    cudaFree(ctx->sess_best); ctx->sess_best = nullptr; ctx->sess_best_cap = 0;
    JX_CK(cudaMalloc((void**)&ctx->sess_best, ((size_t)ctx->n_name_ids + 1) * 4));
    ctx->sess_best_cap = (size_t)ctx->n_name_ids + 1;
  }
  JX_CK(cudaMemsetAsync(ctx->sess_best, 0, ((size_t)ctx->n_name_ids + 1) * 4, ctx->stream));
  TokMode mode;
  mode.strip = strip_names;
  mode.best = ctx->sess_best;
  mode.persist = true;
  JX_ALLOC(d_excl, uint64_t, (size_t)std::max<int64_t>(n_exclude, 1), false);
  if (n_exclude > 0) {
    JX_CK(cudaMemcpyAsync(d_excl, exclude_keys, (size_t)n_exclude * 8, cudaMemcpyHostToDevice, ctx->stream));
    mode.excl = d_excl; mode.n_excl = n_exclude;
  }
  TokOut tok;
  if ((rc = jx_run_tokenizer(ctx, arena, (const char*)(uintptr_t)dptr, nbytes, 1, mode, tok))) return rc;
  if (!ctx->cache.valid) { ctx->err = "jx_cscore_begin: records were not retained"; return JX_EINTERNAL; }
  JX_CK(cudaStreamSynchronize(ctx->stream));
  ctx->sess_open = true;
  *best_ptr = (uint64_t)(uintptr_t)ctx->sess_best;
  *n_ids = ctx->n_name_ids;
  return JX_OK;
}

static int cscore_finish_impl(jx_ctx* ctx, double cscore, char** lines, int64_t* lines_len, int64_t* n_lines, uint64_t* dev_out);
int jx_cscore_finish(jx_ctx* ctx, double cscore, char** lines, int64_t* lines_len, int64_t* n_lines) {
  if (!ctx || !lines || !lines_len || !n_lines) return JX_EARG;
  return cscore_finish_impl(ctx, cscore, lines, lines_len, n_lines, nullptr);
}
int jx_cscore_finish_device(jx_ctx* ctx, double cscore, uint64_t* lines_dev, int64_t* lines_len, int64_t* n_lines) {
  if (!ctx || !lines_dev || !lines_len || !n_lines) return JX_EARG;
  char* unused = nullptr;
  return cscore_finish_impl(ctx, cscore, &unused, lines_len, n_lines, lines_dev);
}
static int cscore_finish_impl(jx_ctx* ctx, double cscore, char** lines, int64_t* lines_len, int64_t* n_lines, uint64_t* dev_out) {
  if (!ctx->sess_open || !ctx->cache.valid) { ctx->err = "jx_cscore_finish without jx_cscore_begin"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  auto& c = ctx->cache;
  const uint32_t n = c.n_lines;
  *lines = nullptr; *lines_len = 0; *n_lines = 0;
  int rc;
  uint32_t cnt = 0;
  uint32_t* idx = nullptr;
  if (n) {
    JX_ALLOC(flags, uint32_t, n, false);
    JX_ALLOC(pos, uint32_t, n, false);
    JX_ALLOC(err, uint32_t, 1, true);
    JX_LAUNCH(k_cscore_flags, nblk(n), 256, 0, c.recs, n, ctx->sess_best, ctx->side[0].name_id, ctx->side[1].name_id, cscore,
              cscore != 0.0 ? 1 : 0, flags, err);
    if ((rc = jx_exclusive_sum(ctx, arena, flags, pos, n))) return rc;
    uint32_t h[3];
    JX_CK(cudaMemcpyAsync(&h[0], pos + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaMemcpyAsync(&h[1], flags + n - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaMemcpyAsync(&h[2], err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    if (h[2]) { ctx->err = "float division by zero"; return JX_EZERODIV; }
    cnt = h[0] + h[1];
    JX_ALLOC(ix, uint32_t, cnt, false);
    if (cnt) JX_LAUNCH(k_compact_idx, nblk(n), 256, 0, flags, pos, n, ix);
    idx = ix;
  }
  LineIndex li; li.desc = c.desc; li.ntiles = c.ntiles; li.cap = c.cap; li.sub_bytes = c.sub_bytes;
  rc = jx_gather_lines(ctx, arena, c.text, c.nbytes, li, c.recs, idx, cnt, lines, lines_len, 0, dev_out);
  if (rc) return rc;
  *n_lines = cnt;
  ctx->sess_open = false;
  c.valid = c.plain;          // plain records stay for jx_near_lines / jx_liftover_text on the same held text
  return JX_OK;
}

void jx_filter_result_free(jx_filter_result* r) {
  if (!r) return;
  free(r->q_rank); free(r->s_rank); free(r->score); free(r->line_off); free(r->q_mother); free(r->s_mother);
  if (!r->text_pinned) free(r->text);
  else if (r->owner && r->text) jx_pin_release((jx_ctx*)r->owner, r->text);   // back to the context's pool
  memset(r, 0, sizeof *r);
}

int jx_blastfilter(jx_ctx* ctx, const char* text, int64_t nbytes, int text_on_device, const jx_filter_opts* opts,
                   jx_filter_result* out) {
  if (!ctx || !opts || !out) return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  const bool use_cscore = opts->cscore != 0.0;
  const bool use_tandem = opts->tandem_nmax != 0;
  int rc;

  JX_TRACE("bf: enter");
  // ---- K1/K2 (+ K4 fused): text -> records, per-name best score -------------------------------
  TokMode mode;
  mode.strip = opts->strip_names;
  JX_ALLOC(d_excl, uint64_t, (size_t)std::max<int64_t>(opts->n_exclude, 1), false);
  if (opts->n_exclude > 0) {
    JX_CK(cudaMemcpyAsync(d_excl, opts->exclude_keys, (size_t)opts->n_exclude * 8, cudaMemcpyHostToDevice, ctx->stream));
    mode.excl = d_excl; mode.n_excl = opts->n_exclude;
  }
  uint32_t* d_best = nullptr;
  if (use_cscore) {
    JX_ALLOC(b, uint32_t, (size_t)ctx->n_name_ids + 1, true);
    d_best = b; mode.best = b;
  }
  mode.fill_cache = ctx->record_reuse;      // jx_liftover_text on the same held text reuses these records
  TokOut tok;
  if ((rc = jx_run_tokenizer(ctx, arena, text, nbytes, text_on_device, mode, tok))) return rc;
  out->stats[0] = (int64_t)(tok.stats[0] - tok.stats[1]);
  out->stats[7] = (int64_t)tok.stats[1];
  out->stats[2] = (int64_t)tok.stats[3];
  out->stats[1] = (int64_t)(tok.stats[2] + tok.stats[8]);
  out->stats[3] = (int64_t)tok.stats[2];
  const uint32_t nrec = tok.n_lines;
  JX_TRACE("bf: tokenized");

  // ---- K5: C-score predicate fused with compaction -------------------------------------------------
  if ((uint64_t)nrec >= (1ull << 31)) { ctx->err = "more than 2^31 record slots"; return JX_EUNSUPPORTED; }
  const size_t capC = (size_t)tok.stats[2] + 1;
  JX_ALLOC(C, uint4, capC, false);
  JX_ALLOC(d_cnt, uint32_t, 8, true);          // 0 candidates, 1 zero-division flag, 2 q edges, 3 s edges, 4 survivors
  if (nrec) {
    unsigned grid = std::min<unsigned>(nblk((nrec + 3) / 4), 148u * JX_SEL_GRID);
    JX_ALLOC(bestq, uint32_t, (size_t)Q.G + 1, false);
    JX_ALLOC(bests, uint32_t, (size_t)S.G + 1, false);
    if (use_cscore) {
      JX_LAUNCH(k_gene_best, nblk((size_t)Q.G), 256, 0, d_best, Q.name_id, (uint32_t)Q.G, bestq);
      JX_LAUNCH(k_gene_best, nblk((size_t)S.G), 256, 0, d_best, S.name_id, (uint32_t)S.G, bests);
    }
    JX_LAUNCH(k_select, grid, 256, 0, tok.recs, nrec, bestq, bests, opts->cscore, use_cscore ? 1 : 0, 0u, C, d_cnt, d_cnt + 1);
  }
  uint32_t h_cnt[2] = {0, 0};
  JX_CK(cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (h_cnt[1]) { ctx->err = "float division by zero"; return JX_EZERODIV; }
  const uint32_t nC = h_cnt[0];
  out->stats[4] = nC;
  JX_TRACE("bf: k_select");

  // ---- K3, K6, K7 on the candidates; everything below is sized by nC and runs without a host round trip ------------
  uint32_t nF = 0;
  uint4* F = nullptr; uint32_t* Forder = nullptr;
  int32_t *d_mq = nullptr, *d_ms = nullptr;
  if (use_tandem) {                                 // (the maps are also a result: want_mothers)
    const uint32_t Gq = (uint32_t)Q.G, Gs = (uint32_t)S.G;
    JX_ALLOC(mq0, int32_t, (size_t)Gq + 1, false);
    d_mq = mq0;
    if (ctx->is_self) d_ms = d_mq;
    else { JX_ALLOC(ms0, int32_t, (size_t)Gs + 1, false); d_ms = ms0; }
  }
  if (nC || use_tandem) {
    uint32_t hcap = 1024;
    while (hcap < 2u * nC + 16u && hcap < (1u << 31)) hcap <<= 1;
    const uint32_t hmask = hcap - 1u;
    const unsigned g1 = std::max(1u, std::min<unsigned>(nblk(std::max<uint32_t>(nC, 1u)), 148u * 8u));
    JX_ALLOC(K1, unsigned long long, hcap, false);
    JX_ALLOC(V1, unsigned long long, hcap, true);
    JX_ALLOC(W1, uint32_t, hcap, false);
    JX_ALLOC(Cslot, uint32_t, (size_t)nC + 1, false);
    JX_ALLOC(win, uint8_t, (size_t)nC + 1, false);
    JX_CK(cudaMemsetAsync(K1, 0xFF, (size_t)hcap * 8, ctx->stream));
    JX_LAUNCH(k_pair_insert, g1, 256, 0, C, d_cnt, (const uint8_t*)nullptr, K1, V1, hmask, Cslot);
    JX_LAUNCH(k_pair_winner, g1, 256, 0, C, d_cnt, (const uint8_t*)nullptr, Cslot, V1, W1, win, d_cnt + 5);
    const uint4* D = C; const uint8_t* fin = win;
    if (use_tandem) {
      const uint32_t Gq = (uint32_t)Q.G, Gs = (uint32_t)S.G;
      JX_ALLOC(parent_q, uint32_t, (size_t)Gq + 1, false);
      uint32_t* parent_s = parent_q;
      JX_LAUNCH(k_iota, nblk(std::max(Gq, 1u)), 256, 0, parent_q, Gq);
      if (!ctx->is_self) {
        JX_ALLOC(ps, uint32_t, (size_t)Gs + 1, false);
        parent_s = ps;
        JX_LAUNCH(k_iota, nblk(std::max(Gs, 1u)), 256, 0, parent_s, Gs);
      }
      JX_ALLOC(qedges, uint2, (size_t)nC + 1, false);
      JX_ALLOC(sedges, uint2, (size_t)nC + 1, false);
      JX_LAUNCH(k_tandem_edges, g1, 256, 0, C, d_cnt, win, K1, W1, hmask, opts->tandem_nmax, Gq, Gs, Q.chr_lex, S.chr_lex,
                qedges, sedges, d_cnt + 2);
      JX_LAUNCH(k_apply_edges, g1, 256, 0, qedges, d_cnt + 2, parent_q);
      JX_LAUNCH(k_apply_edges, g1, 256, 0, sedges, d_cnt + 3, parent_s);
      auto mothers = [&](uint32_t* parent, const SideDev& Dv, int32_t* m) -> int {
        const uint32_t G = (uint32_t)Dv.G;
        JX_ALLOC(root, uint32_t, (size_t)G + 1, false);
        JX_ALLOC(bk, unsigned long long, (size_t)G + 1, true);
        JX_ALLOC(mroot, uint32_t, (size_t)G + 1, false);
        if (G) {
          JX_LAUNCH(k_flatten, nblk(G), 256, 0, parent, G, root);
          JX_LAUNCH(k_mother1, nblk(G), 256, 0, root, Dv.length, Dv.lexrank, G, bk);
          JX_LAUNCH(k_mother2, nblk(G), 256, 0, root, Dv.length, Dv.lexrank, G, bk, mroot);
          JX_LAUNCH(k_mother3, nblk(G), 256, 0, root, mroot, G, m);
        }
        return JX_OK;
      };
      if ((rc = mothers(parent_q, Q, d_mq))) return rc;
      if (!ctx->is_self && (rc = mothers(parent_s, S, d_ms))) return rc;
      // ---- K7: substitute mothers, drop q == s, second per-pair best ---------------------------------------------------
      JX_ALLOC(K2, unsigned long long, hcap, false);
      JX_ALLOC(V2, unsigned long long, hcap, true);
      JX_ALLOC(D2, uint4, (size_t)nC + 1, false);
      JX_ALLOC(Dslot, uint32_t, (size_t)nC + 1, false);
      JX_ALLOC(live, uint8_t, (size_t)nC + 1, false);
      JX_ALLOC(fin2, uint8_t, (size_t)nC + 1, false);
      JX_CK(cudaMemsetAsync(K2, 0xFF, (size_t)hcap * 8, ctx->stream));
      JX_LAUNCH(k_collapse_insert, g1, 256, 0, C, d_cnt, win, d_mq, d_ms, ctx->q_same_s, K2, V2, hmask, D2, Dslot, live);
      JX_LAUNCH(k_pair_winner, g1, 256, 0, D2, d_cnt, live, Dslot, V2, (uint32_t*)nullptr, fin2, (uint32_t*)nullptr);
      D = D2; fin = fin2;
    }
    // ---- K8: the survivors and their output order (score desc, line asc) ----------------------------------------------------
    JX_ALLOC(F0, uint4, (size_t)nC + 1, false);
    JX_ALLOC(okeys, uint64_t, (size_t)nC + 1, false);
    JX_ALLOC(ovals, uint32_t, (size_t)nC + 1, false);
    JX_LAUNCH(k_final_emit, g1, 256, 0, D, d_cnt, fin, F0, (unsigned long long*)okeys, ovals, d_cnt + 4);
    uint32_t h5[6] = {0, 0, 0, 0, 0, 0};
    JX_CK(cudaMemcpyAsync(h5, d_cnt, 24, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    nF = h5[4];
    out->stats[5] = h5[5];
    F = F0;
    uint64_t* ok = okeys; uint32_t* ov = ovals;
    if (nF && (rc = jx_sort_pairs(ctx, arena, ok, ov, nF, 0, 64))) return rc;
    Forder = ov;
  }
  out->stats[6] = nF;
  JX_TRACE("bf: pairs+tandem+order");

  out->n = nF;
  const bool host_arrays = !opts->skip_arrays;
  if (host_arrays) {
    out->q_rank = (int32_t*)malloc(sizeof(int32_t) * (nF + 1));
    out->s_rank = (int32_t*)malloc(sizeof(int32_t) * (nF + 1));
    out->score = (float*)malloc(sizeof(float) * (nF + 1));
    out->line_off = (uint64_t*)malloc(sizeof(uint64_t) * (nF + 1));
  }
  ctx->lastf.n = 0; ctx->lastf.token = 0;
  if (nF) {
    JX_ALLOC(oq, int32_t, nF, false);
    JX_ALLOC(os, int32_t, nF, false);
    JX_ALLOC(osc, float, nF, false);
    JX_ALLOC(ooff, uint64_t, nF, false);
    JX_LAUNCH(k_gather_out, nblk(nF), 256, 0, Forder, nF, F, tok.recs, 0u, tok.li, oq, os, osc, ooff);
    if (host_arrays) {
      JX_CK(cudaMemcpyAsync(out->q_rank, oq, nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(out->s_rank, os, nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(out->score, osc, nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(out->line_off, ooff, nF * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    {   // keep the survivors on the device for jx_scan_filtered
      auto& L = ctx->lastf;
      if (L.cap < nF) {
        cudaFree(L.q); cudaFree(L.s); cudaFree(L.score);
        L.q = nullptr; L.s = nullptr; L.score = nullptr; L.cap = 0;
        const size_t cap = (size_t)nF + (nF >> 2) + 1024;
        JX_CK(cudaMalloc((void**)&L.q, cap * 4));
        JX_CK(cudaMalloc((void**)&L.s, cap * 4));
        JX_CK(cudaMalloc((void**)&L.score, cap * 4));
        L.cap = cap;
      }
      JX_CK(cudaMemcpyAsync(L.q, oq, (size_t)nF * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      JX_CK(cudaMemcpyAsync(L.s, os, (size_t)nF * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      JX_CK(cudaMemcpyAsync(L.score, osc, (size_t)nF * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      L.n = nF; L.token = ++ctx->token_counter;
      out->device_token = L.token;
    }
    if (opts->want_text) {
      // K12: format on the device (exact re-implementation of the "%.2f/%.2g/%.3g" conversions);
      // anything it declines (inf/nan, >int32 integers, decimal rounding ties in the evalue ...)
      // sends the whole output through the host's glibc instead.
      JX_ALLOC(flen, uint32_t, nF, false);
      JX_ALLOC(fpos, uint64_t, nF, false);
      JX_ALLOC(ffb, uint32_t, 1, true);
      JX_ALLOC(ftails, uint8_t, (size_t)nF * FMT_TAIL, false);
      JX_LAUNCH(k_format_len, nblk(nF, FMT_THREADS), FMT_THREADS, 0, tok.d_text, tok.nbytes, ooff, oq, os, Q.name_meta, S.name_meta, nF,
                flen, ftails, ffb);
      // 64-bit byte positions: with the C-score filter off the `.filtered` text of a big table exceeds 4 GiB
      if ((rc = jx_exclusive_sum64(ctx, arena, flen, fpos, nF))) return rc;
      uint32_t h[2] = {0, 0};
      uint64_t hlast = 0;
      JX_CK(cudaMemcpyAsync(&h[0], ffb, 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(&hlast, fpos + nF - 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaMemcpyAsync(&h[1], flen + nF - 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
      JX_CK(cudaStreamSynchronize(ctx->stream));
      const uint32_t ndecl = h[0];
      const uint64_t total = hlast + h[1];
      // jx_pipeline (ctx->defer_text): the text goes to a context-owned buffer and crosses PCIe on the copy stream while
      // the next stages run; the caller waits for the copy stream before it returns
      const bool defer = ctx->defer_text && !ndecl;
      char* dtext = nullptr;
      if (defer) {
        if (ctx->ftext_dev_cap < (size_t)total + 1) {
          cudaFree(ctx->ftext_dev);
          ctx->ftext_dev = nullptr; ctx->ftext_dev_cap = 0;
          const size_t cap = (size_t)total + (size_t)(total >> 2) + 4096;
          JX_CK(cudaMalloc((void**)&ctx->ftext_dev, cap));
          ctx->ftext_dev_cap = cap;
        }
        dtext = (char*)ctx->ftext_dev;
      } else {
        JX_ALLOC(dt0, char, (size_t)total + 1, false);
        dtext = dt0;
      }
      JX_LAUNCH(k_format_write, nblk(nF, FMT_THREADS), FMT_THREADS, 0, oq, os, Q.name_meta, Q.blobw, S.name_meta, S.blobw, nF, flen, fpos,
                ftails, dtext);
      char* host = jx_pin_acquire(ctx, (size_t)total + 1);   // borrowed from the pool until the result is freed
      if (!host) { ctx->err = "cudaMallocHost failed"; return JX_ECUDA; }
      struct PinGuard { jx_ctx* c; char* p; ~PinGuard() { if (p) jx_pin_release(c, p); } } guard{ctx, host};
      if (defer) {
        if (!ctx->ev_text) JX_CK(cudaEventCreateWithFlags(&ctx->ev_text, cudaEventDisableTiming));
        JX_CK(cudaEventRecord(ctx->ev_text, ctx->stream));
        JX_CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_text, 0));
        JX_CK(cudaMemcpyAsync(host, dtext, (size_t)total, cudaMemcpyDeviceToHost, ctx->copy_stream));
        ctx->text_in_flight = true;
        host[total] = 0;                               // (the copy writes [0, total) only)
        out->text = host;
        out->text_pinned = 2;
        out->owner = ctx;
        out->text_len = (int64_t)total;
        guard.p = nullptr;                             // owned by the result
      } else {
      JX_CK(cudaMemcpyAsync(host, dtext, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
      if (!ndecl) {
        JX_CK(cudaStreamSynchronize(ctx->stream));
        host[total] = 0;
        out->text = host;
        out->text_pinned = 2;
        out->owner = ctx;
        out->text_len = (int64_t)total;
        guard.p = nullptr;                           // now owned by the result
      } else {
        // a few lines were declined by the device writer: format exactly those with the host's
        // glibc (the reference's own sscanf/sprintf pair, cblast.pyx:15,17) and splice them in
        std::vector<uint32_t> hlen_all(nF); std::vector<uint64_t> hpos_all(nF);
        JX_CK(cudaMemcpyAsync(hlen_all.data(), flen, (size_t)nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
        JX_CK(cudaMemcpyAsync(hpos_all.data(), fpos, (size_t)nF * 8, cudaMemcpyDeviceToHost, ctx->stream));
        JX_CK(cudaStreamSynchronize(ctx->stream));
        std::vector<int32_t> tmpq, tmps;
        const int32_t* hq = out->q_rank; const int32_t* hs = out->s_rank;
        if (!host_arrays) {
          tmpq.resize(nF); tmps.resize(nF);
          JX_CK(cudaMemcpyAsync(tmpq.data(), oq, (size_t)nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
          JX_CK(cudaMemcpyAsync(tmps.data(), os, (size_t)nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
          JX_CK(cudaStreamSynchronize(ctx->stream));
          hq = tmpq.data(); hs = tmps.data();
        }
        std::vector<uint32_t> didx;
        didx.reserve(ndecl);
        for (uint32_t t = 0; t < nF; ++t) if (!hlen_all[t]) didx.push_back(t);
        const uint32_t nd = (uint32_t)didx.size();
        JX_ALLOC(d_idx, uint32_t, nd, false);
        JX_ALLOC(len, uint32_t, nd, false);
        JX_ALLOC(pos, uint64_t, nd, false);
        JX_CK(cudaMemcpyAsync(d_idx, didx.data(), (size_t)nd * 4, cudaMemcpyHostToDevice, ctx->stream));
        JX_LAUNCH(k_line_len, nblk(nd), 256, 0, tok.d_text, tok.nbytes, ooff, d_idx, nd, len);
        if ((rc = jx_exclusive_sum64(ctx, arena, len, pos, nd))) return rc;
        std::vector<uint32_t> hlen(nd); std::vector<uint64_t> hpos(nd);
        JX_CK(cudaMemcpyAsync(hlen.data(), len, (size_t)nd * 4, cudaMemcpyDeviceToHost, ctx->stream));
        JX_CK(cudaMemcpyAsync(hpos.data(), pos, (size_t)nd * 8, cudaMemcpyDeviceToHost, ctx->stream));
        JX_CK(cudaStreamSynchronize(ctx->stream));
        const uint64_t ltotal = hpos[nd - 1] + hlen[nd - 1];
        JX_ALLOC(packed, uint8_t, (size_t)ltotal, false);
        JX_LAUNCH(k_copy_lines, nblk((size_t)nd * 32), 256, 0, tok.d_text, ooff, d_idx, len, pos, nd, packed);
        std::vector<char> hl((size_t)ltotal);
        JX_CK(cudaMemcpyAsync(hl.data(), packed, (size_t)ltotal, cudaMemcpyDeviceToHost, ctx->stream));
        JX_CK(cudaStreamSynchronize(ctx->stream));
        unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (nd < 4096) nt = 1;
        std::vector<std::string> parts(nt);
        std::vector<size_t> part_lo(nt + 1, 0);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) {
          size_t lo = (size_t)nd * t / nt, hi = (size_t)nd * (t + 1) / nt;
          part_lo[t] = lo; part_lo[t + 1] = hi;
          th.emplace_back([&, t, lo, hi]() {
            format_range(hl.data(), hpos.data(), hlen.data(), hq, hs, didx.data(), lo, hi, Q, S, parts[t]);
          });
        }
        for (auto& x : th) x.join();
        std::string merged;
        size_t extra = 0;
        for (auto& p : parts) extra += p.size();
        merged.reserve((size_t)total + extra + 1);
        // every host-formatted line ends with '\n': walk the parts line by line while copying the
        // device-formatted runs in between
        unsigned pi = 0; size_t po = 0;
        uint64_t copied_to = 0;                       // device text consumed so far
        for (uint32_t k = 0; k < nd; ++k) {
          const uint32_t t = didx[k];
          const uint64_t upto = hpos_all[t];          // device bytes that precede survivor t
          merged.append(host + copied_to, (size_t)(upto - copied_to));
          copied_to = upto;
          while (pi < nt && k >= part_lo[pi + 1]) { ++pi; po = 0; }
          const std::string& P = parts[pi];
          const size_t nl = P.find('\n', po);
          merged.append(P, po, nl - po + 1);
          po = nl + 1;
        }
        merged.append(host + copied_to, (size_t)(total - copied_to));
        out->text = (char*)malloc(merged.size() + 1);
        memcpy(out->text, merged.data(), merged.size());
        out->text[merged.size()] = 0;
        out->text_len = (int64_t)merged.size();
        out->text_pinned = 0;
      }
      }
    }
  } else if (opts->want_text) {
    out->text = (char*)calloc(1, 1);
    out->text_len = 0;
  }
  JX_TRACE("bf: order+gather+text");
  if (opts->want_mothers) {
    out->q_mother = (int32_t*)malloc(sizeof(int32_t) * (size_t)(Q.G + 1));
    out->s_mother = (int32_t*)malloc(sizeof(int32_t) * (size_t)(S.G + 1));
    if (use_tandem) {
      if (Q.G) JX_CK(cudaMemcpyAsync(out->q_mother, d_mq, (size_t)Q.G * 4, cudaMemcpyDeviceToHost, ctx->stream));
      if (S.G) JX_CK(cudaMemcpyAsync(out->s_mother, d_ms, (size_t)S.G * 4, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
      for (int64_t g = 0; g < Q.G; ++g) out->q_mother[g] = (int32_t)g;
      for (int64_t g = 0; g < S.G; ++g) out->s_mother[g] = (int32_t)g;
    }
  }
  JX_CK(cudaStreamSynchronize(ctx->stream));
  JX_CK(cudaGetLastError());
  return JX_OK;
}


}  // extern "C"

// =====================================================================================================================
// One genome pair sharded over ranks by CHROMOSOME PAIR (SURVEY.md 8(e) stages B / C).
//
// The reference's loops are independent per chromosome pair from the pair de-duplication on: (qi, si) determines the
// pair of seqids, tandem families never leave a seqid (blastfilter.py:262-284), `synteny_scan` and `synteny_liftover`
// run per (qseqid, sseqid) (synteny.py:444-450, 1946-1956).  What is global: the per-name best score (reduced with
// ncclAllReduce(max) before these calls) and the tandem families (a subject gene collects evidence from every query
// chromosome): each owner finds the tandem edges of its pairs, the edge lists are all-gathered and every rank runs the
// same small union-find.  The caller (jcvi_b200/sharded.py) moves the 16-byte candidates GPU to GPU (all-to-all) and the
// edge lists (all-gather); these entry points are the device work in between.
//   jx_cscore_begin -> [all-reduce max] -> jx_shard_candidates -> [hist all-reduce, owners] -> jx_shard_partition
//   -> [all-to-all] -> jx_shard_pairs -> [all-gather edges] -> jx_shard_survivors -> jx_shard_scan ...
// Candidate = {qi, si, float32 score bits, (global order << 1) | (evalue < 1e-10)}; global order = ord_base + record
// index, monotone in file order across the ranks' slices, < 2^31.
// =====================================================================================================================
namespace {
__global__ void k_cp_hist(const uint4* __restrict__ C, uint32_t n, const int32_t* __restrict__ qchr, const int32_t* __restrict__ schr,
                          uint32_t nschr, uint32_t* hist) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint4 c = C[j];
    atomicAdd(hist + (uint32_t)qchr[c.x] * nschr + (uint32_t)schr[c.y], 1u);
  }
}
__global__ void k_dest_scatter(const uint4* __restrict__ C, uint32_t n, const int32_t* __restrict__ qchr, const int32_t* __restrict__ schr,
                               uint32_t nschr, const int32_t* __restrict__ owner, uint32_t* cursor, uint4* out) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint4 c = C[j];
    const int32_t d = owner[(uint32_t)qchr[c.x] * nschr + (uint32_t)schr[c.y]];
    out[atomicAdd(cursor + d, 1u)] = c;
  }
}
}  // namespace

extern "C" {

int jx_shard_reset(jx_ctx* ctx) {
  if (!ctx) return JX_EARG;
  cudaSetDevice(ctx->device);
  for (void* p : ctx->sh.allocs) cudaFreeAsync(p, ctx->stream);
  ctx->sh = jx_ctx::Shard();
  return JX_OK;
}

/* record slots of the slice jx_cscore_begin retained (this rank's share of the global order range) */
int jx_shard_slots(jx_ctx* ctx, int64_t* n_slots) {
  if (!ctx || !n_slots) return JX_EARG;
  if (!ctx->cache.valid) { ctx->err = "jx_shard_slots without retained records"; return JX_EARG; }
  *n_slots = ctx->cache.n_lines;
  return JX_OK;
}

int jx_shard_candidates(jx_ctx* ctx, double cscore, uint32_t ord_base, uint64_t* cand_dev, int64_t* n_out, uint32_t* cp_hist,
                        int64_t n_cp) {
  if (!ctx || !cand_dev || !n_out || !cp_hist) return JX_EARG;
  if (!ctx->sess_open || !ctx->cache.valid) { ctx->err = "jx_shard_candidates without jx_cscore_begin"; return JX_EARG; }
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  if (n_cp != (int64_t)Q.n_chr * S.n_chr) { ctx->err = "n_cp must be n_chr(q) * n_chr(s)"; return JX_EARG; }
  auto& c = ctx->cache;
  const uint32_t nrec = c.n_lines;
  if ((uint64_t)ord_base + nrec >= (1ull << 31)) { ctx->err = "more than 2^31 record slots over all ranks"; return JX_EUNSUPPORTED; }
  const bool use_cscore = cscore != 0.0;
  const size_t capC = (size_t)c.stats[2] + 1;
  JX_SALLOC(C, uint4, capC, false);
  JX_ALLOC(d_cnt, uint32_t, 2, true);
  JX_ALLOC(d_hist, uint32_t, (size_t)n_cp + 1, true);
  if (nrec) {
    unsigned grid = std::min<unsigned>(nblk((nrec + 3) / 4), 148u * JX_SEL_GRID);
    JX_ALLOC(bestq, uint32_t, (size_t)Q.G + 1, false);
    JX_ALLOC(bests, uint32_t, (size_t)S.G + 1, false);
    if (use_cscore) {
      JX_LAUNCH(k_gene_best, nblk((size_t)Q.G), 256, 0, ctx->sess_best, Q.name_id, (uint32_t)Q.G, bestq);
      JX_LAUNCH(k_gene_best, nblk((size_t)S.G), 256, 0, ctx->sess_best, S.name_id, (uint32_t)S.G, bests);
    }
    JX_LAUNCH(k_select, grid, 256, 0, c.recs, nrec, bestq, bests, cscore, use_cscore ? 1 : 0, ord_base, C, d_cnt, d_cnt + 1);
  }
  uint32_t h[2] = {0, 0};
  JX_CK(cudaMemcpyAsync(h, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  if (h[1]) { ctx->err = "float division by zero"; return JX_EZERODIV; }
  if (h[0]) JX_LAUNCH(k_cp_hist, std::min<unsigned>(nblk(h[0]), 148u * 8u), 256, 0, C, h[0], Q.chr_lex, S.chr_lex, (uint32_t)S.n_chr, d_hist);
  JX_CK(cudaMemcpyAsync(cp_hist, d_hist, (size_t)n_cp * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  ctx->sess_open = false;
  c.valid = c.plain;          // plain records stay for the liftover stage on the same held slice
  *cand_dev = (uint64_t)(uintptr_t)C; *n_out = h[0];
  return JX_OK;
}

int jx_shard_partition(jx_ctx* ctx, uint64_t cand_dev, int64_t n, const int32_t* owner, int64_t n_cp, int world,
                       uint64_t* out_dev, int64_t* counts, const uint32_t* cp_hist_in, uint32_t* cp_hist_out) {
  if (!ctx || !owner || !out_dev || !counts || world < 1 || n < 0 || n >= (1ll << 31)) return JX_EARG;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  if (n_cp != (int64_t)Q.n_chr * S.n_chr) { ctx->err = "n_cp must be n_chr(q) * n_chr(s)"; return JX_EARG; }
  // per-destination counts from the chromosome-pair histogram (host: n_cp is small)
  JX_ALLOC(d_hist, uint32_t, (size_t)n_cp + 1, true);
  const uint4* C = (const uint4*)(uintptr_t)cand_dev;
  const unsigned grid = std::max(1u, std::min<unsigned>(nblk((size_t)std::max<int64_t>(n, 1)), 148u * 8u));
  std::vector<uint32_t> hist((size_t)n_cp);
  if (cp_hist_in) memcpy(hist.data(), cp_hist_in, (size_t)n_cp * 4);        // (what jx_shard_candidates returned)
  else {
    if (n) JX_LAUNCH(k_cp_hist, grid, 256, 0, C, (uint32_t)n, Q.chr_lex, S.chr_lex, (uint32_t)S.n_chr, d_hist);
    JX_CK(cudaMemcpyAsync(hist.data(), d_hist, (size_t)n_cp * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
  }
  if (cp_hist_out) memcpy(cp_hist_out, hist.data(), (size_t)n_cp * 4);
  std::vector<uint32_t> cursor((size_t)world, 0);
  for (int64_t k = 0; k < n_cp; ++k) {
    if (owner[k] < 0 || owner[k] >= world) { ctx->err = "owner out of range"; return JX_EARG; }
    cursor[(size_t)owner[k]] += hist[(size_t)k];
  }
  uint32_t run = 0;
  for (int d = 0; d < world; ++d) { counts[d] = cursor[(size_t)d]; const uint32_t t = cursor[(size_t)d]; cursor[(size_t)d] = run; run += t; }
  JX_SALLOC(out, uint4, (size_t)n + 1, false);
  JX_ALLOC(d_owner, int32_t, (size_t)n_cp + 1, false);
  JX_ALLOC(d_cursor, uint32_t, (size_t)world, false);
  JX_CK(cudaMemcpyAsync(d_owner, owner, (size_t)n_cp * 4, cudaMemcpyHostToDevice, ctx->stream));
  JX_CK(cudaMemcpyAsync(d_cursor, cursor.data(), (size_t)world * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n) JX_LAUNCH(k_dest_scatter, grid, 256, 0, C, (uint32_t)n, Q.chr_lex, S.chr_lex, (uint32_t)S.n_chr, d_owner, d_cursor, out);
  JX_CK(cudaStreamSynchronize(ctx->stream));
  *out_dev = (uint64_t)(uintptr_t)out;
  return JX_OK;
}

// owner side, first half: per-pair best record of the received candidates + the tandem edges of this rank's pairs
int jx_shard_pairs(jx_ctx* ctx, uint64_t recv_dev, int64_t n, int tandem_nmax, uint64_t* qedges_dev, int64_t* nq,
                   uint64_t* sedges_dev, int64_t* ns) {
  if (!ctx || !qedges_dev || !nq || !sedges_dev || !ns || n < 0 || n >= (1ll << 31)) return JX_EARG;
  cudaSetDevice(ctx->device);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  auto& sh = ctx->sh;
  const uint32_t nC = (uint32_t)n;
  sh.C = (const uint4*)(uintptr_t)recv_dev; sh.nC = nC;
  uint32_t hcap = 1024;
  while (hcap < 2u * nC + 16u && hcap < (1u << 31)) hcap <<= 1;
  sh.hmask = hcap - 1u;
  JX_SALLOC(cnt, uint32_t, 8, true);
  JX_SALLOC(K1, unsigned long long, hcap, false);
  JX_SALLOC(V1, unsigned long long, hcap, true);
  JX_SALLOC(W1, uint32_t, hcap, false);
  JX_SALLOC(Cslot, uint32_t, (size_t)nC + 1, false);
  JX_SALLOC(win, uint8_t, (size_t)nC + 1, false);
  JX_SALLOC(qe, uint2, (size_t)nC + 1, false);
  JX_SALLOC(se, uint2, (size_t)nC + 1, false);
  sh.cnt = cnt; sh.K1 = K1; sh.V1 = V1; sh.W1 = W1; sh.Cslot = Cslot; sh.win = win;
  JX_CK(cudaMemcpyAsync(cnt, &nC, 4, cudaMemcpyHostToDevice, ctx->stream));
  JX_CK(cudaMemsetAsync(K1, 0xFF, (size_t)hcap * 8, ctx->stream));
  const unsigned g1 = std::max(1u, std::min<unsigned>(nblk(std::max<uint32_t>(nC, 1u)), 148u * 8u));
  JX_LAUNCH(k_pair_insert, g1, 256, 0, sh.C, cnt, (const uint8_t*)nullptr, K1, V1, sh.hmask, Cslot);
  JX_LAUNCH(k_pair_winner, g1, 256, 0, sh.C, cnt, (const uint8_t*)nullptr, Cslot, V1, W1, win, cnt + 5);
  if (tandem_nmax)
    JX_LAUNCH(k_tandem_edges, g1, 256, 0, sh.C, cnt, win, K1, W1, sh.hmask, tandem_nmax, (uint32_t)Q.G, (uint32_t)S.G, Q.chr_lex, S.chr_lex,
              qe, se, cnt + 2);
  uint32_t h[4] = {0, 0, 0, 0};
  JX_CK(cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  *qedges_dev = (uint64_t)(uintptr_t)qe; *nq = h[2];
  *sedges_dev = (uint64_t)(uintptr_t)se; *ns = h[3];
  return JX_OK;
}

// owner side, second half: ALL ranks' edges (device arrays of (a, b) rank pairs) -> the same families and mothers on
// every rank -> mother substitution, second per-pair best -> this rank's survivors (candidates with substituted ranks)
int jx_shard_survivors(jx_ctx* ctx, uint64_t qedges_dev, int64_t nq, uint64_t sedges_dev, int64_t ns, int tandem_nmax,
                       uint64_t* surv_dev, int64_t* n_surv, int32_t* q_mother, int32_t* s_mother) {
  if (!ctx || !surv_dev || !n_surv || nq < 0 || ns < 0) return JX_EARG;
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  auto& sh = ctx->sh;
  if (!sh.cnt) { ctx->err = "jx_shard_survivors without jx_shard_pairs"; return JX_EARG; }
  int rc;
  const uint32_t nC = sh.nC;
  const unsigned g1 = std::max(1u, std::min<unsigned>(nblk(std::max<uint32_t>(nC, 1u)), 148u * 8u));
  const uint4* D = sh.C; const uint8_t* fin = sh.win;
  if (tandem_nmax) {
    const uint32_t Gq = (uint32_t)Q.G, Gs = (uint32_t)S.G;
    JX_SALLOC(mq0, int32_t, (size_t)Gq + 1, false);
    sh.mq = mq0;
    if (ctx->is_self) sh.ms = sh.mq;
    else { JX_SALLOC(ms0, int32_t, (size_t)Gs + 1, false); sh.ms = ms0; }
    JX_ALLOC(parent_q, uint32_t, (size_t)Gq + 1, false);
    uint32_t* parent_s = parent_q;
    JX_LAUNCH(k_iota, nblk(std::max(Gq, 1u)), 256, 0, parent_q, Gq);
    if (!ctx->is_self) {
      JX_ALLOC(ps, uint32_t, (size_t)Gs + 1, false);
      parent_s = ps;
      JX_LAUNCH(k_iota, nblk(std::max(Gs, 1u)), 256, 0, parent_s, Gs);
    }
    JX_ALLOC(d_ne, uint32_t, 2, false);
    const uint32_t hne[2] = {(uint32_t)nq, (uint32_t)ns};
    JX_CK(cudaMemcpyAsync(d_ne, hne, 8, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned gq = std::max(1u, std::min<unsigned>(nblk((size_t)std::max<int64_t>(nq, 1)), 148u * 8u));
    const unsigned gs = std::max(1u, std::min<unsigned>(nblk((size_t)std::max<int64_t>(ns, 1)), 148u * 8u));
    if (nq) JX_LAUNCH(k_apply_edges, gq, 256, 0, (const uint2*)(uintptr_t)qedges_dev, d_ne, parent_q);
    if (ns) JX_LAUNCH(k_apply_edges, gs, 256, 0, (const uint2*)(uintptr_t)sedges_dev, d_ne + 1, parent_s);
    auto mothers = [&](uint32_t* parent, const SideDev& Dv, int32_t* m) -> int {
      const uint32_t G = (uint32_t)Dv.G;
      JX_ALLOC(root, uint32_t, (size_t)G + 1, false);
      JX_ALLOC(bk, unsigned long long, (size_t)G + 1, true);
      JX_ALLOC(mroot, uint32_t, (size_t)G + 1, false);
      if (G) {
        JX_LAUNCH(k_flatten, nblk(G), 256, 0, parent, G, root);
        JX_LAUNCH(k_mother1, nblk(G), 256, 0, root, Dv.length, Dv.lexrank, G, bk);
        JX_LAUNCH(k_mother2, nblk(G), 256, 0, root, Dv.length, Dv.lexrank, G, bk, mroot);
        JX_LAUNCH(k_mother3, nblk(G), 256, 0, root, mroot, G, m);
      }
      return JX_OK;
    };
    if ((rc = mothers(parent_q, Q, sh.mq))) return rc;
    if (!ctx->is_self && (rc = mothers(parent_s, S, sh.ms))) return rc;
    const uint32_t hcap = sh.hmask + 1u;
    JX_ALLOC(K2, unsigned long long, hcap, false);
    JX_ALLOC(V2, unsigned long long, hcap, true);
    JX_SALLOC(D2, uint4, (size_t)nC + 1, false);
    JX_ALLOC(Dslot, uint32_t, (size_t)nC + 1, false);
    JX_ALLOC(live, uint8_t, (size_t)nC + 1, false);
    JX_SALLOC(fin2, uint8_t, (size_t)nC + 1, false);
    JX_CK(cudaMemsetAsync(K2, 0xFF, (size_t)hcap * 8, ctx->stream));
    JX_LAUNCH(k_collapse_insert, g1, 256, 0, sh.C, sh.cnt, sh.win, sh.mq, sh.ms, ctx->q_same_s, K2, V2, sh.hmask, D2, Dslot, live);
    JX_LAUNCH(k_pair_winner, g1, 256, 0, D2, sh.cnt, live, Dslot, V2, (uint32_t*)nullptr, fin2, (uint32_t*)nullptr);
    D = D2; fin = fin2;
    if (q_mother && Gq) JX_CK(cudaMemcpyAsync(q_mother, sh.mq, (size_t)Gq * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (s_mother && Gs) JX_CK(cudaMemcpyAsync(s_mother, sh.ms, (size_t)Gs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  JX_SALLOC(F, uint4, (size_t)nC + 1, false);
  JX_ALLOC(okeys, unsigned long long, (size_t)nC + 1, false);
  JX_ALLOC(ovals, uint32_t, (size_t)nC + 1, false);
  JX_LAUNCH(k_final_emit, g1, 256, 0, D, sh.cnt, fin, F, okeys, ovals, sh.cnt + 4);
  uint32_t h5[5] = {0, 0, 0, 0, 0};
  JX_CK(cudaMemcpyAsync(h5, sh.cnt, 20, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  *surv_dev = (uint64_t)(uintptr_t)F; *n_surv = h5[4];
  return JX_OK;
}

// `.filtered` order of a set of survivors (any rank's, or the concatenation of all): score descending, order ascending
// -> host columns; line_off receives the global order of each survivor (the byte offset lives on its source rank)
int jx_shard_order(jx_ctx* ctx, uint64_t surv_dev, int64_t n, jx_filter_result* out) {
  if (!ctx || !out || n < 0 || n >= (1ll << 31)) return JX_EARG;
  memset(out, 0, sizeof *out);
  cudaSetDevice(ctx->device);
  Arena arena(ctx);
  const uint32_t nF = (uint32_t)n;
  out->n = nF;
  out->q_rank = (int32_t*)malloc(sizeof(int32_t) * (nF + 1));
  out->s_rank = (int32_t*)malloc(sizeof(int32_t) * (nF + 1));
  out->score = (float*)malloc(sizeof(float) * (nF + 1));
  out->line_off = (uint64_t*)malloc(sizeof(uint64_t) * (nF + 1));
  if (!nF) return JX_OK;
  const uint4* F = (const uint4*)(uintptr_t)surv_dev;
  JX_ALLOC(cnt, uint32_t, 2, true);
  JX_ALLOC(all, uint8_t, nF, false);
  JX_CK(cudaMemsetAsync(all, 1, nF, ctx->stream));
  JX_CK(cudaMemcpyAsync(cnt, &nF, 4, cudaMemcpyHostToDevice, ctx->stream));
  JX_ALLOC(F2, uint4, nF, false);
  JX_ALLOC(okeys, uint64_t, nF, false);
  JX_ALLOC(ovals, uint32_t, nF, false);
  JX_LAUNCH(k_final_emit, std::min<unsigned>(nblk(nF), 148u * 8u), 256, 0, F, cnt, all, F2, (unsigned long long*)okeys, ovals, cnt + 1);
  uint64_t* ok = okeys; uint32_t* ov = ovals;
  int rc;
  if ((rc = jx_sort_pairs(ctx, arena, ok, ov, nF, 0, 64))) return rc;
  std::vector<uint4> h(nF); std::vector<uint32_t> ord(nF);
  JX_CK(cudaMemcpyAsync(h.data(), F2, (size_t)nF * 16, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaMemcpyAsync(ord.data(), ov, (size_t)nF * 4, cudaMemcpyDeviceToHost, ctx->stream));
  JX_CK(cudaStreamSynchronize(ctx->stream));
  for (uint32_t t = 0; t < nF; ++t) {
    const uint4 c = h[ord[t]];
    out->q_rank[t] = (int32_t)c.x; out->s_rank[t] = (int32_t)c.y;
    memcpy(&out->score[t], &c.z, 4);
    out->line_off[t] = c.w >> 1;
  }
  return JX_OK;
}

}  // extern "C"

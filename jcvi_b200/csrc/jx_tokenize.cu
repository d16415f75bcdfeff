// K1 + K2: BLAST tab12 text -> one 16-byte record per line (gene ranks, float32 score, flags).
//
// Replaces the reference's per-line Python loop:
//   Blast.__iter__            src/jcvi/formats/blast.py:90-95   (skip '#', one BlastLine per row)
//   BlastLine.__init__        src/jcvi/formats/cblast.pyx:106-120 (sscanf of 12 fields)
//   gene_name                 src/jcvi/utils/cbook.py:308-329
//   qorder[query]/sorder[..]  src/jcvi/compara/blastfilter.py:56-84, synteny.py:261-283, 339-357
//   filter_cscore pass 1      src/jcvi/compara/blastfilter.py:227-232 (optional, fused)
//
// Layout.  The text is cut into 8 KiB tiles; one CTA of 128 threads stages a tile (+1 KiB of
// look-ahead) in shared memory with coalesced 16-byte loads and classifies every byte ONCE,
// branch-free, into two bit-per-byte masks (W: byte <= 0x20, N: neither digit nor tab; see
// jx_fast5.cuh) from which the terminator mask follows.  The lines that START in the tile are
// numbered with a block scan, and the tile's line count is published to a decoupled look-back
// chain so that every line learns its global ordinal in the same pass (no separate
// newline-count pass over the text).  Lines are then parsed one per thread from the masks and
// the staged words -- field boundaries are bit positions, validation is mask arithmetic, numbers
// are converted from words -- and the record is written at recs[ordinal]: 16-byte stores,
// consecutive threads -> consecutive addresses.  Irregular lines take the generic
// sscanf-faithful scanner.  HBM traffic per line: the line's bytes once + 16 bytes of record.
#include "jx_internal.cuh"
#include "jx_fast5.cuh"
#include "jx_fast6.cuh"

namespace {

struct TokArgs {
  const uint8_t* text;
  int64_t n;                 // total bytes of text
  uint64_t* desc;            // decoupled look-back descriptors, one per tile of the whole text
  uint32_t* ticket;          // ticket counter of this launch: tile = tile_begin + ticket
  uint32_t tile_begin, tile_end;   // tiles of this launch
  uint4* recs;
  uint32_t rec_cap;
  uint4* hashes;             // names mode only
  uint32_t seed;
  double f_score, f_pctid, f_evalue; long long f_hitlen; int f_noself, f_inverse, f_ids;   // filter mode only
  const uint4* qtab; const uint16_t* qdisp; uint32_t qbshift, qcshift; const uint32_t* qmeta; const uint32_t* qblob;   // 2 x uint4 per slot
  const uint4* stab; const uint16_t* sdisp; uint32_t sbshift, scshift; const uint32_t* smeta; const uint32_t* sblob;
  const uint32_t* qovh; const uint32_t* qovr; uint32_t qnov;   // overflow lists (equal 32-bit hashes), usually empty
  const uint32_t* sovh; const uint32_t* sovr; uint32_t snov;
  const int32_t* qchr; const int32_t* schr;   // chr_lex per gene (near-diagonal rule)
  const uint32_t* qnid; const uint32_t* snid;
  uint32_t* best;
  const uint64_t* excl; int64_t n_excl;
  int strip, is_self, neardiag;
  int merge_q;             // the lines come grouped by query (lastal / BLAST order): merge a warp's equal-query reductions
  unsigned long long* stats;
};

#define FLAG_AGG (1ull << 62)
#define FLAG_INC (2ull << 62)
#define VAL_MASK ((1ull << 62) - 1)

#ifndef JX_LCAP
#define JX_LCAP 1024
#endif
constexpr int LCAP = JX_LCAP;     // line starts of a tile kept in shared memory (more -> search path)

struct WinReader {             // byte view: shared-memory window, global memory beyond it
  const uint8_t* sm;
  const uint8_t* g;
  int64_t base, n;
  int win = JX_WIN;            // bytes of the window
  __device__ __forceinline__ uint32_t operator()(int p) const {
    if (p < win) return sm[p];
    int64_t gp = base + p;
    return gp < n ? g[gp] : (uint32_t)'\n';
  }
};
struct SmemWords {             // aligned 32-bit words of the window
  const uint32_t* w;
  __device__ __forceinline__ uint32_t operator()(int i) const { return w[i]; }
};

// generic (byte reader) probe: any name length < 128, verified against the word blob
struct SideTab {
  const uint4* tab; const uint16_t* disp; uint32_t bshift, cshift;
  const uint32_t* meta; const uint32_t* blob;
  const uint32_t* ovh; const uint32_t* ovr; uint32_t nov;
};
template <class R>
__device__ bool name_equals(R& rd, int a, int b, uint32_t r, const uint32_t* meta, const uint32_t* blob) {
  const uint32_t m = __ldg(meta + r), len = m & 127u;
  if (len != (uint32_t)(b - a)) return false;
  const uint32_t* bw = blob + (m >> 7);
  for (uint32_t i = 0; i < len; ++i)
    if (((__ldg(bw + (i >> 2)) >> (8 * (i & 3))) & 0xFFu) != rd(a + (int)i)) return false;
  return true;
}
__device__ __forceinline__ bool ovf_has(const uint32_t* ovh, uint32_t nov, uint32_t h) {   // is h in the sorted list?
  uint32_t lo = 0, hi = nov;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(ovh + mid) < h) lo = mid + 1; else hi = mid; }
  return lo < nov && __ldg(ovh + lo) == h;
}
template <class R>
__device__ int table_lookup_bytes(R& rd, int a, int b, const SideTab& T) {
  if (b <= a || b - a >= 128) return -1;
  const uint32_t h = jx::slot_hash_bytes(rd, a, b);
  const uint32_t idx = jx::slot_index(h, __ldg(T.disp + (h >> T.bshift)), T.cshift);
  const uint32_t r1 = __ldg(T.tab + 2 * idx).x;
  if (r1 && name_equals(rd, a, b, r1 - 1u, T.meta, T.blob)) return (int)(r1 - 1u);
  if (T.nov) {                                    // names sharing the 32-bit hash of a table entry
    uint32_t lo = 0, hi = T.nov;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(T.ovh + mid) < h) lo = mid + 1; else hi = mid; }
    for (; lo < T.nov && __ldg(T.ovh + lo) == h; ++lo) {
      const uint32_t r = __ldg(T.ovr + lo);
      if (name_equals(rd, a, b, r, T.meta, T.blob)) return (int)r;
    }
  }
  return -1;
}

struct SmemText {              // byte and unaligned-word views of the staged window
  const uint8_t* sm;
  const uint32_t* w;
  __device__ __forceinline__ uint32_t operator()(int p) const { return sm[p]; }
  __device__ __forceinline__ uint32_t load32(int p) const {   // bytes [p, p+4)
    const int i = p >> 2;
    return __funnelshift_r(w[i], w[i + 1], (uint32_t)(p & 3) * 8u);
  }
};
struct MaskWords {             // one of the bit-per-byte class masks of the window
  const uint32_t* m;
  __device__ __forceinline__ uint32_t operator()(int i) const { return m[i]; }
};

#ifndef JX_TOK_MINBLK
#define JX_TOK_MINBLK 8
#endif
constexpr int STAGE_BYTES = JX_WIN + 32;          // one staged window (+ slack for unaligned word reads)

// ---- TMA (bulk async copy) + mbarrier, PTX for sm_90+/sm_100a -------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0u;
}

// Persistent CTAs: each takes tiles from the launch's ticket counter and keeps the NEXT tile's bytes
// in flight (one bulk copy issued by one thread into the other half of a double buffer) while it
// works on the current one, so the HBM latency of staging is off the critical path and costs no
// load/store instructions.
// MODE 0: ranks through the BED name tables (the path's kernel).  MODE 1 (NAMES): no table, the
// record keeps only the line position and `hashes` receives a 64-bit hash of both names
// (jx_distinct_names, formats.blast cscore); a separate instantiation, so the main one is unchanged.
// MODE 2: formats.blast filter -- every line is kept or dropped by a predicate on its own fields.
template <int MODE>
__global__ void __launch_bounds__(JX_TOK_THREADS, JX_TOK_MINBLK) k_tokenize(const TokArgs a) {
  constexpr bool NAMES = MODE == 1;
  constexpr bool FILTER = MODE == 2;
  constexpr int NGRP = JX_TOK_THREADS / JX_GROUP;          // independent tile groups per CTA
  constexpr int GWARPS = JX_GROUP / 32;
  __shared__ __align__(128) uint8_t g_text[NGRP][2][STAGE_BYTES];
  __shared__ __align__(8) uint64_t g_bar[NGRP][2];
  // (+2 words: a 64-bit mask window that starts in the last words reads past the window; those bits are masked off)
  __shared__ uint32_t g_term[NGRP][JX_WIN / 32 + 2];      // terminator bit per byte of the window
  __shared__ uint32_t g_w[NGRP][JX_WIN / 32 + 2];         // byte <= 0x20
  __shared__ uint32_t g_n[NGRP][JX_WIN / 32 + 2];         // byte is neither a digit nor a tab
  __shared__ uint32_t g_pref[NGRP][JX_GROUP];         // exclusive prefix of line starts per 64-byte segment
  __shared__ unsigned long long g_start[NGRP][JX_GROUP];
  __shared__ uint16_t g_lpos[NGRP][LCAP];             // start offset of every line of the tile
  __shared__ uint32_t g_warp[NGRP][GWARPS];
  __shared__ uint32_t g_next[NGRP][2];
  __shared__ unsigned long long g_base[NGRP];

  const int grp = threadIdx.x / JX_GROUP;
  const int tid = threadIdx.x % JX_GROUP, lane = tid & 31, warp = tid >> 5;   // within the group
  uint8_t (*s_text)[STAGE_BYTES] = g_text[grp];
  uint64_t* s_bar = g_bar[grp];
  uint32_t* s_term = g_term[grp]; uint32_t* s_w = g_w[grp]; uint32_t* s_n = g_n[grp];
  uint32_t* s_pref = g_pref[grp]; unsigned long long* s_start = g_start[grp];
  uint16_t* s_lpos = g_lpos[grp]; uint32_t* s_warp = g_warp[grp]; uint32_t* s_next = g_next[grp];
  unsigned long long& s_base = g_base[grp];
  // group-wide synchronisation (one group per CTA: a CTA barrier)
  auto gsync = [&]() { __syncthreads(); };
  auto gany = [&](bool p) -> bool {
    return __syncthreads_or(p) != 0;
  };
  uint32_t c_comment = 0, c_mapped = 0, c_unmapped = 0, c_hard = 0, c_long = 0, c_excl = 0, c_self = 0, c_lines = 0;
  unsigned long long first_bad = 0;
  bool overflow = false;

  // bytes of tile t that exist in the text, capped at one stage; the bulk copy moves the 16-byte multiple
  auto avail_of = [&](uint32_t t) -> uint32_t {
    const int64_t left = a.n - (int64_t)t * JX_TILE;
    return (uint32_t)(left < (int64_t)STAGE_BYTES ? (left < 0 ? 0 : left) : STAGE_BYTES);
  };
  auto issue = [&](uint32_t t, int stage) {      // one thread
    const uint32_t bytes = avail_of(t) & ~15u;
    if (bytes) {
      mbar_expect_tx(&s_bar[stage], bytes);
      bulk_g2s(s_text[stage], a.text + (int64_t)t * JX_TILE, bytes, &s_bar[stage]);
    }
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the CTA's first two tiles: both copies in flight at once
    const uint32_t t0 = a.tile_begin + atomicAdd(a.ticket, 1u);
    s_next[0] = t0;
    if (t0 < a.tile_end) issue(t0, 0);
    const uint32_t t1 = a.tile_begin + atomicAdd(a.ticket, 1u);
    s_next[1] = t1;
    if (t1 < a.tile_end) issue(t1, 1);
  }
  gsync();

  // ---- prepare(tile): wait for its bytes, classify them, count and locate its line starts, PUBLISH the count ----
  // Runs one tile AHEAD of the parse (software pipeline, see the loop below): the tile's aggregate is in the
  // look-back chain before this CTA blocks in the look-back of the tile it is parsing, so no CTA's
  // publication ever waits behind another CTA's look-back.
  auto prepare = [&](const uint32_t tile, const int stage, const uint32_t parity, uint32_t& nlines_out, bool& has_bar_out) {
  const uint32_t* smw = reinterpret_cast<const uint32_t*>(s_text[stage]);
  const uint8_t* sm = s_text[stage];
  const int64_t base = (int64_t)tile * JX_TILE;
  // ---- the window: wait for the bulk copy, '\n' beyond the end of the text ---------------------------
  {
    const uint32_t avail = avail_of(tile), bulk = avail & ~15u;
    if (bulk) {
      while (!mbar_try_wait(&s_bar[stage], parity)) {}
    }
    if (avail < (uint32_t)STAGE_BYTES) {          // last tiles of the text only
      uint8_t* wsm = s_text[stage];
      for (uint32_t i = bulk + tid; i < (uint32_t)STAGE_BYTES; i += JX_GROUP)
        wsm[i] = i < avail ? a.text[base + i] : (uint8_t)'\n';
      gsync();
    }
  }

  // ---- byte classes (SWAR, 32 bytes per mask word) and the terminator mask -----------------------------
  uint32_t hiacc = 0;
  for (int seg = tid; seg < JX_WIN / 32; seg += JX_GROUP) {
    const uint4 v0 = reinterpret_cast<const uint4*>(smw)[2 * seg], v1 = reinterpret_cast<const uint4*>(smw)[2 * seg + 1];
    const uint32_t w8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    uint32_t wm, nm;
    jx::classify32(w8, wm, nm, hiacc);
    s_w[seg] = wm;
    s_n[seg] = nm;
    // a terminator is a byte <= 0x20 that is not a tab: usually the only such bytes, checked by value
    uint32_t c = wm & nm, term = 0;
    while (c) {
      const int b = __ffs((int)c) - 1;
      c &= c - 1u;
      const uint32_t ch = sm[seg * 32 + b];
      if (ch == '\n' || ch == '\r') term |= 1u << b;
    }
    s_term[seg] = term;
  }
  has_bar_out = gany((hiacc & 0x80808080u) != 0u);   // may the tile hold a '|' ?  (also: masks visible)

  // ---- line starts inside the tile (64 bytes per thread) + block scan ----------------------------------
  unsigned long long starts = 0;
  if (tid * JX_SEG < JX_TILE) {
    const unsigned long long term = (unsigned long long)s_term[2 * tid] | ((unsigned long long)s_term[2 * tid + 1] << 32);
    uint32_t prevbit;
    if (tid == 0) prevbit = (base == 0) ? 1u : (uint32_t)jx::is_term(a.text[base - 1]);
    else prevbit = s_term[2 * tid - 1] >> 31;
    starts = (term << 1) | prevbit;
    const int64_t remain = a.n - (base + (int64_t)tid * JX_SEG);   // only positions that exist can start a line
    if (remain <= 0) starts = 0;
    else if (remain < 64) starts &= (1ull << remain) - 1ull;
  }
  const uint32_t cnt = __popcll(starts);
  uint32_t inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
  uint32_t wbase = 0, nlines = 0;
  {
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < GWARPS; ++w) { const uint32_t v = s_warp[w]; if (w < warp) wbase += v; nlines += v; }
  }
  const uint32_t mypref = wbase + inc - cnt;
  s_pref[tid] = mypref;
  s_start[tid] = starts;
  {
    unsigned long long m = starts;
    uint32_t o = mypref;
    while (m && o < LCAP) {
      s_lpos[o++] = (uint16_t)(tid * JX_SEG + __ffsll((long long)m) - 1);
      m &= m - 1;
    }
  }
  // the tile's own count goes into the look-back chain at once
  if (tid == 0) {
    const uint64_t d = (tile == 0 ? FLAG_INC : FLAG_AGG) | (uint64_t)nlines;
    atomicExch((unsigned long long*)(a.desc + tile), (unsigned long long)d);
  }
  nlines_out = nlines;
  };

  // ---- decoupled look-back: global ordinal of the tile's first line ---------------------------------
  // The walk over the predecessors' descriptors is done by warp 0 AFTER the tile's lines are parsed AND the
  // next tile's count is published, when every thread holds at most one line (the normal case).  Tiles with
  // more lines than threads resolve their ordinal first.
  auto look_back = [&](const uint32_t tile, const uint32_t nlines) {      // all 32 lanes of warp 0
    // LBW windows of 32 descriptors are fetched per round (independent loads, one memory latency).
    // Measured on B200: 1 window 0.0728 ms / 64 MiB, 2: 0.0730, 4: 0.0749, 8: 0.0802 -- wider rounds only
    // add polling traffic, so one window it is.
#ifndef JX_LBW
#define JX_LBW 1
#endif
    constexpr int LBW = JX_LBW;
    unsigned long long excl = 0;
    if (tile > 0) {
      int64_t j = (int64_t)tile - 1;
      uint32_t spins = 0;
      bool done = false;
      while (!done) {
        uint64_t d[LBW];
#pragma unroll
        for (int w = 0; w < LBW; ++w) {
          const int64_t idx = j - 32 * w - lane;
          d[w] = idx >= 0 ? *((volatile uint64_t*)(a.desc + idx)) : FLAG_INC;
        }
        bool stalled = false;
#pragma unroll
        for (int w = 0; w < LBW; ++w) {
          if (!done && !stalled) {
            const uint32_t flag = (uint32_t)(d[w] >> 62);
            const unsigned notready = __ballot_sync(0xFFFFFFFFu, flag == 0);
            const unsigned isinc = __ballot_sync(0xFFFFFFFFu, flag == 2);
            const int first = isinc ? __ffs(isinc) - 1 : 32;
            const unsigned upto = first < 31 ? ((2u << first) - 1u) : 0xFFFFFFFFu;
            if (notready & upto) {
              stalled = true;               // a needed predecessor has not published yet: poll again from here
            } else {
              // aggregates are per-tile line counts (<= 8 Ki): their sum fits 32 bits; the one
              // inclusive prefix (lane `first`) is 62 bits wide and is fetched on its own
              const uint32_t agg = ((int)lane < first) ? (uint32_t)d[w] : 0u;
              excl += __reduce_add_sync(0xFFFFFFFFu, agg);
              if (first < 32) {
                excl += __shfl_sync(0xFFFFFFFFu, (unsigned long long)(d[w] & VAL_MASK), first);
                done = true;
              } else {
                j -= 32;
              }
            }
          }
        }
        if (stalled) {
          if (++spins > (1u << 22)) { if (lane == 0) atomicExch(a.stats + 10, 1ull); done = true; }   // never hang the GPU
          else __nanosleep(64);
        }
      }
      if (lane == 0) atomicExch((unsigned long long*)(a.desc + tile), (unsigned long long)(FLAG_INC | (excl + nlines)));
    }
    if (lane == 0) s_base = excl;
  };
  // ---- the pipeline: iteration `it` parses tile T[it] (prepared one iteration earlier), then prepares T[it+1],
  // then resolves T[it]'s ordinal and stores its records.  Tile T[j] lives in stage j & 1; its ticket is
  // taken (and its copy started) right after the parse of T[j-2] has released that stage.
  uint32_t tile = a.tile_end, nlines = 0;
  bool has_bar = false;
  for (int it = -1;; ++it) {
  const bool have_cur = it >= 0;                   // uniform; then tile < a.tile_end
  const int stage = it & 1;
  const uint32_t* smw = reinterpret_cast<const uint32_t*>(s_text[stage]);
  const uint8_t* sm = s_text[stage];
  const int64_t base = (int64_t)tile * JX_TILE;
  const bool defer = nlines <= JX_GROUP;           // uniform over the group
  uint4 held = make_uint4(JX_NOREC, 0u, 0u, 0u);
  bool have_held = false;
  uint4 hheld = make_uint4(0u, 0u, 0u, 0u);
  if (have_cur) {
  if (!defer) {
    gsync();                                       // nobody still reads the previous tile's s_base
    if (warp == 0) look_back(tile, nlines);
    gsync();
  }

  // ---- parse: one line per thread -----------------------------------------------------------------
  WinReader rd{sm, a.text, base, a.n};
  const SmemWords wr{smw};
  const MaskWords mwr{s_w}, mnr{s_n};
  const SmemText B{sm, smw};
  const SideTab TQ{a.qtab, a.qdisp, a.qbshift, a.qcshift, a.qmeta, a.qblob, a.qovh, a.qovr, a.qnov};
  const SideTab TS{a.stab, a.sdisp, a.sbshift, a.scshift, a.smeta, a.sblob, a.sovh, a.sovr, a.snov};
  for (uint32_t k0 = 0; k0 < nlines; k0 += JX_GROUP) {
    const uint32_t k = k0 + tid;
    if (k >= nlines) break;
    int s;
    if (k < LCAP) s = s_lpos[k];
    else {   // overfull tile (average line < 8 bytes): locate the line through the prefix array
      int lo = 0, hi = JX_GROUP - 1;
      while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (s_pref[mid] <= k) lo = mid; else hi = mid - 1; }
      unsigned long long m = s_start[lo];
      for (uint32_t r = k - s_pref[lo]; r; --r) m &= m - 1;
      s = lo * JX_SEG + __ffsll((long long)m) - 1;
    }
    int e;   // end of line: next terminator at or after s
    if (k + 1 < nlines && k + 1 < LCAP) e = (int)s_lpos[k + 1] - 1;
    else {
      int w = s >> 5;
      uint32_t m = s_term[w] & (0xFFFFFFFFu << (s & 31));
      while (!m && ++w < JX_WIN / 32) m = s_term[w];
      if (m) e = w * 32 + __ffs(m) - 1;
      else { e = JX_WIN; while (base + e < a.n && !jx::is_term(a.text[base + e])) ++e; }
    }
    uint4 rec = make_uint4(JX_NOREC, 0u, 0u, (uint32_t)s << 1);
    uint4 hrec = make_uint4(0u, 0u, 0u, 0u);
    if (s < e && sm[s] == '#') {
      ++c_comment;
    } else if (FILTER) {
      // ---- formats.blast filter: BlastLine(row) then the predicate of blast.py:681-699 ------------------
      const uint32_t prevc = s > 0 ? sm[s - 1] : (base > 0 ? a.text[base - 1] : 0u);
      const bool crlf_gap = s >= e && prevc == '\r' && sm[s] == '\n';     // not a row (universal newlines)
      if (!crlf_gap) {
        jx::FilterFields ff;
        jx::filter_parse(B, mwr, mnr, rd, s, e, e + 8 <= JX_WIN, a.f_evalue, ff);
        if (ff.longname) ++c_long;
        if (ff.hard || ff.longname) { ++c_hard; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; }
        else {
          bool noids = false;
          if (a.f_ids) {                                  // c.query in ids and c.subject in ids (raw names)
            const bool inq = table_lookup_bytes(rd, ff.qa, ff.qb, TQ) >= 0;
            const bool ins = inq && table_lookup_bytes(rd, ff.sa, ff.sb, TS) >= 0;
            noids = !(inq && ins);
          }
          const bool remove = jx::filter_remove(rd, ff, a.f_score, a.f_pctid, a.f_hitlen, noids, a.f_inverse != 0, a.f_noself != 0);
          if (!remove) {
            // filter-mode record: {1, length of row.rstrip(), 64-bit offset of the row in the text}
            int r = e;                                   // str.rstrip(): ASCII white space and 0x1c-0x1f
            while (r > s) { const uint32_t ch = rd(r - 1); if (jx::is_space(ch) || (ch - 0x1Cu) <= 3u) --r; else break; }
            const unsigned long long o = (unsigned long long)(base + s);
            rec = make_uint4(1u, (uint32_t)(r - s), (uint32_t)o, (uint32_t)(o >> 32));
            ++c_mapped;
          } else ++c_unmapped;
        }
      }
    } else if (s < e) {
      jx::LineTok t;
      const bool fast = (e + 8 <= JX_WIN) && jx::fast_line5(B, mwr, mnr, s, e, t);
      if (!fast) jx::tokenize_line(rd, s, e, t);
      if (NAMES) {
        if (t.nconv >= 2 && !t.longname && !t.hard) {
          int qa = t.qa, qb = t.qb, sa = t.sa, sb = t.sb;
          uint64_t hq, hs;
          if (fast) {
            if (a.strip) {
              int qbar = -1, sbar = -1;
              if (has_bar) { qbar = jx::find_bar(B, qa, qb); sbar = jx::find_bar(B, sa, sb); }
              qb = jx::strip5(B, qa, qb, qbar);
              sb = jx::strip5(B, sa, sb, sbar);
            }
          } else if (a.strip) { jx::strip_isoform(rd, qa, qb); jx::strip_isoform(rd, sa, sb); }
          const int lq = qb - qa, lsn = sb - sa;
          if (fast && lq > 0 && lsn > 0 && lq < 4 * JX_SLOT_WORDS && lsn < 4 * JX_SLOT_WORDS) {
            uint32_t wq[JX_SLOT_WORDS], ws[JX_SLOT_WORDS];
            jx::name_words6(wr, qa, lq, wq);
            jx::name_words6(wr, sa, lsn, ws);
            hq = jx::name64_words6(wq, lq, a.seed);
            hs = jx::name64_words6(ws, lsn, a.seed);
          } else {
            hq = lq > 0 ? jx::name64_bytes(rd, qa, qb, a.seed) : 0ull;
            hs = lsn > 0 ? jx::name64_bytes(rd, sa, sb, a.seed) : 0ull;
          }
          hrec = make_uint4((uint32_t)hq, (uint32_t)(hq >> 32), (uint32_t)hs, (uint32_t)(hs >> 32));
          ++c_mapped;
        } else {
          if (t.longname) { ++c_long; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; }
          else if (t.hard) { ++c_hard; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; }
          else ++c_unmapped;            // fewer than two fields
        }
      } else {
      bool bad = false, selfhit = false;
      int qi = -1, si = -1;
      uint32_t nidq = 0, nids = 0;
      bool have_nid = false;
      if (t.nconv >= 2) {
        if (t.longname) { bad = true; ++c_long; }
        int qa = t.qa, qb = t.qb, sa = t.sa, sb = t.sb;
        if (a.is_self && (qb - qa) == (sb - sa)) {      // `is_self and query == subject` on raw names
          if (fast) {
            uint32_t diff = 0;
            for (int i = 0; i < qb - qa; i += 4) {
              uint32_t d = B.load32(qa + i) ^ B.load32(sa + i);
              const int rem = qb - qa - i;
              if (rem < 4) d &= (1u << (8 * rem)) - 1u;
              diff |= d;
            }
            selfhit = diff == 0u;
          } else {
            selfhit = true;
            for (int i = 0; i < qb - qa; ++i) if (rd(qa + i) != rd(sa + i)) { selfhit = false; break; }
          }
        }
        if (!selfhit && !bad) {
          if (fast) {
            if (a.strip) {
              int qbar = -1, sbar = -1;
              if (has_bar) { qbar = jx::find_bar(B, qa, qb); sbar = jx::find_bar(B, sa, sb); }
              qb = jx::strip5(B, qa, qb, qbar);
              sb = jx::strip5(B, sa, sb, sbar);
            }
            const int lq = qb - qa, lsn = sb - sa;
            if (lq > 0 && lsn > 0 && lq < 4 * JX_SLOT_WORDS && lsn < 4 * JX_SLOT_WORDS) {
              uint32_t wq[JX_SLOT_WORDS], ws[JX_SLOT_WORDS];
              jx::name_words6(wr, qa, lq, wq);
              jx::name_words6(wr, sa, lsn, ws);
              // perfect hash: one displacement load + one 32-byte slot per name, no chain to walk
              const uint32_t hq = jx::slot_hash6(wq, lq), hs = jx::slot_hash6(ws, lsn);
              const uint32_t dq = __ldg(a.qdisp + (hq >> a.qbshift)), ds = __ldg(a.sdisp + (hs >> a.sbshift));
              const uint4* pq = a.qtab + 2 * jx::slot_index(hq, dq, a.qcshift);
              const uint4* ps = a.stab + 2 * jx::slot_index(hs, ds, a.scshift);
              const uint4 q0 = __ldg(pq), q1 = __ldg(pq + 1), s0 = __ldg(ps), s1 = __ldg(ps + 1);
              const uint32_t dfq = (q0.z ^ wq[0]) | (q0.w ^ wq[1]) | (q1.x ^ wq[2]) | (q1.y ^ wq[3]) | (q1.z ^ wq[4]) | (q1.w ^ wq[5]);
              const uint32_t dfs = (s0.z ^ ws[0]) | (s0.w ^ ws[1]) | (s1.x ^ ws[2]) | (s1.y ^ ws[3]) | (s1.z ^ ws[4]) | (s1.w ^ ws[5]);
              qi = (q0.x && !dfq) ? (int)(q0.x - 1u) : -1;
              si = (s0.x && !dfs) ? (int)(s0.x - 1u) : -1;
              nidq = q0.y; nids = s0.y;
              have_nid = true;
              if ((a.qnov | a.snov) && (qi < 0 || si < 0)) {   // a miss may be a name parked in an overflow list
                const bool maybe = (qi < 0 && ovf_has(a.qovh, a.qnov, hq)) || (si < 0 && ovf_has(a.sovh, a.snov, hs));
                if (maybe) {
                  have_nid = false;
                  qi = table_lookup_bytes(rd, qa, qb, TQ);
                  si = qi >= 0 ? table_lookup_bytes(rd, sa, sb, TS) : -1;
                }
              }
            } else {                                   // long (>= 24 bytes) or empty stripped name
              qi = table_lookup_bytes(rd, qa, qb, TQ);
              if (qi >= 0) si = table_lookup_bytes(rd, sa, sb, TS);
            }
          } else {
            if (a.strip) { jx::strip_isoform(rd, qa, qb); jx::strip_isoform(rd, sa, sb); }
            qi = table_lookup_bytes(rd, qa, qb, TQ);
            if (qi >= 0) si = table_lookup_bytes(rd, sa, sb, TS);
          }
        }
      }
      if (selfhit) ++c_self;
      else if (bad) { if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; }
      else if (qi < 0 || si < 0) ++c_unmapped;
      else if (t.hard) { ++c_hard; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; }
      else {
        if (a.is_self && qi > si) { int tmp = qi; qi = si; si = tmp; }
        bool keep = true;
        if (a.neardiag && a.qchr[qi] == a.schr[si] && si - qi < 40) { keep = false; ++c_self; }
        if (keep && a.n_excl) {
          uint64_t key = ((uint64_t)(uint32_t)qi << 32) | (uint32_t)si;
          int64_t l = 0, h = a.n_excl;
          while (l < h) { int64_t m = (l + h) >> 1; if (a.excl[m] < key) l = m + 1; else h = m; }
          if (l < a.n_excl && a.excl[l] == key) { keep = false; ++c_excl; }
        }
        if (keep) {
          ++c_mapped;
          const uint32_t sbits = __float_as_uint(t.score);
          rec = make_uint4((uint32_t)qi, (uint32_t)si, sbits, ((uint32_t)s << 1) | (uint32_t)t.eflag);
          if (a.best && t.score > 0.f) {            // best_score[name] = max(score), starting at 0.0
            if (!have_nid) { nidq = a.qnid[qi]; nids = a.snid[si]; }   // is_self may have swapped qi/si: both updated
            uint32_t* bq = a.best + nidq;
            uint32_t* bs = a.best + nids;
            const uint32_t vq = __ldcg(bq), vs = __ldcg(bs);     // both reads in flight together
            if (vq < sbits) atomicMax(bq, sbits);
            if (vs < sbits) atomicMax(bs, sbits);
          }
        }
      }
      }  // !NAMES
    } else {
      // an empty segment between the '\r' and the '\n' of a CR LF pair is not a line (universal newlines);
      // a real blank line is a BlastLine with empty names, never in the BED
      const uint32_t prev = s > 0 ? sm[s - 1] : (base > 0 ? a.text[base - 1] : 0u);
      if (!(prev == '\r' && sm[s] == '\n')) ++c_unmapped;
    }
    if (defer) { held = rec; have_held = true; if (NAMES) hheld = hrec; }
    else {
      const unsigned long long ord = s_base + k;
      if (ord < a.rec_cap) { a.recs[ord] = rec; if (NAMES) a.hashes[ord] = hrec; }
    }
  }
  gsync();                                         // this stage and the mask / line-start arrays are free again
  if (tid == 0) {                                  // T[it + 2] goes into the stage just released
    const uint32_t t = a.tile_begin + atomicAdd(a.ticket, 1u);
    s_next[stage] = t;
    if (t < a.tile_end) issue(t, stage);
  }
  }  // have_cur: parse

  // ---- the next tile: classified, counted and published before this CTA waits for anybody ----------------
  const uint32_t nxt = s_next[(it + 1) & 1];
  uint32_t nxt_lines = 0;
  bool nxt_bar = false;
  if (nxt < a.tile_end) prepare(nxt, (it + 1) & 1, (uint32_t)((it + 1) >> 1) & 1u, nxt_lines, nxt_bar);

  if (have_cur && defer) {
    if (warp == 0) look_back(tile, nlines);
  }
  gsync();                                         // s_base; the next tile's arrays; s_next
  if (have_cur) {
    if (defer) {
      const unsigned long long ord = s_base + tid;
      if (have_held && ord < a.rec_cap) { a.recs[ord] = held; if (NAMES) a.hashes[ord] = hheld; }
    }
    if (tid == 0) {
      c_lines += nlines;
      if (s_base + nlines > a.rec_cap) overflow = true;
    }
  }
  if (nxt >= a.tile_end) break;
  tile = nxt; nlines = nxt_lines; has_bar = nxt_bar;
  }  // tiles

  // ---- counters ---------------------------------------------------------------------------------------
  uint32_t vals[7] = {c_comment, c_mapped, c_unmapped, c_hard, c_long, c_excl, c_self};
#pragma unroll
  for (int i = 0; i < 7; ++i) vals[i] = __reduce_add_sync(0xFFFFFFFFu, vals[i]);
  if (lane == 0) {
    if (vals[0]) atomicAdd(a.stats + 1, (unsigned long long)vals[0]);
    if (vals[1]) atomicAdd(a.stats + 2, (unsigned long long)vals[1]);
    if (vals[2]) atomicAdd(a.stats + 3, (unsigned long long)vals[2]);
    if (vals[3]) atomicAdd(a.stats + 4, (unsigned long long)vals[3]);
    if (vals[4]) atomicAdd(a.stats + 5, (unsigned long long)vals[4]);
    if (vals[5]) atomicAdd(a.stats + 8, (unsigned long long)vals[5]);
    if (vals[6]) atomicAdd(a.stats + 9, (unsigned long long)vals[6]);
  }
  if (first_bad) atomicMin(a.stats + 7, first_bad);
  if (tid == 0) {
    if (c_lines) atomicAdd(a.stats + 0, (unsigned long long)c_lines);
    if (overflow) atomicExch(a.stats + 6, 1ull);
  }
}


// =====================================================================================================================
// Version 6 of the path's kernel (MODE 0).  Same machinery as above -- persistent CTAs, bulk-copy double buffer,
// software-pipelined tiles, decoupled look-back -- with the instruction stream cut for the ALU pipe, which is
// what bounds this kernel:
//   * window = 128 x 64 bytes: every thread classifies exactly one 64-byte segment, straight-line;
//   * classification assumes 7-bit text (no per-byte masking, 7 instead of 11 logic ops per word); a tile with a
//     byte >= 0x80 is re-classified with the exact classifier (CTA-uniform branch);
//   * the '|' test moved from every byte of the tile to the name words of each line;
//   * per line: jx_fast6.cuh (32-bit mask windows, evalue decided instead of converted, byte-permute digit groups,
//     3-word names, multiply-add name hash); every line it refuses goes through `slow_line`, the version-5 route
//     followed by the generic sscanf-faithful scanner, kept out of line so that it costs the fast path no registers.
// Results are identical to the version-5 kernel by construction: the fast route accepts a subset of lines and is
// checked against the generic scanner and glibc on the host (tests/test_parse_host.py).
enum { WH_MAPPED = 0, WH_COMMENT, WH_UNMAPPED, WH_HARD, WH_LONG, WH_EXCL, WH_SELF, WH_NONE };
struct SlowOut { uint4 rec; int what; };

// 64 bytes of 7-bit text -> W / N mask words (two each); `any` accumulates the bytes (bit 7 = not 7-bit text)
__device__ __forceinline__ void classify64_ascii(const uint4& v0, const uint4& v1, const uint4& v2, const uint4& v3, uint32_t* wm,
                                                 uint32_t* nm, uint32_t& any) {
  const uint32_t x[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t pw[4], pn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a = x[8 * h + 2 * k], b = x[8 * h + 2 * k + 1];
      any |= a | b;
      // bit 7 of each byte: tw clear <=> byte <= 0x20; un set <=> neither digit nor tab (no carries between 7-bit bytes)
      const uint32_t twa = a + 0x5F5F5F5Fu, twb = b + 0x5F5F5F5Fu;
      const uint32_t una = ((a ^ 0x30303030u) + 0x76767676u) & ((a ^ 0x09090909u) + 0x7F7F7F7Fu);
      const uint32_t unb = ((b ^ 0x30303030u) + 0x76767676u) & ((b ^ 0x09090909u) + 0x7F7F7F7Fu);
      // eight flags -> eight mask bits (text order) in the top byte
      const uint32_t gw = ((~twa >> 4) & 0x08080808u) | (~twb & 0x80808080u);
      const uint32_t gn = ((una >> 4) & 0x08080808u) | (unb & 0x80808080u);
      pw[k] = gw * 0x00204081u;
      pn[k] = gn * 0x00204081u;
    }
    wm[h] = jx::top_bytes(pw[0], pw[1], pw[2], pw[3]);
    nm[h] = jx::top_bytes(pn[0], pn[1], pn[2], pn[3]);
  }
}

// one 256-bit load of a 32-byte table slot (LDG.E.256 on sm_100: half the LSU work of two LDG.128)
__device__ __forceinline__ void ldg256(const uint4* p, uint4& lo, uint4& hi) {       // p: 32-byte aligned, read-only data
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

// does the word hold a '|' (0x7C) byte?  (exact "any zero byte" test on w ^ 0x7C7C7C7C)
__device__ __forceinline__ uint32_t bar_in(uint32_t w) {
  const uint32_t v = w ^ 0x7C7C7C7Cu;
  return (v - 0x01010101u) & ~v & 0x80808080u;
}

// The version-5 route for ONE line of a staged tile: fast_line5 on the masks, else the generic scanner; names through
// the byte-reader table lookup when they are long.  Mirrors the per-line body of k_tokenize<0> above.
__device__ __noinline__ SlowOut slow_line(const TokArgs& a, const uint8_t* sm, const uint32_t* s_w, const uint32_t* s_n,
                                          const int64_t base, const int s, const int e, const int win) {
  SlowOut o;
  o.rec = make_uint4(JX_NOREC, 0u, 0u, (uint32_t)s << 1);
  o.what = WH_NONE;
  const uint32_t* smw = reinterpret_cast<const uint32_t*>(sm);
  if (s >= e) {
    // an empty segment between the '\r' and the '\n' of a CR LF pair is not a line (universal newlines);
    // a real blank line is a BlastLine with empty names, never in the BED
    const uint32_t prev = s > 0 ? sm[s - 1] : (base > 0 ? a.text[base - 1] : 0u);
    if (!(prev == '\r' && sm[s] == '\n')) o.what = WH_UNMAPPED;
    return o;
  }
  WinReader rd{sm, a.text, base, a.n, win};
  const SmemWords wr{smw};
  const MaskWords mwr{s_w}, mnr{s_n};
  const SmemText B{sm, smw};
  const SideTab TQ{a.qtab, a.qdisp, a.qbshift, a.qcshift, a.qmeta, a.qblob, a.qovh, a.qovr, a.qnov};
  const SideTab TS{a.stab, a.sdisp, a.sbshift, a.scshift, a.smeta, a.sblob, a.sovh, a.sovr, a.snov};
  jx::LineTok t;
  const bool fast = (e + 8 <= win) && jx::fast_line5(B, mwr, mnr, s, e, t);
  if (!fast) jx::tokenize_line(rd, s, e, t);
  bool bad = false, selfhit = false;
  int qi = -1, si = -1;
  if (t.nconv >= 2) {
    if (t.longname) bad = true;
    int qa = t.qa, qb = t.qb, sa = t.sa, sb = t.sb;
    if (a.is_self && (qb - qa) == (sb - sa)) {      // `is_self and query == subject` on raw names
      selfhit = true;
      for (int i = 0; i < qb - qa; ++i) if (rd(qa + i) != rd(sa + i)) { selfhit = false; break; }
    }
    if (!selfhit && !bad) {
      if (a.strip) { jx::strip_isoform(rd, qa, qb); jx::strip_isoform(rd, sa, sb); }
      qi = table_lookup_bytes(rd, qa, qb, TQ);
      if (qi >= 0) si = table_lookup_bytes(rd, sa, sb, TS);
    }
  }
  if (bad) { o.what = WH_LONG; return o; }
  if (selfhit) { o.what = WH_SELF; return o; }
  if (qi < 0 || si < 0) { o.what = WH_UNMAPPED; return o; }
  if (t.hard) { o.what = WH_HARD; return o; }
  if (a.is_self && qi > si) { int tmp = qi; qi = si; si = tmp; }
  if (a.neardiag && a.qchr[qi] == a.schr[si] && si - qi < 40) { o.what = WH_SELF; return o; }
  o.rec = make_uint4((uint32_t)qi, (uint32_t)si, __float_as_uint(t.score), ((uint32_t)s << 1) | (uint32_t)t.eflag);
  o.what = WH_MAPPED;             // (the exclusion list and the best-score table are the caller's business)
  return o;
}

__global__ void __launch_bounds__(JX_TOK_THREADS, JX_TOK_MINBLK) k_tokenize6(const __grid_constant__ TokArgs a) {
  __shared__ __align__(128) uint8_t s_text[2][STAGE_BYTES];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_term[JX_WIN / 32 + 2];      // terminator bit per byte of the window (+2: 64-bit windows read past it)
  __shared__ uint32_t s_w[JX_WIN / 32 + 2];         // byte <= 0x20
  __shared__ uint32_t s_n[JX_WIN / 32 + 2];         // byte is neither a digit nor a tab
  __shared__ uint32_t s_pref[JX_TOK_THREADS];       // exclusive prefix of line starts per 64-byte segment
  __shared__ unsigned long long s_start[JX_TOK_THREADS];
  __shared__ uint16_t s_lpos[LCAP];                 // start offset of every line of the tile
  __shared__ uint32_t s_warp[JX_TOK_THREADS / 32], s_wbit[JX_TOK_THREADS / 32];
  __shared__ uint32_t s_next[2];
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t c_comment = 0, c_mapped = 0, c_unmapped = 0, c_hard = 0, c_long = 0, c_excl = 0, c_self = 0, c_lines = 0;
  unsigned long long first_bad = 0;
  bool overflow = false;
  // tiles whose whole stage lies inside the text need no length arithmetic
  const uint32_t n_full = a.n >= (int64_t)STAGE_BYTES ? (uint32_t)((a.n - STAGE_BYTES) / JX_TILE) + 1u : 0u;

  auto avail_of = [&](uint32_t t) -> uint32_t {    // bytes of tile t's stage that exist in the text
    if (t < n_full) return (uint32_t)STAGE_BYTES;
    const int64_t left = a.n - (int64_t)t * JX_TILE;
    return (uint32_t)(left < (int64_t)STAGE_BYTES ? (left < 0 ? 0 : left) : STAGE_BYTES);
  };
  auto issue = [&](uint32_t t, int stage) {        // one thread
    const uint32_t bytes = avail_of(t) & ~15u;
    if (bytes) {
      mbar_expect_tx(&s_bar[stage], bytes);
      bulk_g2s(s_text[stage], a.text + (int64_t)t * JX_TILE, bytes, &s_bar[stage]);
    }
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t t0 = a.tile_begin + atomicAdd(a.ticket, 1u);
    s_next[0] = t0;
    if (t0 < a.tile_end) issue(t0, 0);
    const uint32_t t1 = a.tile_begin + atomicAdd(a.ticket, 1u);
    s_next[1] = t1;
    if (t1 < a.tile_end) issue(t1, 1);
  }
  __syncthreads();

  // ---- prepare(tile): wait for its bytes, classify them, count and locate its line starts, PUBLISH the count ----
  auto prepare = [&](const uint32_t tile, const int stage, const uint32_t parity, uint32_t& nlines_out) {
    const uint8_t* sm = s_text[stage];
    const int64_t base = (int64_t)tile * JX_TILE;
    const bool full = tile < n_full;
    {
      const uint32_t avail = avail_of(tile), bulk = avail & ~15u;
      if (bulk) {
        while (!mbar_try_wait(&s_bar[stage], parity)) {}
      }
      if (!full) {                                  // last tiles of the text only: '\n' beyond its end
        uint8_t* wsm = s_text[stage];
        for (uint32_t i = bulk + tid; i < (uint32_t)STAGE_BYTES; i += JX_TOK_THREADS)
          wsm[i] = i < avail ? a.text[base + i] : (uint8_t)'\n';
        __syncthreads();
      }
    }
    // ---- byte classes of this thread's 64-byte segment ------------------------------------------------------------
    const uint4* seg4 = reinterpret_cast<const uint4*>(sm + tid * JX_SEG);
    const uint4 v0 = seg4[0], v1 = seg4[1], v2 = seg4[2], v3 = seg4[3];
    uint32_t wm[2], nm[2], term[2], any = 0;
    classify64_ascii(v0, v1, v2, v3, wm, nm, any);
    auto finish_masks = [&]() {
      // a terminator is a byte <= 0x20 that is not a tab: usually the only such bytes, checked by value
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t c = wm[h] & nm[h], tm = 0;
        while (c) {
          const int b = __ffs((int)c) - 1;
          c &= c - 1u;
          const uint32_t ch = sm[tid * JX_SEG + h * 32 + b];
          if (ch == '\n' || ch == '\r') tm |= 1u << b;
        }
        term[h] = tm;
      }
      reinterpret_cast<uint2*>(s_w)[tid] = make_uint2(wm[0], wm[1]);
      reinterpret_cast<uint2*>(s_n)[tid] = make_uint2(nm[0], nm[1]);
      reinterpret_cast<uint2*>(s_term)[tid] = make_uint2(term[0], term[1]);
      if (lane == 31) s_wbit[warp] = term[1] >> 31;     // the bit before the next warp's first segment
    };
    finish_masks();
    if (__syncthreads_or((any & 0x80808080u) != 0u)) {   // a byte >= 0x80 somewhere in the window: exact classifier
      uint32_t hiacc = 0;
      const uint32_t w8a[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w}, w8b[8] = {v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
      jx::classify32(w8a, wm[0], nm[0], hiacc);
      jx::classify32(w8b, wm[1], nm[1], hiacc);
      finish_masks();
      __syncthreads();
    }
    // ---- line starts inside the tile + block scan ---------------------------------------------------------------------
    // the bit before this segment: the previous thread's last terminator bit (the byte before the tile for thread 0)
    uint32_t prevbit = __shfl_up_sync(0xFFFFFFFFu, term[1] >> 31, 1);
    if (lane == 0) prevbit = tid ? s_wbit[warp - 1] : ((base == 0) ? 1u : (uint32_t)jx::is_term(a.text[base - 1]));
    unsigned long long starts = 0;
    if (tid < JX_TILE / JX_SEG) {
      const unsigned long long tm = (unsigned long long)term[0] | ((unsigned long long)term[1] << 32);
      starts = (tm << 1) | prevbit;
      if (!full) {                                  // only positions that exist can start a line
        const int64_t remain = a.n - (base + (int64_t)tid * JX_SEG);
        if (remain <= 0) starts = 0;
        else if (remain < 64) starts &= (1ull << remain) - 1ull;
      }
    }
    const uint32_t slo = (uint32_t)starts, shi = (uint32_t)(starts >> 32);
    const uint32_t cnt = __popc(slo) + __popc(shi);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, nlines = 0;
#pragma unroll
    for (int w = 0; w < JX_TOK_THREADS / 32; ++w) { const uint32_t v = s_warp[w]; if (w < warp) wbase += v; nlines += v; }
    uint32_t o = wbase + inc - cnt;
    if (nlines > (uint32_t)LCAP) { s_pref[tid] = o; s_start[tid] = starts; }   // overfull tile: the search path needs them
    {
      uint32_t m = slo;
      const uint32_t p0 = (uint32_t)(tid * JX_SEG);
      while (m && o < (uint32_t)LCAP) { s_lpos[o++] = (uint16_t)(p0 + __ffs((int)m) - 1); m &= m - 1u; }
      m = shi;
      while (m && o < (uint32_t)LCAP) { s_lpos[o++] = (uint16_t)(p0 + 32u + __ffs((int)m) - 1); m &= m - 1u; }
    }
    if (tid == 0) {
      const uint64_t d = (tile == 0 ? FLAG_INC : FLAG_AGG) | (uint64_t)nlines;
      atomicExch((unsigned long long*)(a.desc + tile), (unsigned long long)d);
    }
    nlines_out = nlines;
  };

  // ---- decoupled look-back (warp 0): global ordinal of the tile's first line ------------------------------------
  auto look_back = [&](const uint32_t tile, const uint32_t nlines) {
#ifdef JX_EXP_NOLB
    if (lane == 0) s_base = (unsigned long long)tile * 128ull;     // EXPERIMENT: sparse slots, no chain (timing only)
    return;
#endif
    unsigned long long excl = 0;
    if (tile > 0) {
      int64_t j = (int64_t)tile - 1;
      uint32_t spins = 0;
      bool done = false;
      while (!done) {
        const int64_t idx = j - lane;
        const uint64_t d = idx >= 0 ? *((volatile uint64_t*)(a.desc + idx)) : FLAG_INC;
        const uint32_t flag = (uint32_t)(d >> 62);
        const unsigned notready = __ballot_sync(0xFFFFFFFFu, flag == 0);
        const unsigned isinc = __ballot_sync(0xFFFFFFFFu, flag == 2);
        const int first = isinc ? __ffs(isinc) - 1 : 32;
        const unsigned upto = first < 31 ? ((2u << first) - 1u) : 0xFFFFFFFFu;
        if (notready & upto) {
          if (++spins > (1u << 22)) { if (lane == 0) atomicExch(a.stats + 10, 1ull); done = true; }   // never hang the GPU
          else __nanosleep(64);
        } else {
          const uint32_t agg = ((int)lane < first) ? (uint32_t)d : 0u;
          excl += __reduce_add_sync(0xFFFFFFFFu, agg);
          if (first < 32) {
            excl += __shfl_sync(0xFFFFFFFFu, (unsigned long long)(d & VAL_MASK), first);
            done = true;
          } else {
            j -= 32;
          }
        }
      }
      if (lane == 0) atomicExch((unsigned long long*)(a.desc + tile), (unsigned long long)(FLAG_INC | (excl + nlines)));
    }
    if (lane == 0) s_base = excl;
  };

  const bool v6 = !a.is_self;                        // self comparisons compare and swap raw names: version-5 route
  uint32_t tile = a.tile_end, nlines = 0;
  for (int it = -1;; ++it) {
    const bool have_cur = it >= 0;
    const int stage = it & 1;
    const uint8_t* sm = s_text[stage];
    const uint32_t* smw = reinterpret_cast<const uint32_t*>(sm);
    const int64_t base = (int64_t)tile * JX_TILE;
    const bool defer = nlines <= (uint32_t)JX_TOK_THREADS;
    uint4 held = make_uint4(JX_NOREC, 0u, 0u, 0u);
    bool have_held = false;
    if (have_cur) {
      if (!defer) {
        __syncthreads();                             // nobody still reads the previous tile's s_base
        if (warp == 0) look_back(tile, nlines);
        __syncthreads();
      }
      const SmemWords wr{smw};
      const MaskWords mwr{s_w}, mnr{s_n};
      const SmemText B{sm, smw};
      for (uint32_t k0 = 0; k0 < nlines; k0 += JX_TOK_THREADS) {
        const uint32_t k = k0 + tid;
        if (k >= nlines) break;
        int s, e;
        if (k < (uint32_t)LCAP) s = s_lpos[k];
        else {   // overfull tile (average line < 8 bytes): locate the line through the prefix array
          int lo = 0, hi = JX_TOK_THREADS - 1;
          while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pref[mid] <= k) lo = mid; else hi = mid - 1; }
          unsigned long long m = s_start[lo];
          for (uint32_t r = k - s_pref[lo]; r; --r) m &= m - 1;
          s = lo * JX_SEG + __ffsll((long long)m) - 1;
        }
        if (k + 1 < nlines && k + 1 < (uint32_t)LCAP) e = (int)s_lpos[k + 1] - 1;
        else {   // the tile's last line: next terminator at or after s
          int w = s >> 5;
          uint32_t m = s_term[w] & (0xFFFFFFFFu << (s & 31));
          while (!m && ++w < JX_WIN / 32) m = s_term[w];
          if (m) e = w * 32 + __ffs(m) - 1;
          else { e = JX_WIN; while (base + e < a.n && !jx::is_term(a.text[base + e])) ++e; }
        }
        uint4 rec = make_uint4(JX_NOREC, 0u, 0u, (uint32_t)s << 1);
        int what = WH_NONE;
        bool fastdone = false;
        const int len = e - s;
#ifdef JX_EXP_NOPARSE
        if (true) { what = WH_COMMENT; fastdone = true; rec.x = (uint32_t)len; } else
#endif
        if (len > 0 && sm[s] == '#') {
          what = WH_COMMENT; fastdone = true;
        } else if (v6 && e + 8 <= JX_WIN) {
          jx::Line6 L;
          if (jx::fast_line6_names(mwr, mnr, s, len, L)) {
            // ---- names: three zero-padded words each, gene_name, perfect-hash probe --------------------------------
            const int qa = s, sa = s + L.t0 + 1;
            int lq = L.t0, lsn = L.t1 - L.t0 - 1;
            const int iq = qa >> 2, is = sa >> 2;
            const uint32_t shq = (uint32_t)(qa & 3) * 8u, shs = (uint32_t)(sa & 3) * 8u;
            const uint32_t xq0 = smw[iq], xq1 = smw[iq + 1], xq2 = smw[iq + 2], xq3 = smw[iq + 3];
            const uint32_t xs0 = smw[is], xs1 = smw[is + 1], xs2 = smw[is + 2], xs3 = smw[is + 3];
            uint32_t q0 = __funnelshift_r(xq0, xq1, shq), q1 = __funnelshift_r(xq1, xq2, shq), q2 = __funnelshift_r(xq2, xq3, shq);
            uint32_t s0 = __funnelshift_r(xs0, xs1, shs), s1 = __funnelshift_r(xs1, xs2, shs), s2 = __funnelshift_r(xs2, xs3, shs);
            if (a.strip) { lq = jx::strip6(B, qa, lq, q0); lsn = jx::strip6(B, sa, lsn, s0); }
            if (lq <= 12 && lsn <= 12 && lq >= 1 && lsn >= 1) {
              q0 &= jx::name_mask(lq, 0); q1 &= jx::name_mask(lq, 1); q2 &= jx::name_mask(lq, 2);
              s0 &= jx::name_mask(lsn, 0); s1 &= jx::name_mask(lsn, 1); s2 &= jx::name_mask(lsn, 2);
              const uint32_t bar = a.strip ? (bar_in(q0) | bar_in(q1) | bar_in(q2) | bar_in(s0) | bar_in(s1) | bar_in(s2)) : 0u;
              if (bar == 0u) {
                uint32_t hq = jx::slot_hash_step(jx::slot_hash_init(), q0), hs = jx::slot_hash_step(jx::slot_hash_init(), s0);
                if (lq > 4) hq = jx::slot_hash_step(hq, q1);
                if (lsn > 4) hs = jx::slot_hash_step(hs, s1);
                if (lq > 8) hq = jx::slot_hash_step(hq, q2);
                if (lsn > 8) hs = jx::slot_hash_step(hs, s2);
                hq = jx::slot_hash_finish(hq, (uint32_t)lq); hs = jx::slot_hash_finish(hs, (uint32_t)lsn);
                // the table loads are issued HERE; the rest of the line is parsed while they are in flight
#ifdef JX_EXP_NOLOOKUP
                const uint4 tq0 = make_uint4((hq & 1023u) + 1u, hq & 1023u, q0, q1), tq1 = make_uint4(q2, 0u, 0u, 0u), ts0 = make_uint4((hs & 1023u) + 1u, hs & 1023u, s0, s1), ts1 = make_uint4(s2, 0u, 0u, 0u);
#else
                const uint32_t dq = __ldg(a.qdisp + (hq >> a.qbshift)), ds = __ldg(a.sdisp + (hs >> a.sbshift));
                const uint4* pq = a.qtab + 2 * jx::slot_index(hq, dq, a.qcshift);
                const uint4* ps = a.stab + 2 * jx::slot_index(hs, ds, a.scshift);
                uint4 tq0, tq1, ts0, ts1;                // one 256-bit load per 32-byte slot (LDG.E.256: half the LSU work of two LDG.128)
          ldg256(pq, tq0, tq1);
          ldg256(ps, ts0, ts1);
#endif
                if (jx::fast_line6_rest(B, wr, mwr, mnr, s, len, L)) {
                  const uint32_t dfq = (tq0.z ^ q0) | (tq0.w ^ q1) | (tq1.x ^ q2) | tq1.y | tq1.z | tq1.w;
                  const uint32_t dfs = (ts0.z ^ s0) | (ts0.w ^ s1) | (ts1.x ^ s2) | ts1.y | ts1.z | ts1.w;
                  const bool okq = tq0.x != 0u && dfq == 0u, oks = ts0.x != 0u && dfs == 0u;
                  // a miss may be a name parked in an overflow list (names with equal 32-bit hashes): rare route
                  if ((okq && oks) || (a.qnov | a.snov) == 0u) {
                    fastdone = true;
                    if (!(okq && oks)) what = WH_UNMAPPED;
                    else {
                      const uint32_t qi = tq0.x - 1u, si = ts0.x - 1u;
                      bool keep = true;
                      if (a.n_excl) {
                        const uint64_t key = ((uint64_t)qi << 32) | si;
                        int64_t l = 0, h = a.n_excl;
                        while (l < h) { const int64_t m = (l + h) >> 1; if (a.excl[m] < key) l = m + 1; else h = m; }
                        if (l < a.n_excl && a.excl[l] == key) { keep = false; what = WH_EXCL; }
                      }
                      if (keep) {
                        what = WH_MAPPED;
                        const uint32_t sbits = __float_as_uint(L.score);
                        rec = make_uint4(qi, si, sbits, ((uint32_t)s << 1) | (uint32_t)L.eflag);
#ifndef JX_EXP_NOLOOKUP
                        if (a.best && L.score > 0.f) {        // best_score[name] = max(score), starting at 0.0:
                          atomicMax(a.best + tq0.y, sbits);   // two reductions (no return value, nothing to wait for)
                          atomicMax(a.best + ts0.y, sbits);
                        }
#endif
                      }
                    }
                  }
                }
              }
            }
          }
        }
        if (!fastdone) {
          const SlowOut so = slow_line(a, sm, s_w, s_n, base, s, e, JX_WIN);
          rec = so.rec; what = so.what;
          if (what == WH_MAPPED) {
            if (a.n_excl) {
              const uint64_t key = ((uint64_t)rec.x << 32) | rec.y;
              int64_t l = 0, h = a.n_excl;
              while (l < h) { const int64_t m = (l + h) >> 1; if (a.excl[m] < key) l = m + 1; else h = m; }
              if (l < a.n_excl && a.excl[l] == key) { what = WH_EXCL; rec.x = JX_NOREC; }
            }
            if (what == WH_MAPPED && a.best && __uint_as_float(rec.z) > 0.f) {
              atomicMax(a.best + a.qnid[rec.x], rec.z);
              atomicMax(a.best + a.snid[rec.y], rec.z);
            }
          }
        }
        switch (what) {
          case WH_MAPPED: ++c_mapped; break;
          case WH_COMMENT: ++c_comment; break;
          case WH_UNMAPPED: ++c_unmapped; break;
          case WH_HARD: ++c_hard; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; break;
          case WH_LONG: ++c_long; if (!first_bad) first_bad = (unsigned long long)(base + s) + 1; break;
          case WH_EXCL: ++c_excl; break;
          case WH_SELF: ++c_self; break;
          default: break;
        }
        if (defer) { held = rec; have_held = true; }
        else {
          const unsigned long long ord = s_base + k;
          if (ord < a.rec_cap) a.recs[ord] = rec;
        }
      }
      __syncthreads();                               // this stage and the mask / line-start arrays are free again
      if (tid == 0) {                                // T[it + 2] goes into the stage just released
        const uint32_t t = a.tile_begin + atomicAdd(a.ticket, 1u);
        s_next[stage] = t;
        if (t < a.tile_end) issue(t, stage);
      }
    }

    // ---- the next tile: classified, counted and published before this CTA waits for anybody -------------------------
    const uint32_t nxt = s_next[(it + 1) & 1];
    uint32_t nxt_lines = 0;
    if (nxt < a.tile_end) prepare(nxt, (it + 1) & 1, (uint32_t)((it + 1) >> 1) & 1u, nxt_lines);

    if (have_cur && defer) {
      if (warp == 0) look_back(tile, nlines);
    }
    __syncthreads();                                 // s_base; the next tile's arrays; s_next
    if (have_cur) {
      if (defer) {
        const unsigned long long ord = s_base + tid;
        if (have_held && ord < a.rec_cap) a.recs[ord] = held;
      }
      if (tid == 0) {
        c_lines += nlines;
        if (s_base + nlines > a.rec_cap) overflow = true;
      }
    }
    if (nxt >= a.tile_end) break;
    tile = nxt; nlines = nxt_lines;
  }

  uint32_t vals[7] = {c_comment, c_mapped, c_unmapped, c_hard, c_long, c_excl, c_self};
#pragma unroll
  for (int i = 0; i < 7; ++i) vals[i] = __reduce_add_sync(0xFFFFFFFFu, vals[i]);
  if (lane == 0) {
    if (vals[0]) atomicAdd(a.stats + 1, (unsigned long long)vals[0]);
    if (vals[1]) atomicAdd(a.stats + 2, (unsigned long long)vals[1]);
    if (vals[2]) atomicAdd(a.stats + 3, (unsigned long long)vals[2]);
    if (vals[3]) atomicAdd(a.stats + 4, (unsigned long long)vals[3]);
    if (vals[4]) atomicAdd(a.stats + 5, (unsigned long long)vals[4]);
    if (vals[5]) atomicAdd(a.stats + 8, (unsigned long long)vals[5]);
    if (vals[6]) atomicAdd(a.stats + 9, (unsigned long long)vals[6]);
  }
  if (first_bad) atomicMin(a.stats + 7, first_bad);
  if (tid == 0) {
    if (c_lines) atomicAdd(a.stats + 0, (unsigned long long)c_lines);
    if (overflow) atomicExch(a.stats + 6, 1ull);
  }
}


// =====================================================================================================================
// Version 7 of the path's kernel: WARP-AUTONOMOUS sub-tiles, no CTA barrier and no cross-CTA chain.
//
// The version-6 profile (profiles/r2_k_tokenize6_*) showed the kernel bound by waiting, not by any pipe: 23 % of the
// warp samples sat in CTA barriers (13 % of them behind warp 0's decoupled look-back), 25 % on the table loads.  Both
// waits came from the tile organisation -- 128 threads share a tile, so every phase ends in a barrier, and records are
// stored at dense line ordinals, so every tile needs the line counts of all its predecessors.  Here
//   * one WARP owns a sub-tile: a 4 KiB window (32 lanes x 128 bytes) = 3840 bytes whose lines are the warp's + 256
//     bytes of look-ahead; masks, line starts and the per-line parse only need __syncwarp;
//   * records go to FIXED SLOTS: sub-tile t owns recs[t * cap, (t + 1) * cap) (cap >= its number of lines, the rest is
//     filled with "no record"), so no warp ever needs another warp's count.  A record's index is still monotone in file
//     order -- all the first-seen / stable-sort rules need -- and the line's byte offset follows from the index alone
//     (index / cap * 3840 + offset inside the sub-tile): the per-tile descriptor chain is gone.  cap starts at 64
//     (lines of >= 60 bytes on average) and grows by relaunch when a sub-tile overflows;
//   * the windows arrive by bulk copy into a ring of W7_NBUF buffers per CTA: a warp that finishes a sub-tile re-arms
//     its buffer for the CTA's position p + NBUF and takes the next position; there are always NBUF - 4 copies in flight.
// Per-line work is version 6's (jx_fast6.cuh first, `slow_line` for everything it refuses).
constexpr int W7_WIN = 4096;
constexpr int W7_TILE = 3840;                       // = 30 lanes x 128 bytes = 16 x 240
constexpr int W7_PRE = 16;                          // bytes staged in front of the window (the byte before it decides its first line start)
constexpr int W7_STAGE = W7_PRE + W7_WIN + 32;      // + slack for unaligned word reads past the window
#ifndef JX_W7_NBUF
#define JX_W7_NBUF 5
#endif
constexpr int W7_NBUF = JX_W7_NBUF;
constexpr int W7_LC = 126;                          // line starts of a window kept in shared memory (more: per-lane route)
#ifndef JX_W7_MINBLK
#define JX_W7_MINBLK 7     // 72 registers: 8 CTAs (64 registers) spill inside the per-line code and measure 7 % slower
#endif

// Per-warp counters live in shared memory: only lines that are NOT plain mapped records touch them (mapped = lines -
// everything else), so the common case costs neither registers nor instructions.
struct Counters7 {
  uint32_t* cnt;                 // [WH_NONE + 1] of this warp
  unsigned long long* stats;     // the launch's global statistics (first bad offset: atomicMin, rare)
};

// One line [s, e) of a staged window -> record (version 6's per-line body); false = the line needs `parse_slow7`.
// The slow route is NOT called from here: a call inside the hot loop makes the compiler keep the loop's state (line
// index, slot pointer, mask pointers) in local memory around it -- the reloads were the kernel's largest long-scoreboard
// stall (profiles/r2_v7b_*: 11 % of the samples).  The kernel lists the refused lines and handles them after the loop.
template <bool MERGE_Q>
__device__ __forceinline__ bool parse_fast7(const TokArgs& a, const uint8_t* sm, const uint32_t* s_w, const uint32_t* s_n,
                                            const int s, const int e, const int win, const bool v6, const Counters7& c,
                                            uint4& rec) {
  const uint32_t* smw = reinterpret_cast<const uint32_t*>(sm);
  rec = make_uint4(JX_NOREC, 0u, 0u, (uint32_t)s << 1);
  int what = WH_NONE;
  bool fastdone = false;
  const int len = e - s;
  // the first twelve bytes of the line (the query name of a regular line); also answer "is it a '#' line?" -- the text
  // bytes the per-line code needs come from these words, not from byte loads (shared-memory wavefronts are what the
  // kernel has least of: 32 lanes read 32 different lines, every load is a 2-3-way bank conflict)
  const int qa = s, iq = qa >> 2;
  const uint32_t shq = (uint32_t)(qa & 3) * 8u;
  const uint32_t xq0 = smw[iq], xq1 = smw[iq + 1], xq2 = smw[iq + 2], xq3 = smw[iq + 3];
  uint32_t q0 = __funnelshift_r(xq0, xq1, shq), q1 = __funnelshift_r(xq1, xq2, shq), q2 = __funnelshift_r(xq2, xq3, shq);
  if (len > 0 && (q0 & 0xFFu) == (uint32_t)'#') {
    what = WH_COMMENT; fastdone = true;
  } else if (v6 && e + 8 <= win) {
    const SmemWords wr{smw};
    const MaskWords mwr{s_w}, mnr{s_n};
    const SmemText B{sm, smw};
    jx::Line6 L;
    if (jx::fast_line6_names(mwr, mnr, s, len, L)) {
      const int sa = s + L.t0 + 1;
      int lq = L.t0, lsn = L.t1 - L.t0 - 1;
      const int is = sa >> 2;
      const uint32_t shs = (uint32_t)(sa & 3) * 8u;
      const uint32_t xs0 = smw[is], xs1 = smw[is + 1], xs2 = smw[is + 2], xs3 = smw[is + 3];
      uint32_t s0 = __funnelshift_r(xs0, xs1, shs), s1 = __funnelshift_r(xs1, xs2, shs), s2 = __funnelshift_r(xs2, xs3, shs);
      // (taking the names' last two bytes from q0..q2 / s0..s2 instead of two byte loads each was measured 3.6 % slower:
      //  the selects and shifts cost more issue slots than the loads cost wavefronts)
      if (a.strip) { lq = jx::strip6(B, qa, lq, q0); lsn = jx::strip6(B, sa, lsn, s0); }
      if (lq <= 12 && lsn <= 12 && lq >= 1 && lsn >= 1) {
        q0 &= jx::name_mask(lq, 0); q1 &= jx::name_mask(lq, 1); q2 &= jx::name_mask(lq, 2);
        s0 &= jx::name_mask(lsn, 0); s1 &= jx::name_mask(lsn, 1); s2 &= jx::name_mask(lsn, 2);
        const uint32_t bar = a.strip ? (bar_in(q0) | bar_in(q1) | bar_in(q2) | bar_in(s0) | bar_in(s1) | bar_in(s2)) : 0u;
        if (bar == 0u) {
          uint32_t hq = jx::slot_hash_step(jx::slot_hash_init(), q0), hs = jx::slot_hash_step(jx::slot_hash_init(), s0);
          if (lq > 4) hq = jx::slot_hash_step(hq, q1);
          if (lsn > 4) hs = jx::slot_hash_step(hs, s1);
          if (lq > 8) hq = jx::slot_hash_step(hq, q2);
          if (lsn > 8) hs = jx::slot_hash_step(hs, s2);
          hq = jx::slot_hash_finish(hq, (uint32_t)lq); hs = jx::slot_hash_finish(hs, (uint32_t)lsn);
          // the table loads are issued HERE; the rest of the line is parsed while they are in flight
          const uint32_t dq = __ldg(a.qdisp + (hq >> a.qbshift)), ds = __ldg(a.sdisp + (hs >> a.sbshift));
          const uint4* pq = a.qtab + 2 * jx::slot_index(hq, dq, a.qcshift);
          const uint4* ps = a.stab + 2 * jx::slot_index(hs, ds, a.scshift);
          uint4 tq0, tq1, ts0, ts1;                // one 256-bit load per 32-byte slot (LDG.E.256: half the LSU work of two LDG.128)
          ldg256(pq, tq0, tq1);
          ldg256(ps, ts0, ts1);
          if (jx::fast_line6_rest(B, wr, mwr, mnr, s, len, L)) {
            const uint32_t dfq = (tq0.z ^ q0) | (tq0.w ^ q1) | (tq1.x ^ q2) | tq1.y | tq1.z | tq1.w;
            const uint32_t dfs = (ts0.z ^ s0) | (ts0.w ^ s1) | (ts1.x ^ s2) | ts1.y | ts1.z | ts1.w;
            const bool okq = tq0.x != 0u && dfq == 0u, oks = ts0.x != 0u && dfs == 0u;
            // a miss may be a name parked in an overflow list (names with equal 32-bit hashes): rare route
            if ((okq && oks) || (a.qnov | a.snov) == 0u) {
              fastdone = true;
              if (!(okq && oks)) what = WH_UNMAPPED;
              else {
                const uint32_t qi = tq0.x - 1u, si = ts0.x - 1u;
                bool keep = true;
                if (a.n_excl) {
                  const uint64_t key = ((uint64_t)qi << 32) | si;
                  int64_t l = 0, h = a.n_excl;
                  while (l < h) { const int64_t m = (l + h) >> 1; if (a.excl[m] < key) l = m + 1; else h = m; }
                  if (l < a.n_excl && a.excl[l] == key) { keep = false; what = WH_EXCL; }
                }
                if (keep) {
                  what = WH_MAPPED;
                  const uint32_t sbits = __float_as_uint(L.score);
                  rec = make_uint4(qi, si, sbits, ((uint32_t)s << 1) | (uint32_t)L.eflag);
                  if (a.best && L.score > 0.f) {        // best_score[name] = max(score), starting at 0.0:
                    // two reductions (no return value, nothing to wait for).  lastal / BLAST write all alignments of a
                    // query together, so the lanes of a warp usually carry the SAME query: when they all do, one lane sends
                    // the warp's maximum (32 same-address reductions from one warp serialise in L2 -- measured 9 % of the
                    // kernel on a query-grouped table; a general __match_any_sync costs 26 % on a shuffled one, so only the
                    // all-equal case is merged, and only in the MERGE_Q instantiation of the kernel, which the host picks when
                    // a sample of the text showed runs of equal queries: the test alone costs 2 % on a shuffled table);
                    // subjects are spread, each lane sends its own
                    if constexpr (MERGE_Q) {
                      const unsigned act = __activemask();
                      const int lead = __ffs((int)act) - 1;
                      const uint32_t nid0 = __shfl_sync(act, tq0.y, lead);
                      if (__all_sync(act, tq0.y == nid0)) {
                        const uint32_t gmax = __reduce_max_sync(act, sbits);  // positive float32 bit patterns order like integers
                        if ((int)(threadIdx.x & 31) == lead) atomicMax(a.best + nid0, gmax);
                      } else {
                        atomicMax(a.best + tq0.y, sbits);
                      }
                    } else {
                      atomicMax(a.best + tq0.y, sbits);
                    }
                    atomicMax(a.best + ts0.y, sbits);
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  if (fastdone && what != WH_MAPPED) atomicAdd(c.cnt + what, 1u);   // comment / unmapped / excluded
  return fastdone;
}
// the lines parse_fast7 refused: version-5 route, then the generic scanner
__device__ __noinline__ uint4 parse_slow7(const TokArgs& a, const uint8_t* sm, const uint32_t* s_w, const uint32_t* s_n,
                                          const int64_t base, const int s, const int e, const int win, const Counters7 c) {
  const SlowOut so = slow_line(a, sm, s_w, s_n, base, s, e, win);
  uint4 rec = so.rec;
  int what = so.what;
  if (what == WH_MAPPED) {
    if (a.n_excl) {
      const uint64_t key = ((uint64_t)rec.x << 32) | rec.y;
      int64_t l = 0, h = a.n_excl;
      while (l < h) { const int64_t m = (l + h) >> 1; if (a.excl[m] < key) l = m + 1; else h = m; }
      if (l < a.n_excl && a.excl[l] == key) { what = WH_EXCL; rec.x = JX_NOREC; }
    }
    if (what == WH_MAPPED && a.best && __uint_as_float(rec.z) > 0.f) {
      atomicMax(a.best + a.qnid[rec.x], rec.z);
      atomicMax(a.best + a.snid[rec.y], rec.z);
    }
  }
  if (what != WH_MAPPED) {
    atomicAdd(c.cnt + what, 1u);
    if (what == WH_HARD || what == WH_LONG) atomicMin(c.stats + 7, (unsigned long long)(base + s) + 1ull);
  }
  return rec;
}

// A window full of tiny lines (more line starts than the shared list holds): every lane walks the line starts of its
// own 128-byte segment, ordinals from the warp prefix.  Rare; generic route only.
__device__ __noinline__ void walk_tiny_lines(const TokArgs& a, const uint8_t* sm, const uint32_t* s_w, const uint32_t* s_n,
                                             const int64_t base, const uint32_t st0, const uint32_t st1, const uint32_t st2,
                                             const uint32_t st3, uint32_t k, const uint32_t cap, uint4* slots, Counters7 c) {
  const int lane = threadIdx.x & 31;
  const uint32_t st[4] = {st0, st1, st2, st3};
#pragma unroll 1
  for (int h = 0; h < 4; ++h) {
    uint32_t m = st[h];
    while (m) {
      const int s = lane * 128 + h * 32 + __ffs((int)m) - 1;
      m &= m - 1u;
      int e = s;
      while (e < W7_WIN && !jx::is_term(sm[e])) ++e;
      if (e >= W7_WIN) while (base + e < a.n && !jx::is_term(a.text[base + e])) ++e;
      SlowOut so;
      if (e > s && sm[s] == '#') { so.rec = make_uint4(JX_NOREC, 0u, 0u, 0u); so.what = WH_COMMENT; }
      else so = slow_line(a, sm, s_w, s_n, base, s, e, W7_WIN);
      if (so.what == WH_MAPPED) {
        if (a.n_excl) {
          const uint64_t key = ((uint64_t)so.rec.x << 32) | so.rec.y;
          int64_t l = 0, hh = a.n_excl;
          while (l < hh) { const int64_t mm = (l + hh) >> 1; if (a.excl[mm] < key) l = mm + 1; else hh = mm; }
          if (l < a.n_excl && a.excl[l] == key) { so.what = WH_EXCL; so.rec.x = JX_NOREC; }
        }
        if (so.what == WH_MAPPED && a.best && __uint_as_float(so.rec.z) > 0.f) {
          atomicMax(a.best + a.qnid[so.rec.x], so.rec.z);
          atomicMax(a.best + a.snid[so.rec.y], so.rec.z);
        }
      }
      if (so.what != WH_MAPPED) {
        atomicAdd(c.cnt + so.what, 1u);
        if (so.what == WH_HARD || so.what == WH_LONG) atomicMin(c.stats + 7, (unsigned long long)(base + s) + 1ull);
      }
      if (k < cap) slots[k] = so.rec;
      ++k;
    }
  }
}

// 32 bytes of 7-bit text -> one word of each mask (see classify64_ascii)
__device__ __forceinline__ void classify32_ascii(const uint4& v0, const uint4& v1, uint32_t& wm, uint32_t& nm, uint32_t& any) {
  const uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t pw[4], pn[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a = x[2 * k], b = x[2 * k + 1];
    any |= a | b;
    const uint32_t twa = a + 0x5F5F5F5Fu, twb = b + 0x5F5F5F5Fu;
    const uint32_t una = ((a ^ 0x30303030u) + 0x76767676u) & ((a ^ 0x09090909u) + 0x7F7F7F7Fu);
    const uint32_t unb = ((b ^ 0x30303030u) + 0x76767676u) & ((b ^ 0x09090909u) + 0x7F7F7F7Fu);
    pw[k] = (((~twa >> 4) & 0x08080808u) | (~twb & 0x80808080u)) * 0x00204081u;
    pn[k] = (((una >> 4) & 0x08080808u) | (unb & 0x80808080u)) * 0x00204081u;
  }
  wm = jx::top_bytes(pw[0], pw[1], pw[2], pw[3]);
  nm = jx::top_bytes(pn[0], pn[1], pn[2], pn[3]);
}

// 16 bytes -> 16 bits of each mask (low half-word of the results)
__device__ __forceinline__ void classify16_ascii(const uint4& v, uint32_t& wm, uint32_t& nm, uint32_t& any) {
  const uint32_t x[4] = {v.x, v.y, v.z, v.w};
  uint32_t pw[2], pn[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const uint32_t a = x[2 * k], b = x[2 * k + 1];
    any |= a | b;
    const uint32_t twa = a + 0x5F5F5F5Fu, twb = b + 0x5F5F5F5Fu;
    const uint32_t una = ((a ^ 0x30303030u) + 0x76767676u) & ((a ^ 0x09090909u) + 0x7F7F7F7Fu);
    const uint32_t unb = ((b ^ 0x30303030u) + 0x76767676u) & ((b ^ 0x09090909u) + 0x7F7F7F7Fu);
    pw[k] = (((~twa >> 4) & 0x08080808u) | (~twb & 0x80808080u)) * 0x00204081u;
    pn[k] = (((una >> 4) & 0x08080808u) | (unb & 0x80808080u)) * 0x00204081u;
  }
  wm = __byte_perm(pw[0], pw[1], 0x0073);          // byte 3 of each product; the upper half is not stored
  nm = __byte_perm(pn[0], pn[1], 0x0073);
}

struct Tok7Args {
  TokArgs t;
  int64_t sub_begin, nsub; // this launch handles the sub-tiles [sub_begin, nsub)
  uint32_t cap;            // record slots per sub-tile
};

template <bool MERGE_Q>
__global__ void __launch_bounds__(JX_TOK_THREADS, JX_W7_MINBLK) k_tokenize7(const __grid_constant__ Tok7Args A) {
  const TokArgs& a = A.t;
  constexpr int NW = JX_TOK_THREADS / 32;
  __shared__ __align__(128) uint8_t s_text[W7_NBUF][W7_STAGE];
  __shared__ __align__(8) uint64_t s_bar[W7_NBUF];
  // everything a warp owns sits in ONE struct: the masks, the line list, the to-do list and the counters are constant
  // offsets from a single base register (five separate arrays cost five address registers in the per-line code)
  struct __align__(16) Warp7 {
    uint32_t w[W7_WIN / 32 + 4];   // byte <= 0x20 (+ slack: 64-bit windows read past the window)
    uint32_t n[W7_WIN / 32 + 4];   // byte is neither a digit nor a tab
    uint16_t lp[W7_LC + 2];        // start offset of every line of the window
    uint8_t todo[W7_LC + 2];       // lines of the window left to the slow route
    uint32_t cnt[WH_NONE + 2];     // lines that are not plain records, by kind; [WH_NONE + 1] = lines
  };
  __shared__ Warp7 s_warp[NW];
  __shared__ uint32_t s_head;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Warp7& SW = s_warp[warp];
  uint32_t* const s_w = SW.w;
  uint32_t* const s_n = SW.n;
  uint16_t* const s_lpos = SW.lp;
  const Counters7 c{SW.cnt, a.stats};
  if (lane < WH_NONE + 2) SW.cnt[lane] = 0u;
  bool overflow = false;
  uint32_t maxlines = 0;
  const bool v6 = !a.is_self;
  const int64_t nsub = A.nsub;
  const uint32_t cap = A.cap;

  // sub-tile of the CTA's ring position p; the stage of a sub-tile: bytes [base - 16, base + WIN + 32) of the text
  auto sub_of = [&](uint32_t p) -> int64_t { return A.sub_begin + (int64_t)blockIdx.x + (int64_t)p * gridDim.x; };
  auto arm = [&](uint32_t p) {                      // one thread: start the copy of position p's window into its buffer
    const int64_t sub = sub_of(p);
    if (sub >= nsub) return;
    const int b = (int)(p % W7_NBUF);
    const int64_t base = sub * W7_TILE;
    const int64_t src = base ? base - W7_PRE : 0;
    const int64_t left = a.n - src;
    const uint32_t want = base ? (uint32_t)W7_STAGE : (uint32_t)(W7_STAGE - W7_PRE);
    const uint32_t bytes = (uint32_t)(left < (int64_t)want ? left : (int64_t)want) & ~15u;
    if (bytes) {
      mbar_expect_tx(&s_bar[b], bytes);
      bulk_g2s(s_text[b] + (base ? 0 : W7_PRE), a.text + src, bytes, &s_bar[b]);
    }
  };
  if (tid == 0) {
    for (int b = 0; b < W7_NBUF; ++b) mbar_init(&s_bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_head = 0;
    for (uint32_t p = 0; p < (uint32_t)W7_NBUF; ++p) arm(p);
  }
  if (lane < 4) { s_w[W7_WIN / 32 + lane] = 0u; s_n[W7_WIN / 32 + lane] = 0u; }
  __syncthreads();                                  // the only CTA barrier of the kernel

  for (;;) {
    uint32_t p = 0;
    if (lane == 0) p = atomicAdd(&s_head, 1u);
    p = __shfl_sync(0xFFFFFFFFu, p, 0);
    const int64_t sub = sub_of(p);
    if (sub >= nsub) break;
    const int b = (int)(p % W7_NBUF);
    const uint32_t parity = (p / W7_NBUF) & 1u;
    const int64_t base = sub * W7_TILE;
    uint8_t* const stage = s_text[b];
    const uint8_t* const sm = stage + W7_PRE;
    // ---- the window: wait for the bulk copy; '\n' beyond the end of the text -------------------------------------------
    const int64_t src = base ? base - W7_PRE : 0;
    const int64_t left = a.n - src;
    const uint32_t want = base ? (uint32_t)W7_STAGE : (uint32_t)(W7_STAGE - W7_PRE);
    const bool full = left >= (int64_t)want;
    {
      const uint32_t avail = (uint32_t)(full ? (int64_t)want : left), bulk = avail & ~15u;
      if (bulk) {
        while (!mbar_try_wait(&s_bar[b], parity)) {}
      }
      if (!full) {                                  // last windows of the text only
        uint8_t* dst = stage + (base ? 0 : W7_PRE);
        for (uint32_t i = bulk + lane; i < want; i += 32) dst[i] = i < avail ? a.text[src + i] : (uint8_t)'\n';
        __syncwarp();
      }
    }
    // ---- byte classes: lane handles the 32-byte pieces lane, lane + 32, lane + 64, lane + 96 --------------------------
    // (16-byte chunks lane, lane + 32, ...: a quarter-warp reads 128 contiguous bytes per LDS.128 -- 32-byte pieces per
    //  lane cost twice the shared-memory wavefronts; the two 16-bit halves of a mask word come from neighbouring lanes)
    uint32_t any = 0;
    {
      uint16_t* const w16 = reinterpret_cast<uint16_t*>(s_w);
      uint16_t* const n16 = reinterpret_cast<uint16_t*>(s_n);
#pragma unroll
      for (int j = 0; j < W7_WIN / 512; ++j) {
        const int cidx = lane + 32 * j;
        const uint4 v = reinterpret_cast<const uint4*>(sm)[cidx];
        uint32_t wm, nm;
        classify16_ascii(v, wm, nm, any);
        w16[cidx] = (uint16_t)wm;
        n16[cidx] = (uint16_t)nm;
      }
    }
    if (__any_sync(0xFFFFFFFFu, (any & 0x80808080u) != 0u)) {   // a byte >= 0x80 in the window: exact classifier
      uint32_t hiacc = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4* p4 = reinterpret_cast<const uint4*>(sm + 32 * (lane + 32 * j));
        const uint4 v0 = p4[0], v1 = p4[1];
        const uint32_t w8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t wm, nm;
        jx::classify32(w8, wm, nm, hiacc);
        s_w[lane + 32 * j] = wm;
        s_n[lane + 32 * j] = nm;
      }
    }
    __syncwarp();
    // ---- terminators and line starts of this lane's 128-byte segment ----------------------------------------------------
    const uint4 W4 = reinterpret_cast<const uint4*>(s_w)[lane], N4 = reinterpret_cast<const uint4*>(s_n)[lane];
    const uint32_t wv[4] = {W4.x, W4.y, W4.z, W4.w}, nv[4] = {N4.x, N4.y, N4.z, N4.w};
    uint32_t term[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      uint32_t cnd = wv[h] & nv[h], tm = 0;       // a terminator is a byte <= 0x20 that is not a tab: checked by value
      while (cnd) {
        const int bit = __ffs((int)cnd) - 1;
        cnd &= cnd - 1u;
        const uint32_t ch = sm[lane * 128 + h * 32 + bit];
        if (ch == '\n' || ch == '\r') tm |= 1u << bit;
      }
      term[h] = tm;
    }
    uint32_t prevbit = __shfl_up_sync(0xFFFFFFFFu, term[3] >> 31, 1);
    if (lane == 0) prevbit = base == 0 ? 1u : (uint32_t)jx::is_term(sm[-1]);
    uint32_t st[4];
    st[0] = (term[0] << 1) | prevbit;
    st[1] = (term[1] << 1) | (term[0] >> 31);
    st[2] = (term[2] << 1) | (term[1] >> 31);
    st[3] = (term[3] << 1) | (term[2] >> 31);
    if (!full) {                                    // only positions that exist can start a line
      const int64_t remain = a.n - (base + (int64_t)lane * 128);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int64_t r = remain - 32 * h;
        if (r <= 0) st[h] = 0u; else if (r < 32) st[h] &= (1u << r) - 1u;
      }
    }
    const uint32_t cnt = __popc(st[0]) + __popc(st[1]) + __popc(st[2]) + __popc(st[3]);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += t; }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);                 // starts in the whole window
    const uint32_t nlines = __shfl_sync(0xFFFFFFFFu, inc, W7_TILE / 128 - 1); // ... in the sub-tile: this warp's lines
    const uint32_t pref = inc - cnt;
    if (nlines > cap) overflow = true;
    maxlines = max(maxlines, nlines);
    uint4* const slots = a.recs + (uint64_t)sub * cap;
    const bool listed = total <= (uint32_t)W7_LC;  // (else a window full of tiny lines: every lane walks its own line starts)
    if (listed) {
      uint32_t o = pref;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t m = st[h];
        const uint32_t p0 = (uint32_t)(lane * 128 + h * 32);
        while (m) { s_lpos[o++] = (uint16_t)(p0 + __ffs((int)m) - 1); m &= m - 1u; }
      }
    }
    __syncwarp();
    if (listed) {
      auto line_end = [&](uint32_t k, int s) -> int {
        if (k + 1 < total) return (int)s_lpos[k + 1] - 1;
        int e = s;                                    // no line start after this one inside the window: find the terminator
        while (e < W7_WIN && !jx::is_term(sm[e])) ++e;
        if (e >= W7_WIN) while (base + e < a.n && !jx::is_term(a.text[base + e])) ++e;
        return e;
      };
      uint32_t ntodo = 0;                             // lines the fast route refused (warp-uniform count)
      for (uint32_t k0 = 0; k0 < nlines; k0 += 32) {
        const uint32_t k = k0 + lane;
        bool later = false;
        if (k < nlines) {
          const int s = s_lpos[k];
          const int e = line_end(k, s);
          uint4 rec;
          later = !parse_fast7<MERGE_Q>(a, sm, s_w, s_n, s, e, W7_WIN, v6, c, rec);
          if (!later && k < cap) slots[k] = rec;
        }
        const unsigned lm = __ballot_sync(0xFFFFFFFFu, later);
        if (later) SW.todo[ntodo + __popc(lm & ((1u << lane) - 1u))] = (uint8_t)k;
        ntodo += __popc(lm);
      }
      if (ntodo) {
        __syncwarp();
        for (uint32_t t = lane; t < ntodo; t += 32) {
          const uint32_t k = SW.todo[t];
          const int s = s_lpos[k];
          const uint4 rec = parse_slow7(a, sm, s_w, s_n, base, s, line_end(k, s), W7_WIN, c);
          if (k < cap) slots[k] = rec;
        }
      }
    } else if (lane < W7_TILE / 128) {
      walk_tiny_lines(a, sm, s_w, s_n, base, st[0], st[1], st[2], st[3], pref, cap, slots, c);
    }
    for (uint32_t k = nlines + lane; k < cap; k += 32) slots[k] = make_uint4(JX_NOREC, 0u, 0u, 0u);
    if (lane == 0) SW.cnt[WH_NONE + 1] += nlines;
    __syncwarp();                                     // every lane is done with the window
    if (lane == 0) arm(p + W7_NBUF);
  }

  __syncwarp();
  if (lane == 0) {
    const uint32_t* cw = SW.cnt;
    const uint32_t lines = cw[WH_NONE + 1];
    const uint32_t other = cw[WH_COMMENT] + cw[WH_UNMAPPED] + cw[WH_HARD] + cw[WH_LONG] + cw[WH_EXCL] + cw[WH_SELF] + cw[WH_NONE];
    if (cw[WH_COMMENT]) atomicAdd(a.stats + 1, (unsigned long long)cw[WH_COMMENT]);
    if (lines - other) atomicAdd(a.stats + 2, (unsigned long long)(lines - other));       // mapped = everything else
    if (cw[WH_UNMAPPED]) atomicAdd(a.stats + 3, (unsigned long long)cw[WH_UNMAPPED]);
    if (cw[WH_HARD]) atomicAdd(a.stats + 4, (unsigned long long)cw[WH_HARD]);
    if (cw[WH_LONG]) atomicAdd(a.stats + 5, (unsigned long long)cw[WH_LONG]);
    if (cw[WH_EXCL]) atomicAdd(a.stats + 8, (unsigned long long)cw[WH_EXCL]);
    if (cw[WH_SELF]) atomicAdd(a.stats + 9, (unsigned long long)cw[WH_SELF]);
    if (lines) atomicAdd(a.stats + 0, (unsigned long long)lines);
  }
  if (__any_sync(0xFFFFFFFFu, overflow) && lane == 0) atomicExch(a.stats + 6, 1ull);
  if (lane == 0 && maxlines) atomicMax(a.stats + 11, (unsigned long long)maxlines);
}

}  // namespace

int jx_run_tokenizer(jx_ctx* ctx, Arena& arena, const char* text, int64_t nbytes, int on_device,
                     const TokMode& mode, TokOut& out) {
  if (!mode.names && !(mode.filter && !mode.f_ids) && !ctx->pair_ready) { ctx->err = "jx_bed_load / jx_pair_setup first"; return JX_EARG; }
  if (nbytes < 0 || (nbytes && !text)) return JX_EARG;
  const int64_t ntiles = (nbytes + JX_TILE - 1) / JX_TILE;
  out.ntiles = ntiles;
  out.nbytes = nbytes;
  const uint8_t* d_text = nullptr;
  constexpr int64_t CHUNK = 64ll << 20;   // H2D pipelining granule (multiple of the tile size)
  const int64_t nchunks = (nbytes + CHUNK - 1) / CHUNK;
  std::vector<cudaEvent_t> copied;
  bool held = false;
  if (on_device) {
    if (((uintptr_t)text & 15u) != 0) { ctx->err = "device text must be 16-byte aligned"; return JX_EARG; }
    d_text = (const uint8_t*)text;
    // text uploaded by jx_text_upload: its chunk copies may still be in flight on the copy stream
    held = ctx->held_text && d_text == ctx->held_text && nbytes == ctx->held_n;
    if (held) copied = ctx->held_events;
  } else {
    JX_ALLOC(buf, uint8_t, (size_t)nbytes + 16, false);
    d_text = buf;
    // the copy stream must not run ahead of the allocation on the compute stream
    cudaEvent_t ready;
    JX_CK(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    JX_CK(cudaEventRecord(ready, ctx->stream));
    JX_CK(cudaStreamWaitEvent(ctx->copy_stream, ready, 0));
    cudaEventDestroy(ready);
    for (int64_t c = 0; c < nchunks; ++c) {
      int64_t off = c * CHUNK, len = std::min(CHUNK, nbytes - off);
      JX_CK(cudaMemcpyAsync(buf + off, text + off, (size_t)len, cudaMemcpyHostToDevice, ctx->copy_stream));
      cudaEvent_t ev;
      JX_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      JX_CK(cudaEventRecord(ev, ctx->copy_stream));
      copied.push_back(ev);
    }
  }
  out.d_text = d_text;
  JX_TRACE("tok: text ready");

  // record capacity: lines are >= 24 bytes in any real table; retry with the exact worst case
  // (every byte a line) only if that guess overflows.
  JX_ALLOC(d_stats, unsigned long long, 12, true);
  JX_ALLOC(d_ticket, uint32_t, (size_t)nchunks + 2, true);   // one ticket counter per launch
  const bool to_cache = held && ((mode.fill_cache && !ctx->is_self && mode.n_excl == 0 && !mode.neardiag) || mode.persist);
  if (to_cache) ctx->cache.valid = false;    // about to overwrite the retained records
  const bool use7 = !mode.names && !mode.filter && !ctx->tok_v5 && !ctx->tok_v6;   // the path's kernel: version 7 (fixed slots)
  const int64_t nsub = (nbytes + W7_TILE - 1) / W7_TILE;
  uint64_t* d_desc = nullptr;
  if (use7) {
    // no descriptor chain
  } else if (to_cache) {
    if (ctx->cache.desc_cap < (size_t)ntiles + 1) {
      if (ctx->cache.desc) cudaFree(ctx->cache.desc);
      ctx->cache.desc = nullptr; ctx->cache.desc_cap = 0;
      JX_CK(cudaMalloc((void**)&ctx->cache.desc, ((size_t)ntiles + 1) * sizeof(uint64_t)));
      ctx->cache.desc_cap = (size_t)ntiles + 1;
    }
    d_desc = ctx->cache.desc;
    JX_CK(cudaMemsetAsync(d_desc, 0, ((size_t)ntiles + 1) * sizeof(uint64_t), ctx->stream));
  } else {
    JX_ALLOC(dd, uint64_t, (size_t)ntiles, true);
    d_desc = dd;
  }
  out.desc = d_desc;
  out.li.desc = d_desc; out.li.ntiles = ntiles; out.li.cap = 0; out.li.sub_bytes = 0;
  const SideDev& Q = ctx->side[0];
  const SideDev& S = ctx->side[1];
  uint32_t slot_cap = std::max<uint32_t>(ctx->tok_cap, 8u);     // version 7: record slots per sub-tile
  for (int attempt = 0; attempt < 2; ++attempt) {
    uint64_t cap64 = attempt == 0 ? (uint64_t)(nbytes / 20 + 1024) : (uint64_t)nbytes + 1;
    if (use7) cap64 = (uint64_t)std::max<int64_t>(nsub, 1) * slot_cap;
    if (cap64 > 0xFFFFFFF0ull) { ctx->err = use7 ? "more than 2^32 record slots in one text" : "more than 2^32 lines in one text"; return JX_EARG; }
    uint4* recs = nullptr;
    if (to_cache) {
      if (ctx->cache.recs_cap < (size_t)cap64) {
        if (ctx->cache.recs) cudaFree(ctx->cache.recs);
        ctx->cache.recs = nullptr; ctx->cache.recs_cap = 0;
        JX_CK(cudaMalloc((void**)&ctx->cache.recs, (size_t)cap64 * sizeof(uint4)));
        ctx->cache.recs_cap = (size_t)cap64;
      }
      recs = ctx->cache.recs;
    } else {
      JX_ALLOC(rr, uint4, (size_t)cap64, false);
      recs = rr;
    }
    out.recs = recs;
    uint4* hashes = nullptr;
    if (mode.names) { JX_ALLOC(hh, uint4, (size_t)cap64, false); hashes = hh; }
    out.hashes = hashes;
    if (attempt) {
      JX_CK(cudaMemsetAsync(d_stats, 0, 12 * sizeof(unsigned long long), ctx->stream));
      JX_CK(cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t) * ((size_t)nchunks + 2), ctx->stream));
      if (d_desc) JX_CK(cudaMemsetAsync(d_desc, 0, (size_t)ntiles * sizeof(uint64_t), ctx->stream));
      if (mode.best) JX_CK(cudaMemsetAsync(mode.best, 0, (size_t)ctx->n_name_ids * sizeof(uint32_t), ctx->stream));
    }
    {
      unsigned long long init = ~0ull;
      JX_CK(cudaMemcpyAsync(d_stats + 7, &init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    }
#ifdef JX_EXP_NOLB
    JX_CK(cudaMemsetAsync(recs, 0xFF, (size_t)cap64 * sizeof(uint4), ctx->stream));
#endif
    JX_TRACE("tok: allocs");
    TokArgs a;
    a.text = d_text; a.n = nbytes; a.desc = d_desc; a.ticket = d_ticket; a.tile_begin = 0; a.tile_end = 0; a.recs = recs; a.rec_cap = (uint32_t)cap64;
    a.hashes = hashes; a.seed = mode.seed;
    a.f_score = mode.f_score; a.f_pctid = mode.f_pctid; a.f_evalue = mode.f_evalue; a.f_hitlen = mode.f_hitlen;
    a.f_noself = mode.f_noself; a.f_inverse = mode.f_inverse; a.f_ids = mode.f_ids;
    a.qtab = Q.table; a.qdisp = Q.disp; a.qbshift = Q.bshift; a.qcshift = Q.cshift; a.qmeta = Q.name_meta; a.qblob = Q.blobw;
    a.stab = S.table; a.sdisp = S.disp; a.sbshift = S.bshift; a.scshift = S.cshift; a.smeta = S.name_meta; a.sblob = S.blobw;
    a.qovh = Q.ovf_hash; a.qovr = Q.ovf_rank; a.qnov = Q.n_ovf; a.sovh = S.ovf_hash; a.sovr = S.ovf_rank; a.snov = S.n_ovf;
    a.qchr = Q.chr_lex; a.schr = S.chr_lex; a.qnid = Q.name_id; a.snid = S.name_id;
    a.best = mode.best; a.excl = mode.excl; a.n_excl = mode.n_excl;
    a.strip = mode.strip; a.is_self = ctx->is_self ? 1 : 0; a.neardiag = mode.neardiag;
    a.merge_q = (!on_device && nbytes > 0) ? jx_sample_query_runs((const uint8_t*)text, (size_t)nbytes) : (held ? ctx->tok_grouped : 0);
    a.stats = d_stats;
    const int64_t tiles_per_chunk = CHUNK / JX_TILE;
    // persistent CTAs: as many as are resident at once
    if (!ctx->tok_resident) {
      int per_sm = 0, n_sm = 0;
      if (ctx->tok_v5) JX_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tokenize<0>, JX_TOK_THREADS, 0));
      else if (ctx->tok_v6) JX_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tokenize6, JX_TOK_THREADS, 0));
      else {
        int p0 = 0, p1 = 0;
        JX_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p0, k_tokenize7<false>, JX_TOK_THREADS, 0));
        JX_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p1, k_tokenize7<true>, JX_TOK_THREADS, 0));
        per_sm = std::min(p0, p1);
      }
      JX_CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device));
      ctx->tok_resident = (int64_t)std::max(per_sm, 1) * std::max(n_sm, 1);
    }
    const int64_t resident = ctx->tok_resident;
    // copies still in flight: one launch per 64 MiB chunk, each behind its copy; a text that is already
    // resident is tokenized by ONE launch (no ramp-up / ramp-down per chunk)
    bool pending = (!on_device || held) && attempt == 0 && !copied.empty();
    if (pending && held && cudaEventQuery(copied.back()) == cudaSuccess) pending = false;
#ifdef JX_TOK_ALWAYS_CHUNKED
    const int64_t launch_tiles = tiles_per_chunk;
#else
    const int64_t launch_tiles = pending ? tiles_per_chunk : std::max<int64_t>(ntiles, 1);
#endif
    // (version 7 counts in sub-tiles; 64 MiB is not a multiple of 3840 bytes, so a launch's last window may start in the
    //  next chunk's first bytes: the launch waits for that copy as well, see below)
    const int64_t units = use7 ? nsub : ntiles;
    const int64_t units_per_chunk = use7 ? (CHUNK + W7_TILE - 1) / W7_TILE : tiles_per_chunk;
    const int64_t launch_units = pending ? units_per_chunk : std::max<int64_t>(units, 1);
    const int64_t nlaunch = use7 ? (units + launch_units - 1) / launch_units : (ntiles + launch_tiles - 1) / launch_tiles;
    for (int64_t c = 0; c < nlaunch; ++c) {
      int64_t t0 = c * (use7 ? launch_units : launch_tiles), t1 = std::min(use7 ? nsub : ntiles, t0 + (use7 ? launch_units : launch_tiles));
      if (pending) {
        // a tile's look-ahead may reach into the next chunk: wait for that copy too
        int64_t last_chunk = c + 1;
        if (use7) last_chunk = (std::min<int64_t>(t1 * W7_TILE + W7_WIN + 64, nbytes) - 1) / CHUNK;
        JX_CK(cudaStreamWaitEvent(ctx->stream, copied[(size_t)std::min(last_chunk, nchunks - 1)], 0));
      }
      cudaEvent_t e0, e1;
      JX_CK(cudaEventCreate(&e0));
      JX_CK(cudaEventCreate(&e1));
      JX_CK(cudaEventRecord(e0, ctx->stream));
      a.ticket = d_ticket + c; a.tile_begin = (uint32_t)t0; a.tile_end = (uint32_t)t1;
      if (mode.filter) JX_LAUNCH(k_tokenize<2>, (unsigned)std::min<int64_t>(t1 - t0, resident), JX_TOK_THREADS, 0, a);
      else if (mode.names) JX_LAUNCH(k_tokenize<1>, (unsigned)std::min<int64_t>(t1 - t0, resident), JX_TOK_THREADS, 0, a);
      else if (ctx->tok_v5) JX_LAUNCH(k_tokenize<0>, (unsigned)std::min<int64_t>(t1 - t0, resident), JX_TOK_THREADS, 0, a);
      else if (ctx->tok_v6) JX_LAUNCH(k_tokenize6, (unsigned)std::min<int64_t>(t1 - t0, resident), JX_TOK_THREADS, 0, a);
      else {
        Tok7Args a7; a7.t = a; a7.nsub = t1; a7.sub_begin = t0; a7.cap = slot_cap;
        const unsigned g7 = (unsigned)std::min<int64_t>((t1 - t0 + 3) / 4, resident);
        if (a.merge_q) JX_LAUNCH(k_tokenize7<true>, g7, JX_TOK_THREADS, 0, a7);
        else JX_LAUNCH(k_tokenize7<false>, g7, JX_TOK_THREADS, 0, a7);
      }
      JX_CK(cudaEventRecord(e1, ctx->stream));
      ctx->tok_events.emplace_back(e0, e1);
    }
    JX_TRACE("tok: launched+done");
    ctx->tok_bytes += nbytes;
    JX_CK(cudaGetLastError());
    JX_CK(cudaMemcpyAsync(out.stats, d_stats, 12 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->stream));
    jx_harvest_tok_events(ctx);
    if (use7) {
      // stats[11] = the largest number of lines of any sub-tile: the next text (and the retry, when this one did
      // not fit) gets that many slots plus an eighth
      const uint32_t maxl = (uint32_t)out.stats[11];
      const uint32_t want = std::min<uint32_t>(std::max<uint32_t>(((maxl + maxl / 8 + 2 + 7) / 8) * 8, 16u), (uint32_t)W7_TILE);
      if (!out.stats[6]) {
        out.li.desc = nullptr; out.li.ntiles = 0; out.li.cap = slot_cap; out.li.sub_bytes = (uint32_t)W7_TILE;
        ctx->tok_cap = want;
        break;
      }
      if (attempt == 1) { ctx->err = "record slots overflow"; return JX_EINTERNAL; }
      slot_cap = std::max(want, maxl);
      continue;
    }
    if (!out.stats[6]) break;
    if (attempt == 1) { ctx->err = "record buffer overflow"; return JX_EINTERNAL; }
  }
  if (!held) for (cudaEvent_t ev : copied) cudaEventDestroy(ev);
  out.n_lines = use7 ? (uint32_t)((uint64_t)std::max<int64_t>(nsub, 0) * slot_cap) : (uint32_t)out.stats[0];
  ctx->tok_lines += (int64_t)out.stats[0];
  if (to_cache && !out.stats[4] && !out.stats[5] && !out.stats[10]) {
    auto& c = ctx->cache;
    c.valid = true; c.plain = !ctx->is_self && mode.n_excl == 0 && !mode.neardiag; c.text = d_text; c.nbytes = nbytes; c.strip = mode.strip; c.ntiles = ntiles; c.n_lines = out.n_lines; c.cap = out.li.cap; c.sub_bytes = out.li.sub_bytes;
    memcpy(c.stats, out.stats, sizeof c.stats);
  }
  if (out.stats[10]) { ctx->err = "tokenizer look-back timed out"; return JX_EINTERNAL; }
  if (out.stats[5]) {
    ctx->err = "BLAST name of 128+ bytes near byte offset " + std::to_string(out.stats[7] - 1) +
               " (overflows cblast.pyx's char[128]; undefined behaviour in the reference)";
    return JX_EUNSUPPORTED;
  }
  if (out.stats[4]) {
    ctx->err = "numeric field the device tokenizer cannot convert exactly (inf/nan/hex float, >19 significant "
               "digits or dangling exponent) near byte offset " + std::to_string(out.stats[7] - 1);
    return JX_EUNSUPPORTED;
  }
  return JX_OK;
}

extern "C" {

int jx_set_record_reuse(jx_ctx* ctx, int on) {
  if (!ctx) return JX_EARG;
  ctx->record_reuse = on != 0;
  if (!on) ctx->cache.valid = false;
  return JX_OK;
}

int jx_text_release(jx_ctx* ctx) {
  if (!ctx) return JX_EARG;
  cudaSetDevice(ctx->device);
  ctx->cache.valid = false;
  if (ctx->held_text) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamSynchronize(ctx->stream);
    for (cudaEvent_t ev : ctx->held_events) cudaEventDestroy(ev);
    ctx->held_events.clear();
    cudaFree(ctx->held_text);
    ctx->held_text = nullptr;
    ctx->held_n = 0;
    ctx->held_cap = 0;
  }
  return JX_OK;
}

int jx_text_upload(jx_ctx* ctx, const char* host_text, int64_t nbytes, uint64_t* device_ptr) {
  if (!ctx || !device_ptr || nbytes < 0 || (nbytes && !host_text)) return JX_EARG;
  cudaSetDevice(ctx->device);
  ctx->cache.valid = false;
  constexpr int64_t CHUNK = 64ll << 20;
  if (ctx->held_text && ctx->held_cap >= nbytes + 16) {
    // reuse the allocation: earlier work on it must have finished before it is overwritten
    JX_CK(cudaStreamSynchronize(ctx->stream));
    JX_CK(cudaStreamSynchronize(ctx->copy_stream));
    for (cudaEvent_t ev : ctx->held_events) cudaEventDestroy(ev);
    ctx->held_events.clear();
  } else {
    jx_text_release(ctx);
    uint8_t* buf = nullptr;
    JX_CK(cudaMalloc((void**)&buf, (size_t)nbytes + 16));
    ctx->held_text = buf;
    ctx->held_cap = nbytes + 16;
  }
  uint8_t* buf = ctx->held_text;
  ctx->tok_grouped = jx_sample_query_runs((const uint8_t*)host_text, (size_t)nbytes);
  for (int64_t off = 0; off < nbytes; off += CHUNK) {
    const int64_t len = std::min(CHUNK, nbytes - off);
    JX_CK(cudaMemcpyAsync(buf + off, host_text + off, (size_t)len, cudaMemcpyHostToDevice, ctx->copy_stream));
    cudaEvent_t ev;
    JX_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    JX_CK(cudaEventRecord(ev, ctx->copy_stream));
    ctx->held_events.push_back(ev);
  }
  ctx->held_n = nbytes;
  *device_ptr = (uint64_t)(uintptr_t)buf;
  return JX_OK;
}

}  // extern "C"

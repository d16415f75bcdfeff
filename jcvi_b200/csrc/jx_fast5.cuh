// Tokenizer fast path, version 5: byte classes are computed ONCE per tile, cooperatively and
// branch-free (every lane does the same work), and kept as two bit-per-byte masks in shared
// memory:
//     W  : byte <= 0x20          (tab, the terminators, space and the other control bytes)
//     N  : byte is neither a decimal digit nor a tab
// With them the per-line parse of a REGULAR line (12 fields, 11 single tabs, <= 128 bytes; what
// LAST / BLAST+ / DIAMOND / MMseqs2 write) needs no loop over the text at all: field boundaries
// are the set bits of W, "all digits" is "no bit of N", and the numeric fields are converted with
// word arithmetic.  The instruction stream is the same for every lane of a warp, which is what
// the previous (byte- and word-scanning) versions lacked.  Anything irregular is refused here and
// goes to jx::tokenize_line, the generic sscanf-faithful scanner, so this path only has to be
// exact on what it accepts (tests/test_parse_host.py checks the two against each other).
//
// Reference semantics being reproduced: src/jcvi/formats/cblast.pyx:15,106-112 (sscanf of the 12
// fields), src/jcvi/utils/cbook.py:308-329 (gene_name).
#pragma once
#include "jx_parse.cuh"

namespace jx {

// ---- phase A: classification -------------------------------------------------------------------
// 4 text bytes -> 0x80 flags per byte: `wf` = byte <= 0x20, `df` = byte is a digit or a tab (N is the
// complement).  `hi` accumulates bit 7 of (low7 + 4): set when a byte's low seven bits are >= 0x7C,
// i.e. the tile may contain a '|' (gene_name cuts there).
JX_HD void classify4(uint32_t x, uint32_t& wf, uint32_t& df, uint32_t& hi) {
  const uint32_t a = x & 0x7F7F7F7Fu;
  const uint32_t nx = ~x & 0x80808080u;                          // bytes < 0x80
  wf = ~(a + 0x5F5F5F5Fu) & nx;                                  // low7 < 0x21
  const uint32_t ud = (a ^ 0x30303030u) + 0x76767676u;           // bit7 clear iff low7 in '0'..'9'
  const uint32_t ut = (a ^ 0x09090909u) + 0x7F7F7F7Fu;           // bit7 clear iff low7 == 0x09
  df = ~(ud & ut) & nx;
  hi |= a + 0x04040404u;
}
// flags of 8 consecutive bytes (two words) -> 8 mask bits in text order, in the top byte of the result
JX_HD uint32_t gather8_top(uint32_t f0, uint32_t f1) { return ((f0 >> 4) + f1) * 0x00204081u; }
JX_HD uint32_t top_bytes(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {   // byte3 of each -> one word
#if defined(__CUDA_ARCH__)
  return __byte_perm(__byte_perm(p0, p1, 0x0073), __byte_perm(p2, p3, 0x0073), 0x5410);
#else
  return (p0 >> 24) | ((p1 >> 24) << 8) | ((p2 >> 24) << 16) | (p3 & 0xFF000000u);
#endif
}

// 32 text bytes (8 aligned words) -> one word of each mask (W: byte <= 0x20, N: neither digit nor tab)
JX_HD void classify32(const uint32_t* w, uint32_t& wmask, uint32_t& nmask, uint32_t& hi) {
  uint32_t pw[4], pd[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    uint32_t w0, d0, w1, d1;
    classify4(w[2 * k], w0, d0, hi);
    classify4(w[2 * k + 1], w1, d1, hi);
    pw[k] = gather8_top(w0, w1);
    pd[k] = gather8_top(d0, d1);
  }
  wmask = top_bytes(pw[0], pw[1], pw[2], pw[3]);
  nmask = ~top_bytes(pd[0], pd[1], pd[2], pd[3]);
}

// ---- small bit helpers ----------------------------------------------------------------------------
JX_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {      // low word of (hi:lo) >> sh, 0 <= sh < 32
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, (uint32_t)sh);
#else
  return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
JX_HD int ctz32(uint32_t x) {                                    // x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
JX_HD int ctz64(uint64_t x) {                                    // x != 0
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
JX_HD int hibit32(uint32_t x) {                                  // x != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)x);
#else
  return 31 - __builtin_clz(x);
#endif
}
JX_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
JX_HD uint32_t below32(int k) { return k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u); }   // bits < k, 0 <= k
JX_HD uint64_t below64(int k) { return k >= 64 ? ~0ull : ((1ull << k) - 1ull); }      // bits < k, 0 <= k

// `M` yields mask words: m(i) covers bytes [32 i, 32 i + 32) of the buffer
template <class M>
JX_HD uint32_t bits32(const M& m, int pos) {                     // mask bits of bytes [pos, pos + 32)
  const int i = pos >> 5;
  return funnel_r(m(i), m(i + 1), pos & 31);
}
template <class M>
JX_HD uint64_t bits64(const M& m, int pos) {                     // mask bits of bytes [pos, pos + 64)
  const int i = pos >> 5, sh = pos & 31;
  const uint32_t x0 = m(i), x1 = m(i + 1), x2 = m(i + 2);
  return (uint64_t)funnel_r(x0, x1, sh) | ((uint64_t)funnel_r(x1, x2, sh) << 32);
}

// ---- numeric fields from words ------------------------------------------------------------------------
// four ASCII-minus-'0' digits d0 d1 d2 d3 (d0 in the low byte) -> d0 d1 d2 d3 as a number
JX_HD uint32_t digits4(uint32_t w) {
  const uint32_t t = w * 10u + (w >> 8);              // byte0 = 10 d0 + d1, byte2 = 10 d2 + d3 (no carries: <= 99)
  return (t & 0xFFu) * 100u + ((t >> 16) & 0xFFu);
}
// The first `n` (1..8) bytes of (w1:w0) are text; bytes [0,n) except an optional '.' at index `dot`
// (-1: none) are decimal digits.  Returns the digits as an integer (< 10^8).
JX_HD uint32_t digits_upto8(uint32_t w0, uint32_t w1, int n, int dot) {
  uint64_t v = (uint64_t)w0 | ((uint64_t)w1 << 32);
  int nd = n;
  if (dot >= 0) {                                      // close the gap left by the '.'
    const uint64_t lowm = (1ull << (8 * dot)) - 1ull;
    v = (v & lowm) | ((v >> 8) & ~lowm);
    --nd;
  }
  v = (v - 0x3030303030303030ull) << (8 * (8 - nd));   // right-justify: leading bytes become 0
  return digits4((uint32_t)v) * 10000u + digits4((uint32_t)(v >> 32));
}

#define JX_POW10U32 {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u}
#define JX_POW10F32 {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f}   /* all exact in float32 */
static const uint32_t kPow10u32Host[10] = JX_POW10U32;
static const float kPow10f32Host[11] = JX_POW10F32;
#if defined(__CUDACC__)
static __constant__ uint32_t kPow10u32Dev[10] = JX_POW10U32;
static __constant__ float kPow10f32Dev[11] = JX_POW10F32;
#endif
JX_HD uint32_t pow10_u32(int k) {                      // 0 <= k <= 9
#if defined(__CUDA_ARCH__)
  return kPow10u32Dev[k];
#else
  return kPow10u32Host[k];
#endif
}
JX_HD float pow10_f32(int k) {                         // 0 <= k <= 10
#if defined(__CUDA_ARCH__)
  return kPow10f32Dev[k];
#else
  return kPow10f32Host[k];
#endif
}

// A numeric field = bytes [p, p+L) of the buffer, `nb` = its N bits (bit i: byte p+i is not a digit).
// Accepts  D+ [ '.' D* ] [ (e|E) [+-] D{1,3} ]  and '.' D+ forms with a mantissa of <= 8 bytes and
// yields value = m * 10^e10 (m < 10^8, all digits kept).  Anything else: false (generic scanner).
template <class BR>
JX_HD bool decimal5(const BR& B, int p, int L, uint32_t nb, uint32_t& m, int& e10) {
  if (L < 1 || L > 14) return false;
  int dot = -1, epos = L, estart = L;
  bool eneg = false;
  if (nb) {
    const int b1 = ctz32(nb);
    const uint32_t c1 = B(p + b1);
    uint32_t rest = nb & (nb - 1u);
    int ep = -1;
    if (c1 == '.') {
      dot = b1;
      if (rest) { ep = ctz32(rest); rest &= rest - 1u; }
    } else {
      ep = b1;
    }
    if (ep >= 0) {
      if ((B(p + ep) | 0x20u) != 'e') return false;
      epos = ep; estart = ep + 1;
      if (rest) {                                        // only a sign right after the 'e' may remain
        if (rest != (1u << (ep + 1))) return false;
        const uint32_t cs = B(p + ep + 1);
        if (cs == '-') eneg = true; else if (cs != '+') return false;
        estart = ep + 2;
      }
    } else if (rest) return false;
  }
  const int nman = epos, nexp = L - estart;              // mantissa bytes (with the dot), exponent digits
  if (nman < 1 || nman > 8 || nman - (dot >= 0) < 1) return false;
  if (epos < L && (nexp < 1 || nexp > 3)) return false;
  m = digits_upto8(B.load32(p), B.load32(p + 4), nman, dot);
  int ex = 0;
  if (epos < L) {
    const uint32_t w = B.load32(p + L - 4);              // the exponent digits are the last bytes of the field
    const uint32_t keep = ~below32(8 * (4 - nexp));      // mask before subtracting: no borrow out of 'e' / '-'
    ex = (int)digits4((w & keep) - (0x30303030u & keep));
    if (eneg) ex = -ex;
  }
  e10 = ex - (dot >= 0 ? nman - 1 - dot : 0);
  return true;
}

// evalue: flag = (value < 1e-10), decided like dec_lt_1e10 (<= 8 significant digits: integer compare)
template <class BR>
JX_HD bool evalue5(const BR& B, int p, int L, uint32_t nb, int& flag) {
  uint32_t m; int e10;
  if (!decimal5(B, p, L, nb, m, e10)) return false;
  const int k = -10 - e10;                               // value < 1e-10  <=>  m < 10^k
  const int kc = k < 0 ? 0 : (k > 9 ? 9 : k);
  flag = (m == 0u || k >= 9 || (k > 0 && m < pow10_u32(kc))) ? 1 : 0;
  return true;
}

// score -> float32 exactly like strtof: an integer < 2^24 divided by an exact power of ten is one
// correctly rounded operation; everything else goes through dec_to_f32's exact routes
template <class BR>
JX_HD bool score5(const BR& B, int p, int L, uint32_t nb, float& out) {
  uint32_t m; int e10;
  if (!decimal5(B, p, L, nb, m, e10)) return false;
  if (m < (1u << 24) && e10 <= 0 && e10 >= -10) {
    const float fm = (float)m;
#if defined(__CUDA_ARCH__)
    out = __fdiv_rn(fm, pow10_f32(-e10));
#else
    out = fm / pow10_f32(-e10);
#endif
    return true;
  }
  Dec d; d.m = m; d.e10 = m ? e10 : 0; d.neg = false; d.ok = true; d.dropped = false; d.special = false;
  return dec_to_f32(d, out);
}

// ---- the line ----------------------------------------------------------------------------------------------
// [s, e) = the line without its terminator.  `B` = byte / unaligned-word view of the text,
// `mw` / `mn` = W / N mask words of the same buffer.  Three windows of mask bits cover the line:
//   64 bits from the line start     -> tabs 1..3 (both names and pctid must end inside it)
//   64 bits from tab 3              -> fields 4..10 up to tab 10
//   the last 32 bits of the line    -> tabs 10, 11, evalue and score
// Fills the name spans, score and evalue flag; `qbar`/`sbar` are left at -2 ("not searched": the
// caller knows whether the tile can contain a '|').
struct Line5Tail {              // what the second half of the parse needs: the two decimal fields
  int pe, le, ps, ls;          // absolute start and length of evalue / score
  uint32_t eb, sb;             // their N bits
};
// first half: structure, names, integer fields
template <class BR, class MW, class MN>
JX_HD bool fast_line5_head(const BR& B, const MW& mw, const MN& mn, int s, int e, LineTok& t, Line5Tail& tail) {
  t.score = 0.f; t.nconv = 0; t.eflag = 1; t.hard = false; t.longname = false; t.qbar = t.sbar = -2;
  const int len = e - s;
  if (len < 23) return false;
  // tabs 1..3: the first three bytes <= 0x20; all three must be tabs (a tab is not in N), no empty field
  const uint64_t N0 = bits64(mn, s);
  uint64_t W0 = bits64(mw, s);
  if (len < 64) W0 &= below64(len);
  if (popc64(W0) < 3) return false;
  const int t0 = ctz64(W0);
  const uint64_t W1 = W0 & (W0 - 1ull);
  const int t1 = ctz64(W1);
  const uint64_t W2 = W1 & (W1 - 1ull);
  const int t2 = ctz64(W2);
  const uint64_t upto_t2 = W2 ^ (W2 - 1ull);             // bits <= t2 (W2's lowest bit is t2)
  if ((W0 & N0 & upto_t2) != 0ull || t0 < 1 || t1 - t0 < 2 || t2 - t1 < 2) return false;
  // field 3 (pctid): digits with at most one '.', at least one digit
  const int lp = t2 - t1 - 1;
  if (lp > 32) return false;
  const uint32_t pb = (uint32_t)(N0 >> (t1 + 1)) & below32(lp);
  if (pb) {
    if ((pb & (pb - 1u)) || lp < 2 || B(s + t1 + 1 + ctz32(pb)) != '.') return false;
  }
  // tabs 10, 11 from the end of the line
  const int a = len > 32 ? len - 32 : 0;
  const uint32_t Nt = bits32(mn, s + a);
  const uint32_t Wt = bits32(mw, s + a) & below32(len - a);
  if (!Wt) return false;
  const int h10 = hibit32(Wt);
  const uint32_t Wt2 = Wt & ~(1u << h10);
  if (!Wt2) return false;
  const int h9 = hibit32(Wt2);
  if (((Nt >> h10) | (Nt >> h9)) & 1u) return false;     // both must be tabs
  const int t10 = a + h10, t9 = a + h9;
  // fields 4..10: between tab 3 and tab 10 exactly 6 more tabs, digits otherwise, no empty field
  const int dm = t9 - t2;
  if (dm < 2 || dm > 63) return false;
  const uint64_t Wm = bits64(mw, s + t2), Nm = bits64(mn, s + t2);
  const uint64_t mid = below64(dm + 1);                  // tab 3 .. tab 10 inclusive
  if (popc64(Wm & mid) != 8 || (Nm & mid) != 0ull || (Wm & (Wm >> 1) & mid) != 0ull) return false;
  tail.le = t10 - t9 - 1; tail.ls = len - t10 - 1;
  if (tail.le > 14 || tail.ls > 14) return false;
  tail.eb = (Nt >> (h9 + 1)) & below32(tail.le);
  tail.sb = (h10 < 31 ? (Nt >> (h10 + 1)) : 0u) & below32(tail.ls);
  tail.pe = s + t9 + 1; tail.ps = s + t10 + 1;
  t.qa = s; t.qb = s + t0; t.sa = s + t0 + 1; t.sb = s + t1;
  return true;
}
// second half: evalue and score
template <class BR>
JX_HD bool fast_line5_tail(const BR& B, const Line5Tail& tail, LineTok& t) {
  if (!evalue5(B, tail.pe, tail.le, tail.eb, t.eflag)) return false;
  if (!score5(B, tail.ps, tail.ls, tail.sb, t.score)) return false;
  t.nconv = 12;
  return true;
}
template <class BR, class MW, class MN>
JX_HD bool fast_line5(const BR& B, const MW& mw, const MN& mn, int s, int e, LineTok& t) {
  Line5Tail tail;
  return fast_line5_head(B, mw, mn, s, e, t, tail) && fast_line5_tail(B, tail, t);
}

// ---- formats.blast filter (src/jcvi/formats/blast.py:620-709): every field the predicate reads -----------------
struct FilterFields {
  float score, pct;             // float32 like the Cython object's C floats (0 when not converted)
  long long hitlen;
  int ev_gt;                    // strtod(evalue) > cutoff
  int qa, qb, sa, sb;           // raw name spans
  bool hard, longname, fast;
};
// [s, e) = the row without its terminator (may be empty: BlastLine of "\n", every field 0).  `try_fast`: the
// row lies inside the window the masks cover.  Fast route = fast_line5_head's structure check + word
// arithmetic; everything else goes through the sscanf-faithful scanner.
template <class BR, class MW, class MN, class R>
JX_HD void filter_parse(const BR& B, const MW& mw, const MN& mn, R& rd, int s, int e, bool try_fast, double cutoff_evalue,
                        FilterFields& f) {
  const int ev0 = (0.0 > cutoff_evalue) ? 1 : 0;
  f.score = 0.f; f.pct = 0.f; f.hitlen = 0; f.ev_gt = ev0; f.qa = f.qb = f.sa = f.sb = s;
  f.hard = false; f.longname = false; f.fast = false;
  if (s < e && try_fast) {
    LineTok t; Line5Tail tail;
    if (fast_line5_head(B, mw, mn, s, e, t, tail)) {
      // pctid = [tab2+1, tab3), hitlen = [tab3+1, tab4): the next two separators after the subject name
      const int p0 = t.sb + 1;
      const uint64_t Wx = bits64(mw, p0);
      const uint64_t Wy = Wx & (Wx - 1ull);
      const int lp = Wx ? ctz64(Wx) : 99;
      const int lh = Wy ? ctz64(Wy) - lp - 1 : 0;          // both separators inside the 64-byte window, else generic
      const uint32_t pb = (uint32_t)bits64(mn, p0) & below32(lp > 32 ? 32 : lp);
      uint32_t em = 0; int ee = 0;
      float pct = 0.f, score = 0.f;
      if (lp <= 14 && lh >= 1 && lh <= 8 && score5(B, p0, lp, pb, pct) && score5(B, tail.ps, tail.ls, tail.sb, score) &&
          decimal5(B, tail.pe, tail.le, tail.eb, em, ee)) {
        Dec d; d.m = em; d.e10 = em ? ee : 0; d.neg = false; d.ok = true; d.dropped = false; d.special = false;
        const int g = dec_gt_double(d, cutoff_evalue);
        if (g >= 0) {
          f.hitlen = (long long)digits_upto8(B.load32(p0 + lp + 1), B.load32(p0 + lp + 5), lh, -1);
          f.pct = pct; f.score = score; f.ev_gt = g; f.fast = true;
          f.qa = t.qa; f.qb = t.qb; f.sa = t.sa; f.sb = t.sb;
          return;
        }
      }
    }
  }
  if (s >= e) return;
  FullTok ft;
  tokenize_line_full(rd, s, e, ft);
  f.qa = ft.qa; f.qb = ft.qb; f.sa = ft.sa; f.sb = ft.sb; f.hitlen = ft.hitlen;
  f.hard = ft.hard; f.longname = ft.longname;
  if (ft.pctid.ok && !dec_to_f32(ft.pctid, f.pct)) f.hard = true;
  if (ft.score.ok && !dec_to_f32(ft.score, f.score)) f.hard = true;
  if (ft.evalue.ok) { const int g = dec_gt_double(ft.evalue, cutoff_evalue); if (g < 0) f.hard = true; else f.ev_gt = g; }
}
// the predicate of blast.py:681-699 (noids: `not (query in ids and subject in ids)`, false without --ids)
template <class R>
JX_HD bool filter_remove(R& rd, const FilterFields& f, double c_score, double c_pctid, long long c_hitlen, bool noids,
                         bool inverse, bool noself) {
  bool remove = ((double)f.score < c_score) || ((double)f.pct < c_pctid) || (f.hitlen < c_hitlen) || f.ev_gt || noids;
  if (inverse) remove = !remove;
  if (noself && (f.qb - f.qa) == (f.sb - f.sa)) {          // c.query == c.subject
    bool same = true;
    for (int i = 0; i < f.qb - f.qa; ++i) same = same && (rd(f.qa + i) == rd(f.sa + i));
    remove = remove || same;
  }
  return remove;
}

// position of the first '|' in [a, b), or -1 (only called for tiles that may hold one)
template <class BR>
JX_HD int find_bar(const BR& B, int a, int b) {
  int bar = -1;
  for (int p = a; p < b && bar < 0; p += 4) {
    const uint32_t f = swar_eq(B.load32(p), 0x7C7C7C7Cu);
    if (f) { const int q = p + first_flag(f); if (q < b) bar = q; }
  }
  return bar;
}

// gene_name with the first '|' already located (bar = its position or -1); cbook.py:308-329
template <class BR>
JX_HD int strip5(const BR& B, int a, int b, int bar) {
  const int end = bar >= 0 ? bar : b;
  if (end - a < 2) return end;
  const bool ev = B(a) == 'e' && B(a + 1) == 'v';
  // rsplit('.', 1) leaves a one-character suffix iff the last '.' sits at end-2
  return (!ev && B(end - 2) == '.' && B(end - 1) != '.') ? end - 2 : end;
}

// ---- name table slots -------------------------------------------------------------------------------
// One 32-byte slot (one DRAM/L2 sector) per BED name: { rank + 1 (0 = empty), name id, the first 24
// bytes of the name, zero padded }.  A probe for a name of <= 23 bytes is ONE memory round trip: the
// slot's six name words are compared with the query's six zero-padded words (the zero byte after
// the name makes the comparison length-exact); longer names continue in the word blob.  The hash
// runs over max(6, ceil(len/4)) zero-padded little-endian words.
#define JX_SLOT_WORDS 6
// The slots form a PERFECT hash table (hash-and-displace, built on the host in jx_bed_load): the
// top bits of the name hash select a bucket, the bucket's 16-bit displacement d re-mixes the hash
// into a slot that no other name uses.  A lookup is one small-table load + one slot load and never
// walks a chain, so all lanes of a warp take the same path.
JX_HD uint32_t slot_index(uint32_t h, uint32_t d, uint32_t cshift) {
  return ((h ^ (d * 0x85EBCA6Bu)) * 0x9E3779B1u) >> cshift;
}
// Name hash (host table build in jx_core.cu and every kernel): h = init; h = h * C + w over the ceil(len / 4)
// zero-padded little-endian words of the name (at least one); finish = (h ^ len) * K.  A chain of integer
// multiply-adds -- they issue on the FMA pipe, the ALU pipe being the tokenizer's limiter -- whose TOP bits are
// well mixed; bucket (h >> bshift) and slot (slot_index) use the top bits only.
JX_HD uint32_t slot_hash_init() { return 0x2F0B4A87u; }
JX_HD uint32_t slot_hash_step(uint32_t h, uint32_t w) { return h * 0x9E3779B1u + w; }
JX_HD uint32_t slot_hash_finish(uint32_t h, uint32_t len) { return (h ^ len) * 0x85EBCA77u; }
template <class R>
JX_HD uint32_t slot_hash_bytes(R& rd, int a, int b) {            // generic (byte reader) form
  uint32_t h = slot_hash_init();
  const int nw = (b - a + 3) >> 2;
  for (int k = 0; k < (nw > 1 ? nw : 1); ++k) {
    uint32_t w = 0;
    for (int j = 0; j < 4; ++j) { const int p = a + 4 * k + j; if (p < b) w |= rd(p) << (8 * j); }
    h = slot_hash_step(h, w);
  }
  return slot_hash_finish(h, (uint32_t)(b - a));
}
// the six zero-padded words of a name of <= 24 bytes, from aligned words
template <class TW>
JX_HD void name_words6(const TW& wr, int a, int len, uint32_t* w) {
  const int i = a >> 2, sh = (a & 3) * 8;
  uint32_t x[JX_SLOT_WORDS + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k <= JX_SLOT_WORDS; ++k) x[k] = wr(i + k);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  const int full = len >> 2;                                       // words entirely inside the name
  const uint32_t last = (1u << (8 * (len & 3))) - 1u;              // mask of the partial word (0 when len % 4 == 0)
  for (int k = 0; k < JX_SLOT_WORDS; ++k) {
    const uint32_t m = k < full ? 0xFFFFFFFFu : (k == full ? last : 0u);
    w[k] = funnel_r(x[k], x[k + 1], sh) & m;
  }
}
JX_HD uint32_t slot_hash6(const uint32_t* w, int len) {          // 1 <= len <= 24, w = the six zero-padded words
  uint32_t h = slot_hash_step(slot_hash_init(), w[0]);
  const int nw = (len + 3) >> 2;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 1; k < JX_SLOT_WORDS; ++k) if (k < nw) h = slot_hash_step(h, w[k]);
  return slot_hash_finish(h, (uint32_t)len);
}

// ---- 64-bit name hash for tables without a BED (formats.blast cscore: the set of names is whatever
// the table holds).  Two 32-bit chains over the same zero-padded words as the slot hash; `seed` lets
// the caller retry in the (2^-64 per pair) event that two different names collide.  Never 0.
JX_HD uint32_t name64_step2(uint32_t h, uint32_t w) { return rotl32(h + w, 7) * 0xC2B2AE35u; }
JX_HD uint64_t name64_finish(uint32_t h1, uint32_t h2, uint32_t len) {
  const uint32_t a = fmix32(h1 ^ (len * 0x85EBCA77u)), b = fmix32(h2 ^ (a + len));
  const uint64_t h = ((uint64_t)a << 32) | b;
  return (h == 0ull || h == ~0ull) ? 1ull : h;          // 0 = no name, ~0 = sort sentinel
}
JX_HD uint64_t name64_words6(const uint32_t* w, int len, uint32_t seed) {
  uint32_t h1 = slot_hash_init() ^ seed, h2 = 0x9747B28Cu + seed * 0x85EBCA6Bu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < JX_SLOT_WORDS; ++k) { h1 = slot_hash_step(h1, w[k]); h2 = name64_step2(h2, w[k]); }
  return name64_finish(h1, h2, (uint32_t)len);
}
template <class R>
JX_HD uint64_t name64_bytes(R& rd, int a, int b, uint32_t seed) {   // generic (byte reader) form, any length
  uint32_t h1 = slot_hash_init() ^ seed, h2 = 0x9747B28Cu + seed * 0x85EBCA6Bu;
  const int nw = (b - a + 3) >> 2;
  for (int k = 0; k < (nw > JX_SLOT_WORDS ? nw : JX_SLOT_WORDS); ++k) {
    uint32_t w = 0;
    for (int j = 0; j < 4; ++j) { const int p = a + 4 * k + j; if (p < b) w |= rd(p) << (8 * j); }
    h1 = slot_hash_step(h1, w); h2 = name64_step2(h2, w);
  }
  return name64_finish(h1, h2, (uint32_t)(b - a));
}

// word hash of [a, b) streaming aligned words (one load + one funnel shift per 4 bytes)
template <class TW>
JX_HD uint64_t fast_hash5(const TW& wr, int a, int b) {
  NameHash h; h.init();
  const int sh = (a & 3) * 8, len = b - a;
  int i = a >> 2;
  uint32_t cur = wr(i);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int rem = len; rem > 0; rem -= 4) {
    const uint32_t nxt = wr(++i);
    const uint32_t w = funnel_r(cur, nxt, sh);
    cur = nxt;
    h.word(rem < 4 ? (w & ((1u << (8 * rem)) - 1u)) : w);
  }
  return h.finish((uint32_t)len);
}

}  // namespace jx

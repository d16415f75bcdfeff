// Host+device text primitives of the BLAST tab12 tokenizer (K1/K2).
//
// Replaces, per line, what the reference does with glibc `sscanf` in
// src/jcvi/formats/cblast.pyx:106-112 (format string cblast.pyx:15) and with
// `gene_name` in src/jcvi/utils/cbook.py:308-329.  Everything here is
// `__host__ __device__` so that the exact same code is unit-tested on the CPU against
// glibc strtof/strtod/sscanf (tests/test_parse_host.py) and runs inside the CUDA kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JX_HD __host__ __device__ __forceinline__
#else
#define JX_HD inline
#endif
// Reconvergence point for code that the kernel runs with ALL 32 lanes of a warp (fast_line4 and
// its helpers): after a data-dependent loop the lanes are brought back together explicitly,
// because independent thread scheduling does not guarantee it.
#if defined(__CUDA_ARCH__)
#define JX_CONVERGE() __syncwarp()
#else
#define JX_CONVERGE() ((void)0)
#endif

namespace jx {

// ---- byte classes ------------------------------------------------------------------
// sscanf's isspace() in the C locale; '\n' and '\r' never occur inside a line because the
// line splitter treats both as terminators (Python universal newlines, formats/base.py
// must_open -> open()).
JX_HD bool is_space(uint32_t c) { return c == ' ' || (c - 9u) <= 4u; }
JX_HD bool is_digit(uint32_t c) { return (c - (uint32_t)'0') <= 9u; }
JX_HD bool is_term(uint32_t c) { return c == '\n' || c == '\r'; }

// ---- decimal scanner -------------------------------------------------------------------
struct Dec {
  uint64_t m;     // up to 19 significant digits
  int32_t e10;    // value = m * 10^e10  (ignoring dropped digits)
  bool neg;
  bool ok;        // a conversion happened (at least one digit)
  bool dropped;   // non-zero digits beyond the 19th were dropped
  bool special;   // inf / nan / hex float / dangling exponent: not supported on device
};

// Scans the longest prefix matching  [+-]? (d+ [. d*]? | . d+) ([eE] [+-]? d+)?  starting at p
// (no leading whitespace), never reading at or beyond `e`.  Returns the position after the
// prefix (p itself when nothing matched).  `rd(pos)` yields the byte at pos.
template <class R>
JX_HD int scan_decimal(R& rd, int p, int e, Dec& d) {
  d.m = 0; d.e10 = 0; d.neg = false; d.ok = false; d.dropped = false; d.special = false;
  int q = p;
  if (q < e) { uint32_t c = rd(q); if (c == '-') { d.neg = true; ++q; } else if (c == '+') { ++q; } }
  if (q < e) {
    uint32_t c = rd(q) | 0x20u;
    if ((c == 'i' || c == 'n') && q + 2 < e) {
      uint32_t c1 = rd(q + 1) | 0x20u, c2 = rd(q + 2) | 0x20u;
      if ((c == 'i' && c1 == 'n' && c2 == 'f') || (c == 'n' && c1 == 'a' && c2 == 'n')) { d.special = true; return p; }
    }
    if (c == '0' && q + 1 < e && (rd(q + 1) | 0x20u) == 'x') { d.special = true; return p; }
  }
  int nd = 0;          // significant digits accumulated
  bool any = false;
  while (q < e) {
    uint32_t c = rd(q);
    if (!is_digit(c)) break;
    any = true;
    uint32_t v = c - '0';
    if (nd < 19) { if (nd > 0 || v) { d.m = d.m * 10u + v; ++nd; } }
    else { ++d.e10; if (v) d.dropped = true; }
    ++q;
  }
  if (q < e && rd(q) == '.') {
    int q2 = q + 1;
    bool anyf = false;
    while (q2 < e) {
      uint32_t c = rd(q2);
      if (!is_digit(c)) break;
      anyf = true;
      uint32_t v = c - '0';
      if (nd < 19) { if (nd > 0 || v) { d.m = d.m * 10u + v; ++nd; } --d.e10; }
      else if (v) d.dropped = true;
      ++q2;
    }
    if (any || anyf) { q = q2; any = true; }
  }
  if (!any) return p;
  d.ok = true;
  if (q < e && (rd(q) | 0x20u) == 'e') {
    int q2 = q + 1;
    bool eneg = false;
    if (q2 < e) { uint32_t c = rd(q2); if (c == '-') { eneg = true; ++q2; } else if (c == '+') { ++q2; } }
    if (q2 < e && is_digit(rd(q2))) {
      int ex = 0;
      while (q2 < e) {
        uint32_t c = rd(q2);
        if (!is_digit(c)) break;
        if (ex < 100000) ex = ex * 10 + (int)(c - '0');
        ++q2;
      }
      d.e10 += eneg ? -ex : ex;
      q = q2;
    } else {
      // "1e", "1e+": glibc scanf treats this as a matching failure while strtod would stop
      // before the 'e'; refuse instead of guessing.
      d.special = true;
    }
  }
  if (d.m == 0) d.e10 = 0;
  return q;
}

// [+-]? d+   (scanf %d).  Returns position after the prefix, p when nothing matched.
template <class R>
JX_HD int scan_int(R& rd, int p, int e) {
  int q = p;
  if (q < e) { uint32_t c = rd(q); if (c == '-' || c == '+') ++q; }
  int q0 = q;
  while (q < e && is_digit(rd(q))) ++q;
  return q == q0 ? p : q;
}

// ---- decimal -> binary ---------------------------------------------------------------
#define JX_POW10D {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11, \
                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22}
#define JX_POW10U {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, \
                   1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, \
                   100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull, \
                   1000000000000000000ull, 10000000000000000000ull}
static const double kPow10dHost[23] = JX_POW10D;
static const uint64_t kPow10uHost[20] = JX_POW10U;
#if defined(__CUDACC__)
static __device__ const double kPow10dDev[23] = JX_POW10D;
static __device__ const uint64_t kPow10uDev[20] = JX_POW10U;
#endif
JX_HD double pow10_exact(int k) {  // 10^k, exact in double for 0 <= k <= 22
#if defined(__CUDA_ARCH__)
  return kPow10dDev[k];
#else
  return kPow10dHost[k];
#endif
}
JX_HD uint64_t pow10_u64(int k) {  // 0 <= k <= 19
#if defined(__CUDA_ARCH__)
  return kPow10uDev[k];
#else
  return kPow10uHost[k];
#endif
}

JX_HD void mul64x64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b; hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b; lo = (uint64_t)p; hi = (uint64_t)(p >> 64);
#endif
}
JX_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
JX_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } x; x.u = u; return x.f;
#endif
}
JX_HD uint64_t double_to_bits(double f) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(f);
#else
  union { uint64_t u; double f; } x; x.f = f; return x.u;
#endif
}
JX_HD double bits_to_double(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  union { uint64_t u; double f; } x; x.u = u; return x.f;
#endif
}
JX_HD uint32_t float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { uint32_t u; float f; } x; x.f = f; return x.u;
#endif
}

JX_HD float pow10f_exact(int k) {   // 10^k exact in float for 0 <= k <= 10
  const uint32_t bits = k == 0 ? 0x3F800000u : k == 1 ? 0x41200000u : k == 2 ? 0x42C80000u : k == 3 ? 0x447A0000u
                      : k == 4 ? 0x461C4000u : k == 5 ? 0x47C35000u : k == 6 ? 0x49742400u : k == 7 ? 0x4B189680u
                      : k == 8 ? 0x4CBEBC20u : k == 9 ? 0x4E6E6B28u : 0x501502F9u;
  return bits_to_float(bits);
}

// exact round-to-nearest-even of the 128-bit integer (hi:lo) to float32
JX_HD float u128_to_f32(uint64_t hi, uint64_t lo) {
  if (!hi && !lo) return 0.f;
  int top = hi ? 127 - clz64(hi) : 63 - clz64(lo);   // index of the top set bit
  if (top < 24) return (float)(uint32_t)lo;          // exact
  int sh = top - 23;                                 // bits to drop
  // mant = (N >> sh), rem = N & ((1<<sh)-1)
  uint64_t mant, rem_hi, rem_lo, half_hi, half_lo;
  if (sh >= 64) {
    mant = hi >> (sh - 64);
    rem_hi = (sh - 64) ? (hi & ((1ull << (sh - 64)) - 1)) : 0; rem_lo = lo;
    half_hi = (sh - 64) ? (1ull << (sh - 65)) : 0; half_lo = (sh - 64) ? 0 : (1ull << 63);
  } else {
    mant = (lo >> sh) | (hi ? (hi << (64 - sh)) : 0);
    rem_hi = 0; rem_lo = lo & ((1ull << sh) - 1);
    half_hi = 0; half_lo = 1ull << (sh - 1);
  }
  mant &= 0xFFFFFFull;
  bool gt = rem_hi > half_hi || (rem_hi == half_hi && rem_lo > half_lo);
  bool eq = rem_hi == half_hi && rem_lo == half_lo;
  uint32_t m = (uint32_t)mant;
  int ex = top;
  if (gt || (eq && (m & 1u))) { ++m; if (m == 0x1000000u) { m >>= 1; ++ex; } }
  if (ex > 127) return bits_to_float(0x7F800000u);
  return bits_to_float(((uint32_t)(ex + 127) << 23) | (m & 0x7FFFFFu));
}

// Correctly rounded float32 (== glibc strtof) for the decimals this path sees; returns
// false ("hard") when the cheap exact routes do not apply -- the caller then reports the line
// instead of guessing.  Routes and why they are exact:
//   e10 >= 0 (<= 19): m*10^e10 as an exact 128-bit integer, one rounding.
//   e10 <  0 (>= -12), m < 2^53: one correctly rounded double division followed by the
//        double->float rounding; a float32 midpoint cannot sit within 2^-53 (relative) of
//        m/10^k unless it equals it, because 5^k * 2^25 < 2^53 for k <= 12.
JX_HD bool dec_to_f32(const Dec& d, float& out) {
  if (d.m == 0) { out = d.neg ? -0.f : 0.f; return !d.dropped; }
  if (d.dropped) return false;
  float r;
  if (d.m < (1u << 24) && d.e10 <= 0 && d.e10 >= -10) {
    // both operands exact in float32: one correctly rounded operation == strtof
    const float fm = (float)(uint32_t)d.m;
    if (d.e10 == 0) r = fm;
    else {
#if defined(__CUDA_ARCH__)
      r = __fdiv_rn(fm, pow10f_exact(-d.e10));
#else
      r = fm / pow10f_exact(-d.e10);
#endif
    }
  } else if (d.e10 >= 0) {
    if (d.e10 > 19) {
      if (d.e10 > 60) { r = bits_to_float(0x7F800000u); }
      else return false;
    } else {
      uint64_t hi, lo;
      mul64x64(d.m, pow10_u64(d.e10), hi, lo);
      r = u128_to_f32(hi, lo);
    }
  } else {
    uint64_t m = d.m; int k = -d.e10;
    if (k > 12 || m >= (1ull << 53)) {           // "54.000000000000000" style: strip zeros first
      while (k > 0 && m % 10u == 0) { m /= 10u; --k; }
    }
    if (k == 0) r = u128_to_f32(0, m);
    else if (k <= 12 && m < (1ull << 53)) r = (float)((double)m / pow10_exact(k));
    else return false;
  }
  out = d.neg ? -r : r;
  return true;
}

// Decides  strtod(text) < 1e-10  (tandem_grouper, blastfilter.py:265-271) exactly.
// Returns 1 / 0, or -1 when the value is within 1e-6 (relative) of the threshold and the
// exact route (m < 2^53, |e10| <= 22: single correctly rounded double op) does not apply.
JX_HD int dec_lt_1e10(const Dec& d) {
  if (d.m == 0) return 1;
  if (d.neg) return 1;
  if (!d.dropped && d.m < 1000000000000000ull) {
    // <= 15 significant digits: the decimal is at least 1e-15 (relative) away from 1e-10 unless it
    // is >= 1e-10, so the comparison of the rounded doubles equals the comparison of the decimals:
    // m * 10^e10 < 10^-10  <=>  m < 10^(-10 - e10)
    const int k = -10 - d.e10;
    if (k <= 0) return 0;
    if (k >= 16) return 1;
    return d.m < pow10_u64(k) ? 1 : 0;
  }
  if (!d.dropped && d.m < (1ull << 53) && d.e10 >= -22 && d.e10 <= 22) {
    double v = d.e10 < 0 ? (double)d.m / pow10_exact(-d.e10) : (double)d.m * pow10_exact(d.e10);
    return v < 1e-10 ? 1 : 0;
  }
  // magnitude route: log10(value) = log10(m) + e10; m has at most 19 digits
  int nd = 1; { uint64_t t = d.m; while (t >= 10u) { t /= 10u; ++nd; } }
  int mag = nd - 1 + d.e10;          // value in [10^mag, 10^(mag+1))
  if (mag <= -12) return 1;
  if (mag >= -9) return 0;
  // mag in {-11, -10}: compare leading digits exactly.  value < 1e-10  <=>  mag == -11 unless
  // the digits are 999.. close to rolling over; value >= 1e-10 when mag == -10.  The rounded
  // double can only cross the threshold when the decimal is within 1 ulp (2^-53) of 1e-10.
  if (mag == -11) {
    // value = 0.d1d2..dn * 1e-10 ; within 1e-6 of threshold iff leading digits are 999999..
    uint64_t lead = d.m; int n = nd; while (n > 7) { lead /= 10u; --n; }
    uint64_t all9 = pow10_u64(n) - 1;
    if (n >= 7 && lead == all9) return -1;
    return 1;
  }
  // mag == -10: value in [1e-10, 1e-9).  RN(1e-10) is the double nearest to 1e-10; any decimal
  // >= 1e-10 rounds to a double >= RN(1e-10), hence not "<".
  return 0;
}

// ---- %.3g round trip -------------------------------------------------------------------------
// strtof(sprintf("%.3g", (double)x)) for finite x: what `synteny scan` sees after
// `blastfilter` printed the score with "%.3g" (cblast.pyx:17) -- exact integer arithmetic,
// round-half-even on the exact binary value like glibc printf.  Returns false when the
// magnitude is outside the exactly handled range (caller falls back to text round trip).
// |x| rounded to three significant digits exactly like glibc's "%.3g": |x| ~ q * 10^(p-2), 100 <= q <= 999
JX_HD bool round3_digits(float x, uint32_t& q_out, int& p_out, bool& neg_out, bool& zero_out) {
  uint32_t u = float_to_bits(x);
  bool neg = u >> 31;
  uint32_t ex = (u >> 23) & 0xFFu, fr = u & 0x7FFFFFu;
  if (ex == 0xFFu) return false;
  neg_out = neg; zero_out = false;
  if (ex == 0 && fr == 0) { zero_out = true; q_out = 0; p_out = 0; return true; }
  if (ex == 0) return false;                 // subnormal: not a score
  const uint64_t M = fr | 0x800000u;         // |x| = M * 2^E, 2^23 <= M < 2^24
  const int E = (int)ex - 150;
  uint64_t q; int p;                         // |x| ~ q * 10^(p-2), 100 <= q <= 999
  uint64_t rem, den;                         // exact remainder / denominator of the cut
  if (E >= 0) {
    if (E > 39) return false;                // >= 2^63
    const uint64_t N = M << E;               // exact integer
    p = 0; while (p < 19 && N >= pow10_u64(p + 1)) ++p;
    if (p <= 2) { q = N * pow10_u64(2 - p); rem = 0; den = 1; }
    else { const uint64_t D = pow10_u64(p - 2); q = N / D; rem = N % D; den = D; }
  } else {
    const int s = -E;                        // |x| = M / 2^s
    if (s > 60) return false;
    const uint64_t I = s < 24 ? (M >> s) : 0;
    if (I >= 1000) {
      p = 3; while (I >= pow10_u64(p + 1)) ++p;
      const uint64_t D = pow10_u64(p - 2) << s;   // <= M, fits
      q = M / D; rem = M % D; den = D;
    } else {
      p = 2;
      for (; p >= -10; --p) {
        bool ge;
        if (p >= 0) ge = (s < 40) && ((pow10_u64(p) << s) <= M);
        else ge = ((M * pow10_u64(-p)) >> s) >= 1;
        if (ge) break;
      }
      if (p < -10) return false;
      const uint64_t N = M * pow10_u64(2 - p);     // < 2^24 * 10^12 < 2^64
      q = N >> s; rem = N & ((1ull << s) - 1); den = 1ull << s;
    }
  }
  // round half to even on the exact value (glibc printf)
  const uint64_t twice = rem << 1;            // rem < den <= 2^61: no overflow
  if (twice > den || (twice == den && (q & 1u))) ++q;
  if (q >= 1000) { q = 100; ++p; }
  q_out = (uint32_t)q; p_out = p;
  return true;
}
JX_HD bool round3g_f32(float x, float& out) {
  uint32_t q; int p; bool neg, zero;
  if (!round3_digits(x, q, p, neg, zero)) return false;
  if (zero) { out = x; return true; }
  Dec d; d.m = q; d.e10 = p - 2; d.neg = neg; d.ok = true; d.dropped = false; d.special = false;
  return dec_to_f32(d, out);
}

// ---- names -------------------------------------------------------------------------------
// gene_name (cbook.py:308-329): cut at the first '|'; unless the name starts with "ev", drop
// a trailing ".x" whose suffix is exactly one character.  Adjusts b.
template <class R>
JX_HD void strip_isoform(R& rd, int a, int& b) {
  bool ev = (b - a >= 2) && rd(a) == 'e' && rd(a + 1) == 'v';
  int dot = -1, end = b;
  for (int p = a; p < b; ++p) {
    uint32_t c = rd(p);
    if (c == '|') { end = p; break; }
    if (c == '.') dot = p;
  }
  if (!ev && dot >= 0 && end - dot - 1 == 1) end = dot;
  b = end;
}

// 64-bit hash of the bytes [a,b), defined on little-endian 32-bit words (last word zero padded)
// so that the kernel can hash straight from shared-memory words.  Host (BED table build) and
// device (probe) must agree, hence JX_HD.
JX_HD uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
JX_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
struct NameHash {
  uint32_t h1, h2;
  JX_HD void init() { h1 = 0x2F0B4A87u; h2 = 0x9747B28Cu; }
  JX_HD void word(uint32_t w) {
    h1 = (h1 ^ w) * 0x9E3779B1u;
    h2 = (h2 + w) * 0x85EBCA77u; h2 ^= h2 >> 15;
  }
  JX_HD uint64_t finish(uint32_t len) {
    uint32_t a = fmix32(h1 ^ len), b = fmix32(h2 ^ (a + len));
    return ((uint64_t)a << 32) | b;
  }
};
template <class R>
JX_HD uint64_t name_hash(R& rd, int a, int b) {
  NameHash h; h.init();
  for (int p = a; p < b; p += 4) {
    uint32_t w = rd(p);
    if (p + 1 < b) w |= rd(p + 1) << 8;
    if (p + 2 < b) w |= rd(p + 2) << 16;
    if (p + 3 < b) w |= rd(p + 3) << 24;
    h.word(w);
  }
  return h.finish((uint32_t)(b - a));
}

// ---- one line --------------------------------------------------------------------------------
struct LineTok {
  int qa, qb, sa, sb;   // raw name spans
  int qbar, sbar;       // position of the first '|' in each name, -1 none, -2 not searched
  float score;
  int nconv;            // number of successful sscanf conversions (0..12)
  int eflag;            // evalue < 1e-10 : 1 / 0
  bool hard;            // a needed numeric field could not be converted exactly on device
  bool longname;        // a name of >= 128 bytes (cblast.pyx char[128] overflow)
};

// sscanf(line, "%s\t%s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f") over [s,e) (e = terminator).
// Fields that are not converted keep the zero initialisation of the Cython object.
template <class R>
JX_HD void tokenize_line(R& rd, int s, int e, LineTok& t) {
  t.qa = t.qb = t.sa = t.sb = s; t.score = 0.f; t.nconv = 0; t.eflag = 1; t.hard = false; t.longname = false;
  t.qbar = t.sbar = -2;
  int p = s;
  while (p < e && is_space(rd(p))) ++p;
  t.qa = p; while (p < e && !is_space(rd(p))) ++p; t.qb = p;
  if (t.qb == t.qa) return;
  t.nconv = 1;
  while (p < e && is_space(rd(p))) ++p;
  t.sa = p; while (p < e && !is_space(rd(p))) ++p; t.sb = p;
  if (t.sb == t.sa) return;
  t.nconv = 2;
  if (t.qb - t.qa >= 128 || t.sb - t.sa >= 128) t.longname = true;
  Dec d;
  while (p < e && is_space(rd(p))) ++p;
  int p2 = scan_decimal(rd, p, e, d);            // %f pctid (value not needed here)
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  p = p2; t.nconv = 3;
  for (int k = 0; k < 7; ++k) {                   // 7 x %d
    while (p < e && is_space(rd(p))) ++p;
    p2 = scan_int(rd, p, e);
    if (p2 == p) return;
    p = p2; ++t.nconv;
  }
  while (p < e && is_space(rd(p))) ++p;
  p2 = scan_decimal(rd, p, e, d);                 // %lf evalue
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  p = p2; t.nconv = 11;
  int lt = dec_lt_1e10(d);
  if (lt < 0) { t.hard = true; return; }
  t.eflag = lt;
  while (p < e && is_space(rd(p))) ++p;
  p2 = scan_decimal(rd, p, e, d);                 // %f score
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  t.nconv = 12;
  if (!dec_to_f32(d, t.score)) t.hard = true;
}


// ---- every field (formats.blast filter) ----------------------------------------------------------
// `jcvi.formats.blast filter` (src/jcvi/formats/blast.py:620-709) tests score, pctid, hitlen and evalue
// of every BlastLine.  Fields sscanf does not reach keep the zero initialisation of the Cython object.
struct FullTok {
  int qa, qb, sa, sb;   // raw name spans (empty when not converted)
  Dec pctid, evalue, score;   // .ok = converted
  long long hitlen;     // value of the %d conversion (0 when not converted)
  int nconv;
  bool hard;            // something the device cannot reproduce exactly (special floats, 10+ digit ints, ...)
  bool longname;
};
// [+-]? d+ with its value; returns the position after it (p when nothing matched); > 9 digits -> *big
template <class R>
JX_HD int scan_int_value(R& rd, int p, int e, long long& val, bool& big) {
  int q = p; bool neg = false;
  if (q < e) { uint32_t c = rd(q); if (c == '-') { neg = true; ++q; } else if (c == '+') { ++q; } }
  const int q0 = q; long long v = 0; int nd = 0;
  while (q < e && is_digit(rd(q))) { if (nd < 18) v = v * 10 + (long long)(rd(q) - '0'); if (nd > 0 || rd(q) != '0') ++nd; ++q; }
  if (q == q0) return p;
  big = nd > 9;
  val = neg ? -v : v;
  return q;
}
template <class R>
JX_HD void tokenize_line_full(R& rd, int s, int e, FullTok& t) {
  t.qa = t.qb = t.sa = t.sb = s; t.nconv = 0; t.hard = false; t.longname = false; t.hitlen = 0;
  t.pctid.ok = t.evalue.ok = t.score.ok = false; t.pctid.m = t.evalue.m = t.score.m = 0;
  t.pctid.special = t.evalue.special = t.score.special = false;
  int p = s;
  while (p < e && is_space(rd(p))) ++p;
  t.qa = p; while (p < e && !is_space(rd(p))) ++p; t.qb = p;
  if (t.qb == t.qa) return;
  t.nconv = 1;
  while (p < e && is_space(rd(p))) ++p;
  t.sa = p; while (p < e && !is_space(rd(p))) ++p; t.sb = p;
  if (t.sb == t.sa) { t.sa = t.sb = s; return; }
  t.nconv = 2;
  if (t.qb - t.qa >= 128 || t.sb - t.sa >= 128) t.longname = true;
  while (p < e && is_space(rd(p))) ++p;
  Dec d;
  int p2 = scan_decimal(rd, p, e, d);            // %f pctid
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  t.pctid = d; p = p2; t.nconv = 3;
  for (int k = 0; k < 7; ++k) {                   // 7 x %d; the first is hitlen
    while (p < e && is_space(rd(p))) ++p;
    long long v = 0; bool big = false;
    p2 = scan_int_value(rd, p, e, v, big);
    if (p2 == p) return;
    if (k == 0) { t.hitlen = v; if (big) t.hard = true; }
    p = p2; ++t.nconv;
  }
  while (p < e && is_space(rd(p))) ++p;
  p2 = scan_decimal(rd, p, e, d);                 // %lf evalue
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  t.evalue = d; p = p2; t.nconv = 11;
  while (p < e && is_space(rd(p))) ++p;
  p2 = scan_decimal(rd, p, e, d);                 // %f score
  if (d.special) { t.hard = true; return; }
  if (!d.ok) return;
  t.score = d; t.nconv = 12;
}

// ---- strtod(text) > c, decided exactly ----------------------------------------------------------------
// (formats.blast filter, blast.py:690: `c.evalue > evalue`, c.evalue = the %lf conversion of the field.)
// strtod rounds the decimal x = m * 10^e10 to nearest-even.  For a finite a >= 0 with a = k * 2^q (k < 2^53 the
// significand as an integer, q >= -1074), the next double is (k + 1) * 2^q, so
//     RN(x) > a   <=>   x > mid  or  (x == mid and k odd),   mid = (2 k + 1) * 2^(q - 1)
// (a = DBL_MAX: RN(x) = inf beyond the same midpoint; a = 0: mid = 2^-1075).  x and mid are compared as
// integers after clearing the power of ten: a fixed-size big integer (1088 bits: 5^364 * 2^64 fits) is
// only built when a bit-length estimate cannot separate them, i.e. for values within ~2 decades of a.
struct Big { uint32_t w[34]; int n; };                      // little-endian words, n = words in use (no leading zeros)
JX_HD void big_set(Big& b, uint64_t v) {
  b.w[0] = (uint32_t)v; b.w[1] = (uint32_t)(v >> 32);
  b.n = b.w[1] ? 2 : (b.w[0] ? 1 : 0);
}
JX_HD bool big_mul_small(Big& b, uint32_t f) {              // false: would not fit
  uint64_t carry = 0;
  for (int i = 0; i < b.n; ++i) { const uint64_t t = (uint64_t)b.w[i] * f + carry; b.w[i] = (uint32_t)t; carry = t >> 32; }
  if (carry) { if (b.n >= 34) return false; b.w[b.n++] = (uint32_t)carry; }
  return true;
}
JX_HD bool big_mul_pow5(Big& b, int n) {                    // b *= 5^n
  while (n >= 13) { if (!big_mul_small(b, 1220703125u)) return false; n -= 13; }   // 5^13 < 2^32
  uint32_t f = 1; while (n-- > 0) f *= 5u;
  return f == 1u || big_mul_small(b, f);
}
JX_HD int big_bitlen(const Big& b) {
  if (!b.n) return 0;
  uint32_t t = b.w[b.n - 1]; int l = 0; while (t) { ++l; t >>= 1; }
  return 32 * (b.n - 1) + l;
}
JX_HD int u64_bitlen(uint64_t v) { int l = 0; while (v) { ++l; v >>= 1; } return l; }
// sign of  A * 2^ea  -  S * 2^es   (A big, S 64-bit, both > 0)
JX_HD int big_cmp_scaled(const Big& A, int ea, uint64_t S, int es) {
  const int la = big_bitlen(A) + ea, ls = u64_bitlen(S) + es;
  if (la != ls) return la > ls ? 1 : -1;
  const int sh = es - ea;                                  // compare A with S * 2^sh (sh >= 0) or A * 2^-sh with S
  if (sh < 0) {                                            // then A < 2^64 (equal top bits): shift it instead
    const uint64_t av = (uint64_t)A.w[0] | ((uint64_t)(A.n > 1 ? A.w[1] : 0u) << 32);
    const uint64_t sa = av << (-sh);
    return sa > S ? 1 : (sa < S ? -1 : 0);
  }
  const int ws = sh >> 5, bs = sh & 31;                    // word i of S << sh
  for (int i = A.n - 1; i >= 0; --i) {
    const int j = i - ws;                                  // bits [32 j - bs, 32 j - bs + 32) of S
    uint32_t sw = 0;
    if (j >= 0 && j <= 2) {
      const uint64_t lo = j == 0 ? (S << bs) : (S >> (32 * j - bs));
      sw = (uint32_t)lo;
      if (j == 0 && bs == 0) sw = (uint32_t)S;
    } else if (j == -1 || j > 2) sw = 0;
    if (A.w[i] != sw) return A.w[i] > sw ? 1 : -1;
  }
  return 0;
}
// sign of  m * 10^e10  -  K * 2^Q   (m, K > 0);  2: the big integer would not fit (never for inputs that
// pass dec_gt_double's estimate)
JX_HD int dec_cmp_dyadic(uint64_t m, int e10, uint64_t K, int Q) {
  Big b;
  if (e10 >= 0) {                                          // m 5^e 2^e  vs  K 2^Q
    big_set(b, m);
    if (!big_mul_pow5(b, e10)) return 2;
    return big_cmp_scaled(b, e10, K, Q);
  }
  big_set(b, K);                                           // m  vs  K 5^n 2^(Q + n)
  if (!big_mul_pow5(b, -e10)) return 2;
  return -big_cmp_scaled(b, Q - e10, m, 0);
}
// RN(m * 10^e10 [+ dropped digits]) > a  for finite a >= 0 and m > 0:  1 / 0 / -1 (not decidable: only with
// dropped digits straddling the midpoint)
JX_HD int dec_gt_nonneg(uint64_t m, int e10, bool dropped, double a) {
  const uint64_t bits = double_to_bits(a);
  const int be = (int)((bits >> 52) & 0x7FFu);
  const uint64_t frac = bits & ((1ull << 52) - 1ull);
  const uint64_t k = be ? (frac | (1ull << 52)) : frac;
  const int q = be ? be - 1075 : -1074;
  const uint64_t K = 2ull * k + 1ull; const int Q = q - 1;
  // estimate: log2(mid) in [bitlen(K) - 1 + Q, bitlen(K) + Q), log10(x) in [nd - 1 + e10, nd + e10 (+ dropped))
  int nd = 1; { uint64_t t = m; while (t >= 10u) { t /= 10u; ++nd; } }
  const double l2lo = (double)(u64_bitlen(K) - 1 + Q), l2hi = l2lo + 1.0;
  const double LOG2_10 = 3.321928094887362;
  if ((double)(nd - 1 + e10) * LOG2_10 > l2hi + 1.0) return 1;
  if ((double)(nd + e10 + (dropped ? 1 : 0)) * LOG2_10 < l2lo - 1.0) return 0;
  const int c0 = dec_cmp_dyadic(m, e10, K, Q);
  if (c0 == 2) return -1;
  if (!dropped) return (c0 > 0 || (c0 == 0 && (k & 1ull))) ? 1 : 0;
  if (c0 >= 0) return 1;                                   // x > m 10^e10 >= mid
  const int c1 = dec_cmp_dyadic(m + 1ull, e10, K, Q);
  if (c1 == 2) return -1;
  if (c1 <= 0) return 0;                                   // x < (m + 1) 10^e10 <= mid
  return -1;
}
JX_HD double next_down_pos(double a) { return bits_to_double(double_to_bits(a) - 1ull); }   // a > 0 finite
// strtod(text) > c  for the scanned decimal d:  1 / 0 / -1 (refused)
JX_HD int dec_gt_double(const Dec& d, double c) {
  if (d.special || !d.ok) return -1;
  if (c != c) return 0;                                    // nan
  const uint64_t cb = double_to_bits(c);
  const bool cneg = (cb >> 63) != 0ull;
  const bool cinf = ((cb >> 52) & 0x7FFull) == 0x7FFull;
  if (cinf) return cneg ? 1 : 0;
  if (d.m == 0 && !d.dropped) return 0.0 > c ? 1 : 0;      // +-0.0 > c
  if (!d.dropped && d.m < (1ull << 53) && d.e10 >= -22 && d.e10 <= 22) {   // one correctly rounded operation
    const double v = d.e10 < 0 ? (double)d.m / pow10_exact(-d.e10) : (double)d.m * pow10_exact(d.e10);
    return (d.neg ? -v : v) > c ? 1 : 0;
  }
  if (d.m == 0) return -1;                                 // only dropped digits (20+ leading zeros): refuse
  const double ac = cneg ? -c : c;
  if (!d.neg) return (cneg && ac != 0.0) ? 1 : dec_gt_nonneg(d.m, d.e10, d.dropped, ac);
  // negative value -RN(x):  > c  <=>  c < 0 and RN(x) < |c|  <=>  c < 0 and not RN(x) > pred(|c|)
  if (!cneg || ac == 0.0) return 0;
  const int g = dec_gt_nonneg(d.m, d.e10, d.dropped, next_down_pos(ac));
  return g < 0 ? -1 : 1 - g;
}

// ---- fast path -------------------------------------------------------------------------------
// Word-at-a-time (SWAR) parse of a REGULAR line: 12+ fields separated by single tabs, names made
// of bytes > 0x20, pctid = digits with at most one '.', seven all-digit integers, evalue/score
// plain decimals.  Returns false for anything else; the caller then runs tokenize_line(), the
// generic sscanf-faithful scanner, so the fast path only has to be exact on what it accepts.
// `W` yields aligned 32-bit little-endian words of the buffer: w(i) = bytes [4i, 4i+4).
template <class W>
struct WordBytes {                       // byte/unaligned views on top of a word reader
  const W& w;
  JX_HD explicit WordBytes(const W& w_) : w(w_) {}
  JX_HD uint32_t operator()(int p) const { return (w(p >> 2) >> ((p & 3) * 8)) & 0xFFu; }
  JX_HD uint32_t load32(int p) const {   // bytes [p, p+4)
    const int i = p >> 2, sh = (p & 3) * 8;
    uint32_t lo = w(i);
    if (!sh) return lo;
    return (lo >> sh) | (w(i + 1) << (32 - sh));
  }
};

JX_HD uint32_t swar_le20(uint32_t x) {   // 0x80 in every byte that is <= 0x20 (and < 0x80)
  uint32_t t = (x & 0x7F7F7F7Fu) + 0x5F5F5F5Fu;      // bit7 set iff low7 >= 0x21
  return ~(t | x) & 0x80808080u;
}
JX_HD uint32_t swar_eq(uint32_t x, uint32_t pat) {   // 0x80 in every byte equal to pat's byte
  uint32_t y = x ^ pat;
  uint32_t t = (y & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
  return ~(t | y) & 0x80808080u;
}
JX_HD uint32_t swar_nondigit(uint32_t x) {           // 0x80 in every byte that is not '0'..'9'
  uint32_t a = x & 0x7F7F7F7Fu;
  uint32_t ge0 = a + 0x50505050u;                    // bit7 iff low7 >= 0x30
  uint32_t gt9 = a + 0x46464646u;                    // bit7 iff low7 >= 0x3A
  return ~(ge0 & ~gt9 & ~x) & 0x80808080u;
}
JX_HD int first_flag(uint32_t f) {                   // byte index (0..3) of the lowest 0x80 flag
  // ALU/FMA only: isolate the lowest flag, move it to bit 0 of its byte, table-by-multiply
  const uint32_t x = (f & (0u - f)) >> 7;            // 1, 2^8, 2^16 or 2^24
  return (int)((x * 0x00010203u) >> 24);
}
JX_HD int count_flags4(uint32_t f) {                 // number of 0x80 flags in the word (0..4)
  return (int)((((f >> 7) & 0x01010101u) * 0x01010101u) >> 24);
}

// end of the name starting at p: first byte <= 0x20 must be a tab.  Returns the tab position or
// -1 (irregular / beyond limit).  *bar receives the position of the first '|' (or -1).
template <class W>
JX_HD int fast_name_end(const WordBytes<W>& B, int p, int limit, int* bar) {
  *bar = -1;
  for (int q = p; q < limit; q += 4) {
    const uint32_t x = B.load32(q);
    if (*bar < 0) { uint32_t f = swar_eq(x, 0x7C7C7C7Cu); if (f) *bar = q + first_flag(f); }
    const uint32_t lo = swar_le20(x);
    if (lo) {
      const int k = first_flag(lo), t = q + k;
      if (((x >> (8 * k)) & 0xFFu) != '\t' || t >= limit) return -1;
      if (*bar >= t) *bar = -1;
      return t;
    }
  }
  return -1;
}

// The fast path in two halves so that the kernel can start the two table probes (long-latency
// loads) as soon as the names are known and overlap them with the validation of the numeric
// fields:  fast_line_names  ->  [hash, issue probes]  ->  fast_line_rest.
template <class W>
JX_HD bool fast_line_names(const W& wr, int s, int e, LineTok& t) {
  const WordBytes<W> B(wr);
  t.score = 0.f; t.nconv = 0; t.eflag = 1; t.hard = false; t.longname = false;
  if (e - s < 23) return false;
  const int t1 = fast_name_end(B, s, e, &t.qbar);
  if (t1 <= s || t1 - s >= 128) return false;
  t.qa = s; t.qb = t1;
  const int t2 = fast_name_end(B, t1 + 1, e, &t.sbar);
  if (t2 <= t1 + 1 || t2 - t1 - 1 >= 128) return false;
  t.sa = t1 + 1; t.sb = t2;
  return true;
}
template <class W>
JX_HD bool fast_line_rest(const W& wr, int e, LineTok& t) {
  const WordBytes<W> B(wr);
  const int t2 = t.sb;
  // fields 3..10: non-digit bytes must be exactly: at most one '.' inside field 3, then 8 tabs,
  // no empty field
  int p = t2 + 1;               // start of the current field
  int ntab = 0, ndot = 0, prev = t2, t10 = -1;
  {
    int q = p & ~3;
    uint32_t nd = swar_nondigit(wr(q >> 2));
    nd &= 0xFFFFFFFFu << ((p & 3) * 8);     // ignore bytes before p
    for (;;) {
      while (!nd) { q += 4; if (q >= e) return false; nd = swar_nondigit(wr(q >> 2)); }
      const int pos = q + first_flag(nd);
      nd &= nd - 1;
      if (pos >= e) return false;
      const uint32_t c = B(pos);
      if (c == '\t') {
        const int ndig = pos - prev - 1 - (ntab == 0 ? ndot : 0);
        if (ndig < 1) return false;          // empty field (or a lone '.')
        prev = pos;
        if (++ntab == 8) { t10 = pos; break; }
      } else if (c == '.' && ntab == 0 && ndot == 0) {
        ndot = 1;
      } else {
        return false;
      }
    }
  }
  // field 11 (evalue) and field 12 (score): short decimals, scanned bytewise
  Dec d;
  p = t10 + 1;
  int p2 = scan_decimal(B, p, e, d);
  if (d.special || !d.ok || p2 >= e || B(p2) != '\t') return false;
  const int lt = dec_lt_1e10(d);
  if (lt < 0) return false;
  t.eflag = lt;
  p = p2 + 1;
  p2 = scan_decimal(B, p, e, d);
  if (d.special || !d.ok) return false;
  // whatever follows the score's numeric prefix is ignored by sscanf (last conversion)
  if (!dec_to_f32(d, t.score)) return false;
  t.nconv = 12;
  return true;
}
template <class W>
JX_HD bool fast_line(const W& wr, int s, int e, LineTok& t) {
  return fast_line_names(wr, s, e, t) && fast_line_rest(wr, e, t);
}

// gene_name with the first '|' already located (bar = its position or -1)
template <class W>
JX_HD int fast_strip(const WordBytes<W>& B, int a, int b, int bar) {
  const bool ev = (b - a >= 2) && B(a) == 'e' && B(a + 1) == 'v';
  int end = bar >= 0 ? bar : b;
  // rsplit('.', 1) leaves a one-character suffix iff the last '.' sits at end-2
  if (!ev && end - a >= 2 && B(end - 2) == '.' && B(end - 1) != '.') end -= 2;
  return end;
}

template <class W>
JX_HD uint64_t fast_hash(const WordBytes<W>& B, int a, int b) {
  NameHash h; h.init();
  for (int p = a; p < b; p += 4) {
    uint32_t w = B.load32(p);
    const int rem = b - p;
    if (rem < 4) w &= (1u << (8 * rem)) - 1u;
    h.word(w);
  }
  return h.finish((uint32_t)(b - a));
}


// ---- fast path, version 3 ---------------------------------------------------------------------
// One pass over the line's words finds the bytes <= 0x20 (separators): a regular line has exactly
// 11 tabs followed by the terminator (or a 12th tab / CR LF).  Everything else is validated with
// population counts of SWAR flags, so the instruction stream is nearly identical for all lanes.
// `sep` is a 12-entry scratch array addressed as sep[j * stride].

JX_HD uint32_t range_mask(int q, int a, int b) {    // 0x80 flags of the bytes of word [q,q+4) inside [a,b)
  int lo = a - q; if (lo < 0) lo = 0;
  int hi = b - q; if (hi > 4) hi = 4;
  if (hi <= lo) return 0u;
  uint32_t m = 0xFFFFFFFFu << (8 * lo);
  if (hi < 4) m &= 0xFFFFFFFFu >> (8 * (4 - hi));
  return m & 0x80808080u;
}
JX_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
// number of flagged bytes in [a,b), b > a: edge words masked once, interior words unmasked
template <class W, class F>
JX_HD int count_flags(const W& wr, int a, int b, F fl) {
  const int q0 = a >> 2, q1 = (b - 1) >> 2;
  const uint32_t first = 0x80808080u & (0xFFFFFFFFu << (8 * (a & 3)));
  const uint32_t last = 0x80808080u & (0xFFFFFFFFu >> (8 * (3 - ((b - 1) & 3))));
  if (q0 == q1) return popc32(fl(wr(q0)) & first & last);
  int c = popc32(fl(wr(q0)) & first) + popc32(fl(wr(q1)) & last);
  for (int q = q0 + 1; q < q1; ++q) c += popc32(fl(wr(q)));
  return c;
}
struct FlNonDigit { JX_HD uint32_t operator()(uint32_t x) const { return swar_nondigit(x); } };
struct FlDot { JX_HD uint32_t operator()(uint32_t x) const { return swar_eq(x, 0x2E2E2E2Eu); } };
struct FlCR { JX_HD uint32_t operator()(uint32_t x) const { return swar_eq(x, 0x0D0D0D0Du); } };
template <class W>
JX_HD int count_nondigit(const W& wr, int a, int b) { return b > a ? count_flags(wr, a, b, FlNonDigit()) : 0; }
template <class W>
JX_HD int count_eq(const W& wr, int a, int b, uint32_t) { return b > a ? count_flags(wr, a, b, FlDot()) : 0; }

// evalue field [a,b): decides (value < 1e-10) for the shapes BLAST/LAST print; false = not handled.
// One uniform loop over the characters with a 4-state machine (lanes stay converged).
template <class BR>
JX_HD bool fast_evalue(const BR& B, int a, int b, int& flag) {
  if (b - a > 14 || b <= a) return false;
  const uint32_t first = B(a);
  if (!is_digit(first)) return false;
  int st = 0, nint = 0, nfrac = 0, X = 0, nexp = 0;
  bool nz = false, neg = false, ok = true;
  for (int p = a; p < b; ++p) {
    const uint32_t c = B(p);
    const uint32_t d = c - '0';
    const bool isd = d <= 9u, ise = (c | 0x20u) == 'e';
    if (st == 0) {
      if (isd) { ++nint; nz |= (d != 0); } else if (c == '.') st = 1; else if (ise) st = 2; else ok = false;
    } else if (st == 1) {
      if (isd) { ++nfrac; nz |= (d != 0); } else if (ise) st = 2; else ok = false;
    } else if (st == 2) {
      if (isd) { X = (int)d; nexp = 1; } else if (c == '-' || c == '+') neg = (c == '-'); else ok = false;
      st = 3;
    } else {
      if (isd) { X = X * 10 + (int)d; ++nexp; } else ok = false;
    }
  }
  const bool has_exp = st >= 2;
  if (!ok || (has_exp && (nexp == 0 || nexp > 3))) return false;
  if (neg) X = -X;
  if (!nz) { flag = 1; return true; }                    // 0, 0.0, 0e-5 ...
  if (nint + nfrac > 8) return false;
  if (first == '0') {                                    // "0.ddd": >= 10^-nfrac >= 1e-8
    if (nint == 1 && !has_exp) { flag = 0; return true; }
    return false;
  }
  // value in [10^(nint-1+X), 10^(nint+X)); with <= 8 digits it stays >= 1e-8 (relative) away
  // from 1e-10 unless it is >= 1e-10 exactly, so the rounded double compares the same way
  flag = (nint + X <= -10) ? 1 : 0;
  return true;
}

// score field starting at a (field ends before b): digits [. digits], <= 9 digits, no exponent;
// anything after the numeric prefix is ignored (last sscanf conversion)
template <class BR>
JX_HD bool fast_score(const BR& B, int a, int b, float& score) {
  int st = 0, cnt = 0, k = 0;
  uint32_t m = 0;
  bool expo = false;
  const int bb = b - a > 12 ? a + 12 : b;               // longer fields go to the generic scanner
  if (bb != b) return false;
  for (int p = a; p < b; ++p) {
    const uint32_t c = B(p);
    const uint32_t d = c - '0';
    const bool isd = d <= 9u;
    if (st == 0) {
      if (isd) { m = m * 10u + d; ++cnt; } else if (c == '.') st = 1; else { expo = (c | 0x20u) == 'e'; st = 2; }
    } else if (st == 1) {
      if (isd) { m = m * 10u + d; ++cnt; ++k; } else { expo = (c | 0x20u) == 'e'; st = 2; }
    }
  }
  if (cnt == 0 || cnt > 9 || expo) return false;
  if (m >= (1u << 24)) return false;
  // both operands exact in float -> one correctly rounded division == strtof
#if defined(__CUDA_ARCH__)
  score = k ? __fdiv_rn((float)m, pow10f_exact(k)) : (float)m;
#else
  score = k ? (float)m / pow10f_exact(k) : (float)m;
#endif
  return true;
}

template <class W, class S>
JX_HD bool fast_line3(const W& wr, int s, int limit, S* sep, int stride, LineTok& t) {
  const WordBytes<W> B(wr);
  t.score = 0.f; t.nconv = 0; t.eflag = 1; t.hard = false; t.longname = false; t.qbar = t.sbar = -2;
  // ---- separators: all lanes advance one word per iteration ------------------------------------
  int nsep = 0, end12 = -1;
  {
    uint32_t keep = 0xFFFFFFFFu << (8 * (s & 3));
    bool bad = false;
    for (int q = s & ~3; end12 < 0 && !bad; q += 4) {
      if (q + 8 > limit) { bad = true; break; }
      const uint32_t x = wr(q >> 2);
      uint32_t f = swar_le20(x) & keep;
      keep = 0xFFFFFFFFu;
      while (f) {
        const int j = first_flag(f);
        f &= f - 1;
        const uint32_t c = (x >> (8 * j)) & 0xFFu;
        const int pos = q + j;
        if (c == '\t' && nsep < 11) {
          sep[nsep * stride] = (S)pos;
          ++nsep;
        } else {
          if (c == '\t' || c == '\n' || (c == '\r' && B(pos + 1) == '\n')) end12 = pos; else bad = true;
          break;
        }
      }
    }
    if (bad) return false;
  }
  if (nsep != 11) return false;
  int prev = s - 1;
  bool ok = true;
  for (int j = 0; j < 11; ++j) { const int tj = sep[j * stride]; ok &= (tj - prev >= 2); prev = tj; }
  ok &= (end12 - prev >= 2);
  const int t1 = sep[0], t2 = sep[stride], t3 = sep[2 * stride], t10 = sep[9 * stride], t11 = sep[10 * stride];
  ok &= (t1 - s < 128) & (t2 - t1 - 1 < 128);
  if (!ok) return false;
  t.qa = s; t.qb = t1; t.sa = t1 + 1; t.sb = t2;
  // ---- pctid: digits with at most one '.' ---------------------------------------------------------
  {
    const int nnd = count_nondigit(wr, t2 + 1, t3);
    if (nnd > 1) return false;
    if (nnd == 1 && (count_eq(wr, t2 + 1, t3, 0x2E2E2E2Eu) != 1 || t3 - t2 - 1 < 2)) return false;
  }
  // ---- seven integers: the only non-digits between t3 and t10 are the six tabs ------------------
  if (count_nondigit(wr, t3 + 1, t10) != 6) return false;
  // ---- evalue, score ----------------------------------------------------------------------------------
  int flag;
  if (!fast_evalue(B, t10 + 1, t11, flag)) {
    Dec d;
    const int p2 = scan_decimal(B, t10 + 1, t11, d);
    if (d.special || !d.ok || p2 != t11) return false;
    flag = dec_lt_1e10(d);
    if (flag < 0) return false;
  }
  t.eflag = flag;
  if (!fast_score(B, t11 + 1, end12, t.score)) {
    Dec d;
    scan_decimal(B, t11 + 1, end12, d);
    if (d.special || !d.ok) return false;
    if (!dec_to_f32(d, t.score)) return false;
  }
  t.nconv = 12;
  return true;
}

// gene_name + hash in one go: returns the stripped end, sets h.  '|' is searched while hashing;
// when one turns up the (rare) cut is applied and the hash recomputed.
template <class W>
JX_HD int fast_strip_hash(const WordBytes<W>& B, int a, int b, bool strip, uint64_t& h) {
  int end = b;
  if (strip) {
    const bool ev = (b - a >= 2) && B(a) == 'e' && B(a + 1) == 'v';
    if (!ev && b - a >= 2 && B(b - 2) == '.' && B(b - 1) != '.') end = b - 2;
  }
  for (int pass = 0; pass < 2; ++pass) {
    NameHash nh; nh.init();
    int bar = -1;
    for (int p = a; p < end; p += 4) {
      uint32_t w = B.load32(p);
      const int rem = end - p;
      if (rem < 4) w &= (1u << (8 * rem)) - 1u;
      if (strip && bar < 0) { const uint32_t f = swar_eq(w, 0x7C7C7C7Cu); if (f) bar = p + first_flag(f); }
      nh.word(w);
    }
    if (bar < 0 || pass == 1) { h = nh.finish((uint32_t)(end - a)); return end; }
    // a '|' inside: gene_name cuts there first, then looks for the one-character suffix
    end = bar;
    const bool ev = (b - a >= 2) && B(a) == 'e' && B(a + 1) == 'v';
    if (!ev && end - a >= 2 && B(end - 2) == '.' && B(end - 1) != '.') end -= 2;
    strip = false;   // second pass: plain hash of [a, end)
  }
  return end;
}


// ---- fast path, version 4 ---------------------------------------------------------------------
// Five events instead of twelve: the two name ends and the end of pctid are found walking
// forward, the last two tabs walking backward from the end of the line; the seven integer fields
// in between are validated by counting (every non-digit byte must be a tab, six of them, never
// adjacent), which is one uniform loop over ~8 words.
JX_HD int last_flag(uint32_t f) {                    // byte index (0..3) of the highest 0x80 flag
  return (f & 0x80000000u) ? 3 : (f & 0x00800000u) ? 2 : (f & 0x00008000u) ? 1 : 0;
}
// NOTE on control flow: every loop below has a single exit and the function has a single return.
// With independent thread scheduling a `return`/`break` out of a loop leaves the lanes of a warp
// diverged for the rest of the line (measured: 13 active lanes instead of 25), so validity is
// accumulated in `ok` and all reads are clamped to the line instead of bailing out early.

// last tab in [lo, hi), -1 if none
template <class W>
JX_HD int prev_tab(const W& wr, int hi, int lo) {
  int res = -2;
  uint32_t keep = 0x80808080u & (0xFFFFFFFFu >> (8 * (3 - ((hi - 1) & 3))));
  for (int q = (hi - 1) & ~3; q + 4 > lo && res == -2; q -= 4) {
    const uint32_t f = swar_eq(wr(q >> 2), 0x09090909u) & keep;
    keep = 0x80808080u;
    if (f) { const int pos = q + last_flag(f); res = pos >= lo ? pos : -1; }
  }
  JX_CONVERGE();
  return res == -2 ? -1 : res;
}
// first byte <= 0x20 in [lo, hi): position, or hi when none
template <class W>
JX_HD int next_low(const W& wr, int lo, int hi) {
  int res = -1;
  uint32_t keep = 0x80808080u & (0xFFFFFFFFu << (8 * (lo & 3)));
  for (int q = lo & ~3; q < hi && res < 0; q += 4) {
    const uint32_t f = swar_le20(wr(q >> 2)) & keep;
    keep = 0x80808080u;
    if (f) { const int pos = q + first_flag(f); res = pos < hi ? pos : hi; }
  }
  JX_CONVERGE();
  return res < 0 ? hi : res;
}

#if defined(__CUDACC__)
#define JX_COLD __host__ __device__ __noinline__
#else
#define JX_COLD
#endif
// rare: the uniform state machines declined; generic scanner, kept out of line
template <class BR>
JX_COLD bool slow_evalue(const BR& B, int ea, int eb, int& flag) {
  Dec d;
  const int p2 = scan_decimal(B, ea, eb, d);
  const int lt = dec_lt_1e10(d);
  flag = lt;
  return !d.special && d.ok && p2 == eb && lt >= 0;
}
template <class BR>
JX_COLD bool slow_score(const BR& B, int sa, int e1, float& score) {
  Dec d;
  scan_decimal(B, sa, e1, d);
  return !d.special && d.ok && dec_to_f32(d, score);
}

// s = first byte of the line, e = position of its '\n' (e > s).  All words up to (e + 4) must be
// readable.
template <class W>
JX_HD bool fast_line4(const W& wr, int s, int e, LineTok& t) {
  const WordBytes<W> B(wr);
  t.score = 0.f; t.nconv = 0; t.eflag = 1; t.hard = false; t.longname = false; t.qbar = t.sbar = -2;
  if (e > s && B(e - 1) == '\r') --e;                // CR LF: the CR is trailing whitespace
  bool ok = e - s >= 23;
  const int e1 = ok ? e : s + 1;                     // keep every position inside the line when !ok
  // ---- forward: the two names and pctid ---------------------------------------------------------
  const int t1 = next_low(wr, s, e1);
  ok &= (t1 > s) & (t1 < e1) & (t1 - s < 128);
  const int a2 = t1 + 1 < e1 ? t1 + 1 : e1 - 1;
  const int t2 = next_low(wr, a2, e1);
  ok &= (t2 > a2) & (t2 < e1) & (t2 - a2 < 128);
  const int a3 = t2 + 1 < e1 ? t2 + 1 : e1 - 1;
  const int t3 = next_low(wr, a3, e1);
  ok &= (t3 > a3) & (t3 < e1);
  ok &= (B(t1 < e1 ? t1 : s) == '\t') & (B(t2 < e1 ? t2 : s) == '\t') & (B(t3 < e1 ? t3 : s) == '\t');
  // ---- backward: score and evalue ------------------------------------------------------------------
  const int lo = t3 + 1 < e1 ? t3 + 1 : e1 - 1;
  int t11 = prev_tab(wr, e1, lo);
  ok &= t11 >= 0;
  if (t11 < 0) t11 = e1 - 1;
  int t10 = prev_tab(wr, t11 > lo ? t11 : lo + 1, lo);
  ok &= t10 >= 0;
  if (t10 < 0) t10 = lo;
  ok &= (t11 - t10 >= 2) & (e1 - t11 >= 2);
  // ---- pctid [a3, t3): digits with at most one '.' ------------------------------------------------
  {
    const int a = a3, b = t3 > a3 ? t3 : a3 + 1;
    const int q0 = a >> 2, q1 = (b - 1) >> 2;
    const uint32_t first = 0x80808080u & (0xFFFFFFFFu << (8 * (a & 3)));
    const uint32_t last = 0x80808080u & (0xFFFFFFFFu >> (8 * (3 - ((b - 1) & 3))));
    int nnd = 0; uint32_t bad = 0;
    for (int q = q0; q <= q1; ++q) {
      const uint32_t x = wr(q);
      uint32_t m = 0x80808080u;
      if (q == q0) m &= first;
      if (q == q1) m &= last;
      const uint32_t nd = swar_nondigit(x) & m;
      bad |= nd ^ (swar_eq(x, 0x2E2E2E2Eu) & m);      // a non-digit that is not '.'
      nnd += count_flags4(nd);
    }
    JX_CONVERGE();
    ok &= (bad == 0) & (nnd <= 1) & (b - a - nnd >= 1);
  }
  // ---- the seven integers [t3+1, t10): non-digits == tabs, six of them, none adjacent ----------------
  {
    const int a = lo, b = t10 > lo ? t10 : lo + 1;
    ok &= b - a >= 13;
    const int q0 = a >> 2, q1 = (b - 1) >> 2;
    const uint32_t first = 0x80808080u & (0xFFFFFFFFu << (8 * (a & 3)));
    const uint32_t last = 0x80808080u & (0xFFFFFFFFu >> (8 * (3 - ((b - 1) & 3))));
    int ntab = 0; uint32_t bad = 0, carry = 0;
    // the byte before `a` is the tab t3: as the "previous byte" of a it sits in the carry (a at a
    // word start) or one byte below a inside the first word
    const uint32_t prevflag = (a & 3) ? (0x80u << (8 * ((a & 3) - 1))) : 0u;
    carry = (a & 3) ? 0u : 0x80u;
    for (int q = q0; q <= q1; ++q) {
      const uint32_t x = wr(q);
      uint32_t m = 0x80808080u;
      if (q == q0) m &= first;
      if (q == q1) m &= last;
      const uint32_t nd = swar_nondigit(x) & m;
      const uint32_t tb = swar_eq(x, 0x09090909u) & m;
      const uint32_t tbp = (q == q0) ? (tb | prevflag) : tb;
      bad |= nd ^ tb;                                 // a non-digit that is not a tab
      bad |= ((tbp << 8) & tb) | (tb & carry);        // adjacent tabs inside / across words
      carry = tb >> 24;
      ntab += count_flags4(tb);
    }
    JX_CONVERGE();
    ok &= (bad == 0) & (ntab == 6) & (B(b - 1) != '\t');
  }
  t.qa = s; t.qb = t1; t.sa = a2; t.sb = t2;
  // ---- evalue, score (uniform state machines; the generic scanner only when they decline) ------------
  {
    const int ea = t10 + 1, eb = t11 > t10 + 1 ? t11 : t10 + 2;
    int flag = 1;
    bool done = fast_evalue(B, ea, eb, flag);
    JX_CONVERGE();
    if (ok && !done) done = slow_evalue(B, ea, eb, flag);
    ok &= done;
    t.eflag = flag;
    const int sa = t11 + 1 < e1 ? t11 + 1 : e1 - 1;
    JX_CONVERGE();
    done = fast_score(B, sa, e1, t.score);
    JX_CONVERGE();
    if (ok && !done) done = slow_score(B, sa, e1, t.score);
    ok &= done;
    // junk after the score is ignored by sscanf, but a '\r' in it would be a line break for Python
    JX_CONVERGE();
    ok &= count_flags(wr, sa, e1, FlCR()) == 0;
    JX_CONVERGE();
  }
  t.nconv = ok ? 12 : 0;
  return ok;
}

// A name of <= 16 bytes held in four registers: strip, hash and (later) verify without touching
// shared memory again.  Longer names use fast_strip_hash + the generic probe.
struct NameRegs { uint32_t w[4]; int len; };

template <class W>
JX_HD bool name_to_regs(const WordBytes<W>& B, int a, int b, bool strip, NameRegs& n, uint64_t& h) {
  if (b - a > 16 || b <= a) return false;
  uint32_t w0 = B.load32(a), w1 = 0, w2 = 0, w3 = 0;
  int len = b - a;
  if (len > 4) w1 = B.load32(a + 4);
  if (len > 8) w2 = B.load32(a + 8);
  if (len > 12) w3 = B.load32(a + 12);
  if (strip) {
    // gene_name: cut at the first '|', then drop ".x" unless the name starts with "ev"
    const uint32_t f0 = swar_eq(w0, 0x7C7C7C7Cu), f1 = swar_eq(w1, 0x7C7C7C7Cu), f2 = swar_eq(w2, 0x7C7C7C7Cu),
                   f3 = swar_eq(w3, 0x7C7C7C7Cu);
    int bar = f0 ? first_flag(f0) : f1 ? 4 + first_flag(f1) : f2 ? 8 + first_flag(f2) : f3 ? 12 + first_flag(f3) : 99;
    if (bar < len) len = bar;
    const bool ev = (b - a >= 2) && (w0 & 0xFFFFu) == 0x7665u;    // "ev"
    if (!ev && len >= 2) {
      const int i2 = len - 2;                                    // bytes len-2, len-1
      const uint32_t wa = i2 < 4 ? w0 : i2 < 8 ? w1 : i2 < 12 ? w2 : w3;
      const uint32_t wb = (i2 + 1) < 4 ? w0 : (i2 + 1) < 8 ? w1 : (i2 + 1) < 12 ? w2 : w3;
      const uint32_t c2 = (wa >> (8 * (i2 & 3))) & 0xFFu, c1 = (wb >> (8 * ((i2 + 1) & 3))) & 0xFFu;
      if (c2 == '.' && c1 != '.') len -= 2;
    }
  }
  // zero the bytes beyond len
  const uint32_t k0 = len >= 4 ? 0xFFFFFFFFu : (len <= 0 ? 0u : (1u << (8 * len)) - 1u);
  const uint32_t k1 = len >= 8 ? 0xFFFFFFFFu : (len <= 4 ? 0u : (1u << (8 * (len - 4))) - 1u);
  const uint32_t k2 = len >= 12 ? 0xFFFFFFFFu : (len <= 8 ? 0u : (1u << (8 * (len - 8))) - 1u);
  const uint32_t k3 = len >= 16 ? 0xFFFFFFFFu : (len <= 12 ? 0u : (1u << (8 * (len - 12))) - 1u);
  n.w[0] = w0 & k0; n.w[1] = w1 & k1; n.w[2] = w2 & k2; n.w[3] = w3 & k3; n.len = len;
  NameHash nh; nh.init();
  if (len > 0) nh.word(n.w[0]);
  if (len > 4) nh.word(n.w[1]);
  if (len > 8) nh.word(n.w[2]);
  if (len > 12) nh.word(n.w[3]);
  h = nh.finish((uint32_t)len);
  return true;
}

}  // namespace jx

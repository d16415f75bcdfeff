// Host+device exact re-implementation of the three printf conversions of cblast.pyx:17
//     "%s\t%s\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.3g"
// (BlastLine.__str__, cblast.pyx:144-156) for the values the tokenizer parsed.  Every routine
// returns false when the input is outside what it reproduces exactly; the caller then formats
// the whole output with the host's glibc instead (never a silent difference).
// Unit-tested on the CPU against glibc snprintf (tests/test_format_host.py).
#pragma once
#include "jx_parse.cuh"

namespace jx {

JX_HD int fmt_uint(char* o, uint64_t v) {
  char tmp[24]; int n = 0;
  do { tmp[n++] = (char)('0' + (int)(v % 10u)); v /= 10u; } while (v);
  for (int i = 0; i < n; ++i) o[i] = tmp[n - 1 - i];
  return n;
}
JX_HD int fmt_int(char* o, long long v) {
  if (v < 0) { o[0] = '-'; return 1 + fmt_uint(o + 1, (uint64_t)(-v)); }
  return fmt_uint(o, (uint64_t)v);
}

// "%.2f" of a float32 (promoted to double by printf: exact), round-half-even on the exact value
JX_HD bool fmt_f32_2f(char* o, int& n, float x) {
  const uint32_t u = float_to_bits(x);
  const bool neg = u >> 31;
  const uint32_t ex = (u >> 23) & 0xFFu, fr = u & 0x7FFFFFu;
  if (ex == 0xFFu) return false;
  uint64_t N = 0;                                  // round(|x| * 100)
  if (ex != 0) {
    const uint64_t M = (uint64_t)(fr | 0x800000u) * 100u;      // < 2^31
    const int E = (int)ex - 150;
    if (E >= 0) { if (E > 32) return false; N = M << E; }
    else {
      const int s = -E;
      if (s >= 64) N = 0;
      else {
        N = M >> s;
        const uint64_t rem = M & ((1ull << s) - 1), half = 1ull << (s - 1);
        if (rem > half || (rem == half && (N & 1u))) ++N;
      }
    }
  } else if (fr != 0) {
    N = 0;                                         // subnormal: far below 0.005
  }
  n = 0;
  if (neg) o[n++] = '-';
  n += fmt_uint(o + n, N / 100u);
  o[n++] = '.';
  o[n++] = (char)('0' + (int)((N / 10u) % 10u));
  o[n++] = (char)('0' + (int)(N % 10u));
  return true;
}

// tail shared by %.2g / %.3g: `digits` has exactly `nd` decimal digits (first one non-zero) and
// the value is digits * 10^(P - nd + 1), P = decimal exponent of the rounded value
JX_HD int fmt_g_tail(char* o, bool neg, uint32_t digits, int nd, int P) {
  int n = 0;
  if (neg) o[n++] = '-';
  char d[4];
  for (int i = nd - 1; i >= 0; --i) { d[i] = (char)('0' + (int)(digits % 10u)); digits /= 10u; }
  int last = nd;                                   // strip trailing zeros (no '#' flag)
  while (last > 1 && d[last - 1] == '0') --last;
  if (P < -4 || P >= nd) {                         // exponential style
    o[n++] = d[0];
    if (last > 1) { o[n++] = '.'; for (int i = 1; i < last; ++i) o[n++] = d[i]; }
    o[n++] = 'e';
    int e = P;
    if (e < 0) { o[n++] = '-'; e = -e; } else o[n++] = '+';
    if (e < 10) o[n++] = '0';
    n += fmt_uint(o + n, (uint64_t)e);
    return n;
  }
  if (P >= 0) {                                    // d[0..P] before the point
    for (int i = 0; i <= P; ++i) o[n++] = d[i];    // P < nd, digits exist
    if (last > P + 1) { o[n++] = '.'; for (int i = P + 1; i < last; ++i) o[n++] = d[i]; }
  } else {
    o[n++] = '0'; o[n++] = '.';
    for (int i = 0; i < -P - 1; ++i) o[n++] = '0';
    for (int i = 0; i < last; ++i) o[n++] = d[i];
  }
  return n;
}

// "%.3g" of a float32
JX_HD bool fmt_f32_3g(char* o, int& n, float x) {
  uint32_t q; int p; bool neg, zero;
  if (!round3_digits(x, q, p, neg, zero)) return false;
  if (zero) { n = 0; if (neg) o[n++] = '-'; o[n++] = '0'; return true; }
  n = fmt_g_tail(o, neg, q, 3, p);
  return true;
}

// "%.2g" of strtod(text) given the scanned decimal.  Exact whenever the decimal is not within
// ~1e-15 (relative) of a two-digit rounding tie and stays far from double under/overflow:
// then the correctly rounded double and the decimal itself round to the same two digits.
JX_HD bool fmt_dec_2g(char* o, int& n, const Dec& d) {
  if (d.dropped) return false;
  if (d.m == 0) { n = 0; if (d.neg) o[n++] = '-'; o[n++] = '0'; return true; }
  uint64_t m = d.m;
  int nd = 1; { uint64_t t = m; while (t >= 10u) { t /= 10u; ++nd; } }
  int P = nd - 1 + d.e10;                          // decimal exponent of the leading digit
  if (P < -300 || P > 300) return false;          // stay inside the normal double range
  uint32_t two;
  if (nd <= 2) {
    two = (uint32_t)m; if (nd == 1) two *= 10u;
  } else {
    const int k = nd - 2;                          // digits to drop
    if (k > 14) return false;
    const uint64_t pw = pow10_u64(k);
    const uint64_t head = m / pw, tail = m % pw, half = pw / 2u;
    if (tail == half) return false;                // decimal tie: the binary value decides
    two = (uint32_t)head + (tail > half ? 1u : 0u);
    if (two == 100u) { two = 10u; ++P; }
  }
  n = fmt_g_tail(o, d.neg, two, 2, P);
  return true;
}

// One survivor line: sscanf-faithful field scan (same cursor rules as tokenize_line) that keeps
// every numeric field, then the reference's formatting with the substituted names.
struct LineFields {
  Dec pct, ev, sc;
  long long iv[7];
  int nconv;
  bool hard;
};

template <class R>
JX_HD void scan_fields(R& rd, int s, int e, LineFields& f) {
  f.nconv = 0; f.hard = false;
  for (int k = 0; k < 7; ++k) f.iv[k] = 0;
  f.pct.ok = f.ev.ok = f.sc.ok = false; f.pct.m = f.ev.m = f.sc.m = 0; f.pct.neg = f.ev.neg = f.sc.neg = false;
  f.pct.e10 = f.ev.e10 = f.sc.e10 = 0; f.pct.dropped = f.ev.dropped = f.sc.dropped = false;
  f.pct.special = f.ev.special = f.sc.special = false;
  int p = s;
  while (p < e && is_space(rd(p))) ++p;
  int a = p; while (p < e && !is_space(rd(p))) ++p;
  if (p == a) return;
  f.nconv = 1;
  while (p < e && is_space(rd(p))) ++p;
  a = p; while (p < e && !is_space(rd(p))) ++p;
  if (p == a) return;
  f.nconv = 2;
  while (p < e && is_space(rd(p))) ++p;
  int p2 = scan_decimal(rd, p, e, f.pct);
  if (f.pct.special) { f.hard = true; return; }
  if (!f.pct.ok) return;
  p = p2; f.nconv = 3;
  for (int k = 0; k < 7; ++k) {
    while (p < e && is_space(rd(p))) ++p;
    p2 = scan_int(rd, p, e);
    if (p2 == p) return;
    int q = p; bool neg = false;
    uint32_t c = rd(q);
    if (c == '-') { neg = true; ++q; } else if (c == '+') ++q;
    while (q < p2 && rd(q) == '0') ++q;
    if (p2 - q > 9) { f.hard = true; return; }     // beyond int32: strtol clamping, leave to glibc
    long long v = 0;
    for (; q < p2; ++q) v = v * 10 + (long long)(rd(q) - '0');
    f.iv[k] = neg ? -v : v;
    p = p2; ++f.nconv;
  }
  while (p < e && is_space(rd(p))) ++p;
  p2 = scan_decimal(rd, p, e, f.ev);
  if (f.ev.special) { f.hard = true; return; }
  if (!f.ev.ok) return;
  p = p2; f.nconv = 11;
  while (p < e && is_space(rd(p))) ++p;
  scan_decimal(rd, p, e, f.sc);
  if (f.sc.special) { f.hard = true; return; }
  if (!f.sc.ok) return;
  f.nconv = 12;
}

// numeric part of the output line: "\t%.2f\t%d...\t%.2g\t%.3g\n" into o; false = leave to glibc
JX_HD bool fmt_numeric_tail(char* o, int& n, const LineFields& f) {
  if (f.hard) return false;
  float pct = 0.f, sc = 0.f;
  if (f.pct.ok && !dec_to_f32(f.pct, pct)) return false;
  if (f.sc.ok && f.nconv == 12 && !dec_to_f32(f.sc, sc)) return false;
  long long v[7];
  for (int k = 0; k < 7; ++k) v[k] = f.iv[k];
  if (v[3] > v[4]) { long long t = v[3]; v[3] = v[4]; v[4] = t; }     // cblast.pyx:115-120
  if (v[5] > v[6]) { long long t = v[5]; v[5] = v[6]; v[6] = t; }
  n = 0;
  int k2;
  o[n++] = '\t';
  if (!fmt_f32_2f(o + n, k2, pct)) return false;
  n += k2;
  for (int k = 0; k < 7; ++k) { o[n++] = '\t'; n += fmt_int(o + n, v[k]); }
  o[n++] = '\t';
  Dec ev = f.ev;
  if (f.nconv < 11) { ev.m = 0; ev.neg = false; ev.dropped = false; }
  if (!fmt_dec_2g(o + n, k2, ev)) return false;
  n += k2;
  o[n++] = '\t';
  if (!fmt_f32_3g(o + n, k2, sc)) return false;
  n += k2;
  o[n++] = '\n';
  return true;
}

}  // namespace jx

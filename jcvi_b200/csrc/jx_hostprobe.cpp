// Test tooling: exposes the host build of jx_parse.cuh so that the CPU test-suite can check
// the tokenizer arithmetic against glibc (strtof / strtod / sscanf / printf) without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "jx_parse.cuh"
#include "jx_format.cuh"
#include "jx_fast5.cuh"
#include <vector>
#include <string>

namespace {
struct PtrReader { const unsigned char* p; inline uint32_t operator()(int i) const { return p[i]; } };
}

extern "C" {

// returns 1 ok / 0 hard / -1 no conversion; *consumed = chars consumed
int probe_f32(const char* s, float* out, int* consumed) {
  PtrReader rd{(const unsigned char*)s};
  jx::Dec d;
  int e = (int)strlen(s);
  int q = jx::scan_decimal(rd, 0, e, d);
  *consumed = q;
  if (d.special) return 0;
  if (!d.ok) return -1;
  return jx::dec_to_f32(d, *out) ? 1 : 0;
}

// returns 1/0 for (strtod(s) < 1e-10), -1 hard, -2 no conversion
int probe_lt1e10(const char* s) {
  PtrReader rd{(const unsigned char*)s};
  jx::Dec d;
  jx::scan_decimal(rd, 0, (int)strlen(s), d);
  if (d.special) return -1;
  if (!d.ok) return -2;
  return jx::dec_lt_1e10(d);
}

int probe_round3g(float x, float* out) { return jx::round3g_f32(x, *out) ? 1 : 0; }

// bulk check of round3g against glibc for every float32 with bit patterns in [lo, hi) step
// returns number of mismatches; first mismatching bit pattern in *bad
long probe_round3g_range(unsigned lo, unsigned hi, unsigned step, unsigned* bad, long* unsupported) {
  long mism = 0; *unsupported = 0;
  char buf[64];
  for (unsigned long u = lo; u < hi; u += step) {
    float x = jx::bits_to_float((uint32_t)u), got;
    if (!jx::round3g_f32(x, got)) { ++*unsupported; continue; }
    snprintf(buf, sizeof buf, "%.3g", (double)x);
    float want = strtof(buf, nullptr);
    if (jx::float_to_bits(want) != jx::float_to_bits(got)) { if (!mism) *bad = (unsigned)u; ++mism; }
  }
  return mism;
}

unsigned long long probe_hash(const char* s, int n) {
  PtrReader rd{(const unsigned char*)s};
  return jx::name_hash(rd, 0, n);
}

int probe_strip(const char* s, int n) {
  PtrReader rd{(const unsigned char*)s};
  int b = n;
  jx::strip_isoform(rd, 0, b);
  return b;
}

// tokenize one line (no terminator needed); outputs spans/score/nconv/eflag/flags
int probe_line(const char* s, int n, int* spans, float* score, int* nconv, int* eflag) {
  PtrReader rd{(const unsigned char*)s};
  jx::LineTok t;
  jx::tokenize_line(rd, 0, n, t);
  spans[0] = t.qa; spans[1] = t.qb; spans[2] = t.sa; spans[3] = t.sb;
  *score = t.score; *nconv = t.nconv; *eflag = t.eflag;
  return (t.hard ? 1 : 0) | (t.longname ? 2 : 0);
}

// the reference's own parse of the same line, via glibc sscanf with the cblast.pyx:15 format
int probe_line_glibc(const char* s, char* q, char* sb, float* score, double* evalue, int* ints) {
  float pct = 0; *score = 0; *evalue = 0; q[0] = sb[0] = 0;
  for (int i = 0; i < 7; ++i) ints[i] = 0;
  return sscanf(s, "%127s\t%127s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f", q, sb, &pct, ints, ints + 1, ints + 2,
                ints + 3, ints + 4, ints + 5, ints + 6, evalue, score);
}


// fast path vs generic scanner on one line; buffer must be padded (16 readable bytes past n)
namespace { struct WordRd { const unsigned char* p; inline uint32_t operator()(int i) const { uint32_t w; memcpy(&w, p + 4 * i, 4); return w; } }; }
int probe_fast_line(const char* buf, int s, int n, int* spans, float* score, int* eflag, int* strip_ends,
                    unsigned long long* hashes) {
  WordRd wr{(const unsigned char*)buf};
  jx::LineTok t;
  if (!jx::fast_line(wr, s, n, t)) return 0;
  spans[0] = t.qa; spans[1] = t.qb; spans[2] = t.sa; spans[3] = t.sb;
  *score = t.score; *eflag = t.eflag;
  jx::WordBytes<WordRd> B(wr);
  strip_ends[0] = jx::fast_strip(B, t.qa, t.qb, t.qbar);
  strip_ends[1] = jx::fast_strip(B, t.sa, t.sb, t.sbar);
  hashes[0] = jx::fast_hash(B, t.qa, strip_ends[0]);
  hashes[1] = jx::fast_hash(B, t.sa, strip_ends[1]);
  return 1;
}


int probe_fast_line3(const char* buf, int s, int limit, int* spans, float* score, int* eflag, int* strip_ends,
                     unsigned long long* hashes) {
  WordRd wr{(const unsigned char*)buf};
  jx::LineTok t;
  unsigned short sep[12];
  if (!jx::fast_line3(wr, s, limit, sep, 1, t)) return 0;
  spans[0] = t.qa; spans[1] = t.qb; spans[2] = t.sa; spans[3] = t.sb;
  *score = t.score; *eflag = t.eflag;
  jx::WordBytes<WordRd> B(wr);
  uint64_t h0, h1;
  strip_ends[0] = jx::fast_strip_hash(B, t.qa, t.qb, true, h0);
  strip_ends[1] = jx::fast_strip_hash(B, t.sa, t.sb, true, h1);
  hashes[0] = h0; hashes[1] = h1;
  return 1;
}


int probe_fmt_2f(float x, char* out) { int n; if (!jx::fmt_f32_2f(out, n, x)) return -1; out[n] = 0; return n; }
int probe_fmt_3g(float x, char* out) { int n; if (!jx::fmt_f32_3g(out, n, x)) return -1; out[n] = 0; return n; }
int probe_fmt_2g_text(const char* s, char* out) {
  PtrReader rd{(const unsigned char*)s};
  jx::Dec d;
  jx::scan_decimal(rd, 0, (int)strlen(s), d);
  if (d.special || !d.ok) return -2;
  int n; if (!jx::fmt_dec_2g(out, n, d)) return -1; out[n] = 0; return n;
}
// exhaustive-ish sweeps against glibc; return mismatches
long probe_fmt_3g_range(unsigned lo, unsigned hi, unsigned step, unsigned* bad, long* unsupported) {
  long mism = 0; *unsupported = 0; char a[64], b[64];
  for (unsigned long u = lo; u < hi; u += step) {
    float x = jx::bits_to_float((uint32_t)u); int n;
    if (!jx::fmt_f32_3g(a, n, x)) { ++*unsupported; continue; }
    a[n] = 0; snprintf(b, sizeof b, "%.3g", (double)x);
    if (strcmp(a, b)) { if (!mism) *bad = (unsigned)u; ++mism; }
  }
  return mism;
}
long probe_fmt_2f_range(unsigned lo, unsigned hi, unsigned step, unsigned* bad, long* unsupported) {
  long mism = 0; *unsupported = 0; char a[64], b[64];
  for (unsigned long u = lo; u < hi; u += step) {
    float x = jx::bits_to_float((uint32_t)u); int n;
    if (!jx::fmt_f32_2f(a, n, x)) { ++*unsupported; continue; }
    a[n] = 0; snprintf(b, sizeof b, "%.2f", (double)x);
    if (strcmp(a, b)) { if (!mism) *bad = (unsigned)u; ++mism; }
  }
  return mism;
}
// whole line: returns length of the numeric tail or -1 (leave to glibc)
int probe_fmt_line(const char* line, int n, char* out) {
  PtrReader rd{(const unsigned char*)line};
  jx::LineFields f;
  jx::scan_fields(rd, 0, n, f);
  int k;
  if (!jx::fmt_numeric_tail(out, k, f)) return -1;
  out[k] = 0;
  return k;
}


int probe_fast_line4(const char* buf, int s, int e, int* spans, float* score, int* eflag, int* strip_ends,
                     unsigned long long* hashes) {
  WordRd wr{(const unsigned char*)buf};
  jx::LineTok t;
  if (!jx::fast_line4(wr, s, e, t)) return 0;
  spans[0] = t.qa; spans[1] = t.qb; spans[2] = t.sa; spans[3] = t.sb;
  *score = t.score; *eflag = t.eflag;
  jx::WordBytes<WordRd> B(wr);
  for (int k = 0; k < 2; ++k) {
    const int a = k ? t.sa : t.qa, b = k ? t.sb : t.qb;
    jx::NameRegs nr; uint64_t h;
    if (jx::name_to_regs(B, a, b, true, nr, h)) { strip_ends[k] = a + nr.len; hashes[k] = h; }
    else { uint64_t h2; strip_ends[k] = jx::fast_strip_hash(B, a, b, true, h2); hashes[k] = h2; }
  }
  return 1;
}

// v5: masks built with the kernel's own classifier over a padded copy of the buffer
int probe_fast_line5(const char* buf, int buflen, int s, int e, int* spans, float* score, int* eflag, int* strip_ends,
                     unsigned long long* hashes) {
  const int padded = ((buflen + 31) / 32) * 32 + 256;
  std::vector<unsigned char> text((size_t)padded, (unsigned char)'\n');
  memcpy(text.data(), buf, (size_t)buflen);
  std::vector<uint32_t> mw((size_t)padded / 32), mn((size_t)padded / 32);
  uint32_t hi = 0;
  for (int seg = 0; seg < padded / 32; ++seg) {
    uint32_t w[8];
    memcpy(w, text.data() + 32 * seg, 32);
    jx::classify32(w, mw[(size_t)seg], mn[(size_t)seg], hi);
  }
  struct MaskRd { const uint32_t* p; inline uint32_t operator()(int i) const { return p[i]; } };
  WordRd wr{text.data()};
  MaskRd rw{mw.data()}, rn{mn.data()};
  jx::LineTok t;
  jx::WordBytes<WordRd> B(wr);
  if (!jx::fast_line5(B, rw, rn, s, e, t)) return 0;
  spans[0] = t.qa; spans[1] = t.qb; spans[2] = t.sa; spans[3] = t.sb;
  *score = t.score; *eflag = t.eflag;
  const bool maybe_bar = (hi & 0x80808080u) != 0u;
  for (int k = 0; k < 2; ++k) {
    const int a = k ? t.sa : t.qa, b = k ? t.sb : t.qb;
    const int bar = maybe_bar ? jx::find_bar(B, a, b) : -1;
    strip_ends[k] = jx::strip5(B, a, b, bar);
    hashes[k] = jx::fast_hash5(wr, a, strip_ends[k]);
  }
  return 1;
}

// formats.blast filter: the kernel's per-row decision (filter_parse + filter_remove, masks built with the kernel's
// classifier) for rows [rs[i], re[i]) of buf, on both routes, next to the reference's own arithmetic (glibc sscanf with
// the cblast.pyx:15 format, C float / int / double compares).  out[3 i + 0]: fast allowed, + 1: generic only,
// + 2: glibc;  1 kept, 0 removed, -1 refused (hard), -2 name of 128+ bytes.  route[i] = 1 when the fast route decided.
void probe_filter_rows(const char* buf, int buflen, const int* rs, const int* re, long n, double c_score, double c_pctid,
                       long long c_hitlen, double c_evalue, int noself, int inverse, signed char* out, signed char* route) {
  const int padded = ((buflen + 31) / 32) * 32 + 256;
  std::vector<unsigned char> text((size_t)padded, (unsigned char)'\n');
  memcpy(text.data(), buf, (size_t)buflen);
  std::vector<uint32_t> mw((size_t)padded / 32 + 2), mn((size_t)padded / 32 + 2);
  uint32_t hi = 0;
  for (int seg = 0; seg < padded / 32; ++seg) {
    uint32_t w[8];
    memcpy(w, text.data() + 32 * seg, 32);
    jx::classify32(w, mw[(size_t)seg], mn[(size_t)seg], hi);
  }
  struct MaskRd { const uint32_t* p; inline uint32_t operator()(int i) const { return p[i]; } };
  WordRd wr{text.data()};
  MaskRd rw{mw.data()}, rn{mn.data()};
  jx::WordBytes<WordRd> B(wr);
  PtrReader rd{text.data()};
  for (long i = 0; i < n; ++i) {
    for (int pass = 0; pass < 2; ++pass) {
      jx::FilterFields f;
      jx::filter_parse(B, rw, rn, rd, rs[i], re[i], pass == 0, c_evalue, f);
      if (pass == 0) route[i] = f.fast ? 1 : 0;
      signed char r;
      if (f.longname) r = -2;
      else if (f.hard) r = -1;
      else r = jx::filter_remove(rd, f, c_score, c_pctid, c_hitlen, false, inverse != 0, noself != 0) ? 0 : 1;
      out[3 * i + pass] = r;
    }
    std::string row((const char*)text.data() + rs[i], (size_t)(re[i] - rs[i]));
    char q[4096], sb[4096]; float pct = 0, score = 0; double ev = 0; int ints[7] = {0, 0, 0, 0, 0, 0, 0};
    q[0] = sb[0] = 0;
    sscanf(row.c_str(), "%4095s\t%4095s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f", q, sb, &pct, ints, ints + 1, ints + 2,
           ints + 3, ints + 4, ints + 5, ints + 6, &ev, &score);
    bool remove = (double)score < c_score || (double)pct < c_pctid || (long long)ints[0] < c_hitlen || ev > c_evalue;
    if (inverse) remove = !remove;
    remove = remove || (noself && strcmp(q, sb) == 0);
    out[3 * i + 2] = (strlen(q) >= 128 || strlen(sb) >= 128) ? -2 : (remove ? 0 : 1);
  }
}

// byte classes of one 32-byte block: the kernel's SWAR classifier (for comparison with the definition)
void probe_classify32(const unsigned char* bytes32, unsigned* wmask, unsigned* nmask, unsigned* hi) {
  uint32_t w[8];
  memcpy(w, bytes32, 32);
  uint32_t h = 0, a, b;
  jx::classify32(w, a, b, h);
  *wmask = a; *nmask = b; *hi = h & 0x80808080u;
}

// 32-bit slot hash (name table) of n names stored back to back with offsets off[0..n]
void probe_slot_hash_many(const char* blob, const unsigned* off, long n, unsigned* out) {
  PtrReader rd{(const unsigned char*)blob};
  for (long i = 0; i < n; ++i) out[i] = jx::slot_hash_bytes(rd, (int)off[i], (int)off[i + 1]);
}

// the kernel's route to the slot hash: six zero-padded words read from aligned text words (buffer must
// have 32 readable bytes after the name); must equal probe_slot_hash_many's byte-wise definition
unsigned probe_slot_hash6(const char* buf, int a, int len, unsigned* words6) {
  WordRd wr{(const unsigned char*)buf};
  uint32_t w[JX_SLOT_WORDS];
  jx::name_words6(wr, a, len, w);
  for (int k = 0; k < JX_SLOT_WORDS; ++k) words6[k] = w[k];
  return jx::slot_hash6(w, len);
}

}  // extern "C"

// strtod(s) > c decided by the kernel's exact routine: 1 / 0 / -1 refused / -2 no conversion
extern "C" int probe_gt_double(const char* s, double c) {
  PtrReader rd{(const unsigned char*)s};
  jx::Dec d;
  jx::scan_decimal(rd, 0, (int)strlen(s), d);
  if (d.special) return -1;
  if (!d.ok) return -2;
  return jx::dec_gt_double(d, c);
}

// v6: the kernel's first-level route (jx_fast6.cuh) on one line of a padded buffer, masks from the kernel's classifier.
// Returns 1 when accepted: spans (raw names), score, eflag, stripped ends (no '|' handling: the kernel uses this route
// only for tiles without one) and the version-2 slot hashes of the stripped names; 0 when refused.
#include "jx_fast6.cuh"
extern "C" int probe_fast_line6(const char* buf, int buflen, int s, int e, int* spans, float* score, int* eflag, int* strip_ends,
                                unsigned* hashes) {
  const int padded = ((buflen + 31) / 32) * 32 + 256;
  std::vector<unsigned char> text((size_t)padded, (unsigned char)'\n');
  memcpy(text.data(), buf, (size_t)buflen);
  std::vector<uint32_t> mw((size_t)padded / 32 + 2), mn((size_t)padded / 32 + 2);
  uint32_t hi = 0;
  for (int seg = 0; seg < padded / 32; ++seg) {
    uint32_t w[8];
    memcpy(w, text.data() + 32 * seg, 32);
    jx::classify32(w, mw[(size_t)seg], mn[(size_t)seg], hi);
  }
  struct MaskRd { const uint32_t* p; inline uint32_t operator()(int i) const { return p[i]; } };
  WordRd wr{text.data()};
  MaskRd rw{mw.data()}, rn{mn.data()};
  jx::WordBytes<WordRd> B(wr);
  jx::Line6 o;
  if (!jx::fast_line6(B, wr, rw, rn, s, e - s, o)) return 0;
  spans[0] = s; spans[1] = s + o.t0; spans[2] = s + o.t0 + 1; spans[3] = s + o.t1;
  *score = o.score; *eflag = o.eflag;
  PtrReader rd{text.data()};
  for (int k = 0; k < 2; ++k) {
    const int a = spans[2 * k], len = spans[2 * k + 1] - a;
    const int sl = jx::strip6(B, a, len, B.load32(a));
    strip_ends[k] = a + sl;
    hashes[k] = jx::slot_hash_bytes(rd, a, a + sl);
  }
  return 1;
}

// the name table jx_bed_load builds for n names (blob + offsets): out = {slot bits, bucket bits, overflow names,
// table words}; every name must then be found by the lookup the kernels do.  Returns the number of names NOT found.
#include "jx_nametable.h"
extern "C" long probe_build_table(const char* blob, const unsigned* off, long n, long* out) {
  std::vector<uint32_t> nid((size_t)n);
  for (long i = 0; i < n; ++i) nid[(size_t)i] = (uint32_t)i;
  jx_bed bed; memset(&bed, 0, sizeof bed);
  bed.n_genes = n; bed.names = blob; bed.name_off = off; bed.name_id = nid.data();
  jxnt::BuiltSide B;
  const std::string e = jxnt::build_side(&bed, B);
  if (!e.empty()) return -1;
  out[0] = 32 - (long)B.cshift; out[1] = 32 - (long)B.bshift; out[2] = (long)B.ovf_hash.size(); out[3] = (long)B.tab.size();
  PtrReader rd{(const unsigned char*)blob};
  long missing = 0;
  for (long i = 0; i < n; ++i) {
    const int a = (int)off[i], b = (int)off[i + 1];
    const uint32_t h = jx::slot_hash_bytes(rd, a, b);
    const uint32_t idx = jx::slot_index(h, B.disp[h >> B.bshift], B.cshift);
    const uint32_t* slot = &B.tab[(size_t)idx * 8];
    bool ok = slot[0] != 0;
    if (ok) {                                   // compare the first 24 bytes like the kernel
      for (int k = 0; k < 24 && ok; ++k) {
        const unsigned char want = a + k < b ? (unsigned char)blob[a + k] : 0;
        ok = (unsigned char)((slot[2 + (k >> 2)] >> (8 * (k & 3))) & 0xFFu) == want;
      }
    }
    if (!ok) {                                  // parked in the overflow list?
      bool found = false;
      for (size_t k = 0; k < B.ovf_hash.size(); ++k) if (B.ovf_hash[k] == h) found = true;
      if (!found) ++missing;
    }
  }
  return missing;
}

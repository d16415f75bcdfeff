// Tokenizer fast path, version 6: the per-line parse of a REGULAR line in ~300 thread instructions.
//
// Same contract as jx_fast5.cuh (which stays as the second-level route): the byte-class masks W (byte <= 0x20)
// and N (neither digit nor tab) of the staged window exist; this function accepts only lines of the shape LAST /
// BLAST+ / DIAMOND / MMseqs2 write -- 12 non-empty fields, 11 single tabs, both names within the first 32 bytes,
// evalue <= 14 bytes, score <= 8 bytes -- and returns exactly what sscanf with the format of
// src/jcvi/formats/cblast.pyx:15 yields for them; everything else is REFUSED (false) and goes to fast_line5 /
// tokenize_line.  What is new against version 5:
//   * 32-bit mask windows wherever the field layout allows it (names, tail), one 64-bit window for the middle;
//   * evalue is DECIDED (value < 1e-10), not converted: for the normalised form d[.ddd]e-XX only the exponent
//     is read; plain decimals of <= 11 bytes only need "all digits zero";
//   * score = I * 10^f + F from two right-justified 4-digit groups picked with byte permutes, then one correctly
//     rounded float division (both operands exact in float32);
//   * names: three zero-padded words for names of <= 12 bytes (masks from a constant table); the name hash
//     (jx_fast5.cuh slot_hash_*) is a multiply-add chain that runs on the FMA pipe.
// Reference semantics: src/jcvi/formats/cblast.pyx:15,106-112 (sscanf), src/jcvi/utils/cbook.py:308-329 (gene_name).
#pragma once
#include "jx_fast5.cuh"

namespace jx {

JX_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {      // bytes of (b:a) picked by the four selector nibbles
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  const uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) {
    const uint32_t n = (sel >> (4 * k)) & 0xFu;
    uint32_t byte = (uint32_t)(v >> (8 * (n & 7u))) & 0xFFu;
    if (n & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;          // sign replication
    r |= byte << (8 * k);
  }
  return r;
#endif
}

// bits < 8 r, 0 <= r <= 4 (r = 4 gives all ones)
#define JX_BYTEMASKS {0x00000000u, 0x000000FFu, 0x0000FFFFu, 0x00FFFFFFu, 0xFFFFFFFFu}
static const uint32_t kByteMaskHost[5] = JX_BYTEMASKS;
// masks of the three zero-padded words of a name of 0..12 bytes
#define JX_NAMEMASKS { {0u,0u,0u,0u}, {0xFFu,0u,0u,0u}, {0xFFFFu,0u,0u,0u}, {0xFFFFFFu,0u,0u,0u}, {0xFFFFFFFFu,0u,0u,0u}, \
  {0xFFFFFFFFu,0xFFu,0u,0u}, {0xFFFFFFFFu,0xFFFFu,0u,0u}, {0xFFFFFFFFu,0xFFFFFFu,0u,0u}, {0xFFFFFFFFu,0xFFFFFFFFu,0u,0u}, \
  {0xFFFFFFFFu,0xFFFFFFFFu,0xFFu,0u}, {0xFFFFFFFFu,0xFFFFFFFFu,0xFFFFu,0u}, {0xFFFFFFFFu,0xFFFFFFFFu,0xFFFFFFu,0u}, \
  {0xFFFFFFFFu,0xFFFFFFFFu,0xFFFFFFFFu,0u} }
static const uint32_t kNameMaskHost[13][4] = JX_NAMEMASKS;
#if defined(__CUDACC__)
static __constant__ uint32_t kByteMaskDev[5] = JX_BYTEMASKS;
static __constant__ uint32_t kNameMaskDev[13][4] = JX_NAMEMASKS;
#endif
JX_HD uint32_t byte_mask(int r) {
#if defined(__CUDA_ARCH__)
  return kByteMaskDev[r];
#else
  return kByteMaskHost[r];
#endif
}
JX_HD uint32_t name_mask(int len, int k) {
#if defined(__CUDA_ARCH__)
  return kNameMaskDev[len][k];
#else
  return kNameMaskHost[len][k];
#endif
}

// the two aligned-word reads of an unaligned 8-byte load share their middle word
template <class TW>
JX_HD void load64(const TW& wr, int p, uint32_t& lo, uint32_t& hi) {   // bytes [p, p + 8)
  const int i = p >> 2, sh = (p & 3) * 8;
  const uint32_t x0 = wr(i), x1 = wr(i + 1), x2 = wr(i + 2);
  lo = funnel_r(x0, x1, sh);
  hi = funnel_r(x1, x2, sh);
}

struct Line6 {
  int t0, t1;        // tabs 1 and 2, relative to the line start: query = [0, t0), subject = [t0 + 1, t1)
  float score;
  int eflag;         // evalue < 1e-10
};

// evalue field = bytes [p, p + L) (1 <= L <= 14), nb = its N bits.  Decides value < 1e-10 for
//   d[.d*]e-X{1,3}  (one leading digit)      and      plain decimals d*[.d*] of <= 11 bytes;
// false = some other shape (generic route).
template <class BR, class TW>
JX_HD bool evalue6(const BR& B, const TW& wr, int p, int L, uint32_t nb, int& flag) {
  if (nb & 1u) return false;                             // must start with a digit
  const bool dot1 = (nb & 2u) != 0u && B(p + 1) == '.';  // a '.' right after the first digit
  const uint32_t nb2 = dot1 ? (nb & ~2u) : nb;
  if (nb2 == 0u) {
    // no exponent: digits with at most one '.' after the first digit.  With <= 11 bytes a non-zero value is
    // >= 1e-9 ("0." + 9 digits), so value < 1e-10 iff every digit is '0'
    if (L > 11) return false;
    uint32_t lo, hi, x2;
    load64(wr, p, lo, hi);
    x2 = L > 8 ? B.load32(p + 8) : 0u;
    // bytes that are not '0' (the '.' is neutralised): xor with "0000", mask to the field
    const uint32_t dotm = dot1 ? 0x0000FF00u : 0u;
    const int r0 = L < 4 ? L : 4, r1 = L < 4 ? 0 : (L < 8 ? L - 4 : 4), r2 = L < 8 ? 0 : L - 8;
    const uint32_t d = ((lo ^ 0x30303030u) & byte_mask(r0) & ~dotm) | ((hi ^ 0x30303030u) & byte_mask(r1)) |
                       ((x2 ^ 0x30303030u) & byte_mask(r2 > 4 ? 4 : r2));
    flag = d == 0u ? 1 : 0;
    return true;
  }
  // exponent form: the remaining N bits must be exactly 'e' and '-' next to each other, then 1..3 digits
  const int pe = ctz32(nb2);
  if ((nb2 >> pe) != 3u) return false;
  const int nx = L - pe - 2;
  if (nx < 1 || nx > 3) return false;                    // (pe >= 1: bit 0 is clear)
  const uint32_t tail = B.load32(p + L - 4);             // ... the last four bytes of the field: [e][-]XX / [-]XXX / ...
  // 'e' and '-' sit right before the nx exponent digits
  const uint32_t em = prmt(tail, 0u, nx == 1 ? 0x4421u : (nx == 2 ? 0x4410u : 0x4403u));   // byte0 = 'e' (or junk for nx == 3), byte1 = '-'
  if (((em >> 8) & 0xFFu) != (uint32_t)'-') return false;
  if (nx < 3) { if (((em & 0xFFu) | 0x20u) != (uint32_t)'e') return false; }
  else if ((B(p + pe) | 0x20u) != (uint32_t)'e') return false;
  // exponent value from the top nx bytes of `tail`
  const int xs = 8 * (4 - nx);
  const uint32_t dg = (tail >> xs) - (0x30303030u >> xs);              // the nx digits, first digit in the lowest byte (no borrows: 'e' / '-' are shifted out first)
  const uint32_t X = nx == 1 ? (dg & 0xFFu)
                   : nx == 2 ? (dg & 0xFFu) * 10u + ((dg >> 8) & 0xFFu)
                             : (dg & 0xFFu) * 100u + ((dg >> 8) & 0xFFu) * 10u + ((dg >> 16) & 0xFFu);
  const uint32_t d0 = B(p) - (uint32_t)'0';              // leading digit
  if (X >= 11u) { flag = 1; return true; }               // d.ddd * 10^-X < 10^(1-X) <= 1e-10
  if (d0 == 0u) return false;                            // 0.ddde-X: not normalised, generic route
  flag = 0;                                              // d >= 1, X <= 10: value >= 1e-10
  return true;
}

// score field = bytes [p, p + L) (1 <= L <= 8), nb = its N bits: digits with at most one '.' that is not the first
// byte; integer part <= 8 digits, fraction <= 4 digits, all digits < 2^24 -> float32 exactly like strtof.
template <class BR, class TW>
JX_HD bool score6(const BR& B, const TW& wr, int p, int L, uint32_t nb, float& out) {
  int nI = L, nF = 0;
  if (nb) {
    if (nb & (nb - 1u)) return false;
    const int dot = ctz32(nb);
    if (dot == 0 || B(p + dot) != '.') return false;
    nI = dot; nF = L - 1 - dot;
    if (nF > 4) return false;
  }
  uint32_t lo, hi;
  load64(wr, p, lo, hi);
  // integer part: last four digits (right-justified group) and the digits before them
  uint32_t ilo, ihi = 0u;
  if (nI <= 4) {
    ilo = (lo - 0x30303030u) << (8 * (4 - nI));
  } else {
    ilo = prmt(lo, hi, 0x3210u + 0x1111u * (uint32_t)(nI - 4)) - 0x30303030u;
    ihi = (lo - 0x30303030u) << (8 * (8 - nI));
  }
  uint32_t m = digits4(ilo);
  if (nI > 4) m += digits4(ihi) * 10000u;
  if (nF) {
    const uint32_t fw = prmt(lo, hi, 0x3210u + 0x1111u * (uint32_t)(nI + 1));     // the bytes after the '.'
    const uint32_t f = digits4((fw - 0x30303030u) << (8 * (4 - nF)));
    m = m * pow10_u32(nF) + f;
  }
  if (m >= (1u << 24)) return false;
  const float fm = (float)m;
  if (nF == 0) { out = fm; return true; }
#if defined(__CUDA_ARCH__)
  out = __fdiv_rn(fm, pow10_f32(nF));
#else
  out = fm / pow10_f32(nF);
#endif
  return true;
}

// [s, s + len) = the line without its terminator; the window must hold s + len + 8 bytes.
// First half: tabs 1 and 2 (the two names).  The kernel issues the name-table loads between the two halves, so that
// their latency is covered by the second half's arithmetic.
template <class MW, class MN>
JX_HD bool fast_line6_names(const MW& mw, const MN& mn, const int s, const int len, Line6& o) {
  if (len < 23) return false;                            // 12 non-empty fields and 11 tabs
  // tabs 1 and 2 are the first two bytes <= 0x20, inside the first 32 bytes
  const uint32_t WA = bits32(mw, s), NA = bits32(mn, s);
  const uint32_t WA1 = WA & (WA - 1u);
  if (WA1 == 0u) return false;
  const int t0 = ctz32(WA), t1 = ctz32(WA1);
  if ((((NA >> t0) | (NA >> t1)) & 1u) != 0u || t0 < 1 || t1 - t0 < 2) return false;      // both tabs; names not empty
  o.t0 = t0; o.t1 = t1;
  return true;
}
// Second half: structure of the rest of the line, evalue decision, score.
template <class BR, class TW, class MW, class MN>
JX_HD bool fast_line6_rest(const BR& B, const TW& wr, const MW& mw, const MN& mn, const int s, const int len, Line6& o) {
  const int t1 = o.t1;
  // ---- tail: tabs 10 and 11 are the last two bytes <= 0x20, inside the last 32 bytes ------------------------------
  const int az = len > 32 ? len - 32 : 0;               // window = [az, az + 32) relative to s
  const int zl = len - az;                               // its bytes that belong to the line (32, or len when shorter)
  uint32_t WZ = bits32(mw, s + az);
  const uint32_t NZ = bits32(mn, s + az);
  if (zl < 32) WZ &= below32(zl);
  if (WZ == 0u) return false;
  const int h10 = hibit32(WZ);
  const uint32_t WZ1 = WZ ^ (1u << h10);
  if (WZ1 == 0u) return false;
  const int h9 = hibit32(WZ1);
  const int ls = zl - 1 - h10, le = h10 - h9 - 1;
  if ((((NZ >> h10) | (NZ >> h9)) & 1u) != 0u || ls < 1 || ls > 8 || le < 1 || le > 14) return false;
  const uint32_t sbits = (NZ >> (h10 + 1)) & below32(ls);          // h10 <= 30 because ls >= 1
  const uint32_t ebits = (NZ >> (h9 + 1)) & below32(le);
  // ---- middle: from tab 2 to tab 10 -- pctid and seven integers, eight single tabs in between ------------------
  const int t9 = az + h9;
  const int dm = t9 - t1;                                // tab 2 = bit 0, tab 10 = bit dm of the 64-bit window
  if (dm < 16 || dm > 63) return false;
  const uint64_t mid = below64(dm + 1);
  const uint64_t Wm = bits64(mw, s + t1) & mid;
  const uint64_t Pm = bits64(mn, s + t1) & mid;
  if (popc64(Wm) != 9 || (Wm & (Wm >> 1)) != 0ull) return false;  // 9 separators, no empty field
  const uint32_t wl = (uint32_t)Wm & ~1u;
  if (wl == 0u) return false;
  const int t2r = ctz32(wl);                             // pctid = bytes [1, t2r) of the window
  if (Pm) {                                              // the only non-digit allowed: one '.' inside pctid
    const uint32_t pl = (uint32_t)Pm;
    if ((Pm >> 32) != 0ull || (pl & (pl - 1u)) != 0u) return false;
    const int dp = ctz32(pl);
    if (dp >= t2r || t2r < 3 || B(s + t1 + dp) != '.') return false;
  }
  // ---- evalue and score ---------------------------------------------------------------------------------------------
  if (!evalue6(B, wr, s + t9 + 1, le, ebits, o.eflag)) return false;
  return score6(B, wr, s + az + h10 + 1, ls, sbits, o.score);
}
template <class BR, class TW, class MW, class MN>
JX_HD bool fast_line6(const BR& B, const TW& wr, const MW& mw, const MN& mn, const int s, const int len, Line6& o) {
  return fast_line6_names(mw, mn, s, len, o) && fast_line6_rest(B, wr, mw, mn, s, len, o);
}

// gene_name (cbook.py:308-329) on a name WITHOUT '|' (the caller knows that the tile holds none): new length
template <class BR>
JX_HD int strip6(const BR& B, int a, int len, uint32_t first_word) {
  if (len < 2) return len;
  const bool ev = (first_word & 0xFFFFu) == 0x7665u;     // "ev"
  return (!ev && B(a + len - 2) == '.' && B(a + len - 1) != '.') ? len - 2 : len;
}

}  // namespace jx

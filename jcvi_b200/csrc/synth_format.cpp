// Bench/test tooling (not on the product path): formats synthetic BLAST tab12 lines from
// per-line integer arrays produced by jcvi_b200/synth.py.  Multi-threaded so that the
// 20 M-line BASELINE workload is generated in seconds on the GPU box's host cores.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {
inline char* put_uint(char* p, uint32_t v) {
  char tmp[12]; int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char* put_name(char* p, const char* prefix, int plen, int32_t gene, int iso, int digits) {
  memcpy(p, prefix, (size_t)plen); p += plen;
  char tmp[16];
  int v = gene;
  for (int i = digits - 1; i >= 0; --i) { tmp[i] = (char)('0' + v % 10); v /= 10; }
  memcpy(p, tmp, (size_t)digits); p += digits;
  if (iso > 0) { *p++ = '.'; p = put_uint(p, (uint32_t)iso); }
  return p;
}
}  // namespace

extern "C" {

// Each line: <qprefix><qgene %0*d>.<qiso>\t<sprefix><sgene>.<siso>\t%.2f pctid\t7 ints\t%.1e evalue\t%.1f score
// score10 = score*10 (int); evalue = 10^(-score/ediv).  iso == 0 -> no suffix.
// Returns bytes written (<= cap) or -1 when cap is too small.
int64_t synth_format_lines(int64_t n, const int32_t* qgene, const int32_t* sgene, const uint8_t* qiso,
                           const uint8_t* siso, const int32_t* score10, const int32_t* pct100,
                           const int32_t* ints7, const char* qprefix, const char* sprefix, int digits,
                           double ediv, int score_style, char* out, int64_t cap, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  const int qpl = (int)strlen(qprefix), spl = (int)strlen(sprefix);
  const int64_t maxline = qpl + spl + 2 * (digits + 4) + 10 + 7 * 12 + 16 + 16 + 16;
  std::vector<std::string> parts((size_t)nthreads);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    th.emplace_back([&, t]() {
      int64_t lo = n * t / nthreads, hi = n * (t + 1) / nthreads;
      std::string& s = parts[(size_t)t];
      s.resize((size_t)((hi - lo) * maxline));
      char* base = &s[0];
      char* p = base;
      for (int64_t i = lo; i < hi; ++i) {
        p = put_name(p, qprefix, qpl, qgene[i], qiso[i], digits); *p++ = '\t';
        p = put_name(p, sprefix, spl, sgene[i], siso[i], digits); *p++ = '\t';
        p += snprintf(p, 16, "%.2f", pct100[i] / 100.0); *p++ = '\t';
        for (int k = 0; k < 7; ++k) { p = put_uint(p, (uint32_t)ints7[i * 7 + k]); *p++ = '\t'; }
        double score = score10[i] / 10.0;
        double ev = std::pow(10.0, -score / ediv);
        if (ev < 1e-300) { *p++ = '0'; *p++ = '.'; *p++ = '0'; }
        else p += snprintf(p, 16, "%.1e", ev);
        *p++ = '\t';
        if (score_style == 1 && score10[i] % 10 == 0) p = put_uint(p, (uint32_t)(score10[i] / 10));  // LAST-like ints
        else p += snprintf(p, 16, "%.1f", score);
        *p++ = '\n';
      }
      s.resize((size_t)(p - base));
    });
  }
  for (auto& x : th) x.join();
  int64_t total = 0;
  for (auto& s : parts) total += (int64_t)s.size();
  if (total > cap) return -1;
  char* o = out;
  for (auto& s : parts) { memcpy(o, s.data(), s.size()); o += s.size(); }
  return total;
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
//
// Sequential CPU restatement of the MCscan anchor-detection hot path of
// tanghaibao/jcvi v1.6.6 (blastfilter -> synteny scan -> synteny liftover), used as
// the parity oracle for the CUDA implementation in jcvi_b200/csrc and as the
// `cpu_baseline` leg of bench.py.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
//
// It deliberately follows the reference's *control flow* (per-line loop, dicts keyed
// by name strings, stable score sort, first-seen sets, Grouper list merging, KD-tree
// traversal) rather than the data-parallel formulation of the product, so that the
// two implementations are independent.  Every function cites the reference
// file:line (relative to /root/reference/) it restates.
//
// Parity pinning: tests/test_oracle_golden.py checks this oracle byte-for-byte
// against (a) the reference's own golden files
// (tests/compara/synteny.py/references/*.anchors) and (b) fixtures produced by
// running the unmodified reference (Cython cblast build) in this container
// (tests/golden/make_golden.py).
//
// Also restated, each with its own golden vectors from the unmodified reference (tests/golden/make_golden_*.py) and
// cross-checked against the live reference on randomized inputs (tests/live_reference_check.py): formats.blast cscore
// and filter, synteny mcscan and simple, compara.synfind.
//
// Third-party arithmetic restated here because it is not under /root/reference:
//   * glibc sscanf/sprintf  -- called directly, exactly like cblast.pyx does.
//   * natsort (unpinned dep, pyproject.toml:49) -- default natural-sort key.
//   * scipy.spatial.cKDTree (scipy 1.18.1 installed here; pyproject.toml:61 unpinned)
//     -- build (balanced, compact, leafsize=16) and k=1, p=1 query; validated
//     black-box against the installed scipy in tests/test_oracle_kdtree.py.
//   * numpy.polyfit (synteny simple's orientation) -- the sign of the exact integer covariance; a covariance of
//     exactly 0 is rounding noise in the reference and "+" here (parity unpinned for such flat blocks).
#include <algorithm>
#include <set>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {

thread_local std::string g_err;

// ---------------------------------------------------------------------------------
// file helpers
// ---------------------------------------------------------------------------------
bool read_file(const char* path, std::string& out) {
  FILE* f = fopen(path, "rb");
  if (!f) { g_err = std::string("cannot open ") + path; return false; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize((size_t)n);
  if (n > 0 && fread(&out[0], 1, (size_t)n, f) != (size_t)n) { fclose(f); g_err = "short read"; return false; }
  fclose(f);
  return true;
}

// Python text-mode iteration (universal newlines): "\n", "\r\n" and a lone "\r" all end a
// line.  Calls fn(begin, end) for every line, [begin,end) excluding the terminator.
template <class F>
void for_each_line(const char* p, size_t n, F fn) {
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    while (j < n && p[j] != '\n' && p[j] != '\r') ++j;
    fn(p + i, p + j);
    if (j < n && p[j] == '\r' && j + 1 < n && p[j + 1] == '\n') ++j;
    i = j + 1;
  }
}

std::string strip(const std::string& s) {  // str.strip(): ASCII whitespace
  size_t a = 0, b = s.size();
  auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); };
  while (a < b && ws(s[a])) ++a;
  while (b > a && ws(s[b - 1])) --b;
  return s.substr(a, b - a);
}

std::vector<std::string> split_ws(const std::string& s) {  // str.split()
  std::vector<std::string> out;
  size_t i = 0, n = s.size();
  auto ws = [](char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); };
  while (i < n) {
    while (i < n && ws(s[i])) ++i;
    size_t j = i;
    while (j < n && !ws(s[j])) ++j;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}

// ---------------------------------------------------------------------------------
// natsort default key (third-party, restated): chunks alternate str / int.
// ---------------------------------------------------------------------------------
int natcmp(const std::string& a, const std::string& b) {
  // tuple compare of: split on ([0-9]+), empty strings dropped, "" prepended when the first
  // chunk is numeric  ==> even positions are strings, odd positions are ints.
  size_t i = 0, j = 0;
  bool want_str = true;
  for (;;) {
    bool enda = i >= a.size(), endb = j >= b.size();
    if (want_str) {
      size_t i2 = i, j2 = j;
      while (i2 < a.size() && !(a[i2] >= '0' && a[i2] <= '9')) ++i2;
      while (j2 < b.size() && !(b[j2] >= '0' && b[j2] <= '9')) ++j2;
      // a string chunk exists at this position iff (not at end); an empty chunk only
      // happens at position 0 (the prepended "").
      bool hasa = !enda, hasb = !endb;
      if (!hasa || !hasb) return hasa ? 1 : (hasb ? -1 : 0);
      int c = a.compare(i, i2 - i, b, j, j2 - j);
      if (c) return c < 0 ? -1 : 1;
      i = i2; j = j2;
    } else {
      if (enda || endb) return enda ? (endb ? 0 : -1) : 1;
      size_t i2 = i, j2 = j;
      while (i2 < a.size() && a[i2] >= '0' && a[i2] <= '9') ++i2;
      while (j2 < b.size() && b[j2] >= '0' && b[j2] <= '9') ++j2;
      size_t ia = i, jb = j;
      while (ia + 1 < i2 && a[ia] == '0') ++ia;
      while (jb + 1 < j2 && b[jb] == '0') ++jb;
      size_t la = i2 - ia, lb = j2 - jb;
      if (la != lb) return la < lb ? -1 : 1;
      int c = a.compare(ia, la, b, jb, lb);
      if (c) return c < 0 ? -1 : 1;
      i = i2; j = j2;
    }
    want_str = !want_str;
  }
}

// ---------------------------------------------------------------------------------
// BED  (src/jcvi/formats/bed.py:55-73 BedLine, 142-167 Bed.__init__, 196-198 order)
// ---------------------------------------------------------------------------------
struct BedLine {
  std::string seqid, accn;
  long start = 0, end = 0;  // start = col2 + 1 (bed.py:59)
  std::vector<std::string> args;  // tab-split columns, for __str__ (bed.py:75-89)
};

struct Bed {
  std::string filename;
  std::vector<BedLine> rows;                      // sorted
  std::unordered_map<std::string, int> order;     // accn -> rank (later duplicate wins)
};

bool load_bed(const char* path, Bed& bed) {
  std::string txt;
  if (!read_file(path, txt)) return false;
  bed.filename = path;
  bool ok = true;
  for_each_line(txt.data(), txt.size(), [&](const char* b, const char* e) {
    std::string line(b, e);
    // bed.py:154-160  skip blank, '#', "browser ", "track name"
    if (strip(line).empty() || line[0] == '#' || line.rfind("browser ", 0) == 0 ||
        line.rfind("track name", 0) == 0)
      return;
    std::string s = strip(line);
    BedLine r;
    size_t i = 0;
    while (true) {
      size_t j = s.find('\t', i);
      r.args.push_back(s.substr(i, j == std::string::npos ? std::string::npos : j - i));
      if (j == std::string::npos) break;
      i = j + 1;
    }
    if (r.args.size() < 4) { ok = false; g_err = "BED line with <4 columns"; return; }
    r.seqid = r.args[0];
    r.start = atol(r.args[1].c_str()) + 1;
    r.end = atol(r.args[2].c_str());
    if (r.start > r.end) { ok = false; g_err = "BED start > end"; return; }  // bed.py:61 assert
    r.accn = r.args[3];
    bed.rows.push_back(std::move(r));
  });
  if (!ok) return false;
  // bed.py:147,166-167: sort by (natsort_key(seqid), start, accn); list.sort is stable
  std::stable_sort(bed.rows.begin(), bed.rows.end(), [](const BedLine& a, const BedLine& b) {
    int c = natcmp(a.seqid, b.seqid);
    if (c) return c < 0;
    if (a.start != b.start) return a.start < b.start;
    return a.accn < b.accn;
  });
  for (size_t i = 0; i < bed.rows.size(); ++i) bed.order[bed.rows[i].accn] = (int)i;
  return true;
}

// ---------------------------------------------------------------------------------
// gene_name  (src/jcvi/utils/cbook.py:308-329)
// ---------------------------------------------------------------------------------
std::string gene_name(const std::string& st0) {
  bool use_sep = !(st0.size() >= 2 && st0[0] == 'e' && st0[1] == 'v');
  std::string st = st0.substr(0, st0.find('|'));
  std::string name = st, suffix;
  if (use_sep) {
    size_t k = st.rfind('.');
    if (k != std::string::npos) { name = st.substr(0, k); suffix = st.substr(k + 1); }
  }
  if (suffix.size() != 1) name = st;
  return name;
}

// ---------------------------------------------------------------------------------
// BlastLine  (src/jcvi/formats/cblast.pyx:15,84-120 parse; 17,144-156 __str__)
// ---------------------------------------------------------------------------------
struct BlastLine {
  char query[136];   // cblast.pyx:85-86 uses char[128]; longer names are rejected below
  char subject[136];
  int hitlen = 0, nmismatch = 0, ngaps = 0, qstart = 0, qstop = 0, sstart = 0, sstop = 0;
  float pctid = 0.f, score = 0.f;
  double evalue = 0.0;
  int qi = 0, si = 0;
  long line = 0;  // ordinal among parsed lines (for stable sorting)
};

const char* kBlastFormat = "%135s\t%135s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f";
const char* kBlastOutput = "%s\t%s\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.3g";

bool parse_blast_line(const char* b, const char* e, BlastLine& bl, std::string& scratch) {
  scratch.assign(b, e);
  scratch.push_back('\n');
  bl = BlastLine();
  bl.query[0] = bl.subject[0] = 0;
  sscanf(scratch.c_str(), kBlastFormat, bl.query, bl.subject, &bl.pctid, &bl.hitlen, &bl.nmismatch,
         &bl.ngaps, &bl.qstart, &bl.qstop, &bl.sstart, &bl.sstop, &bl.evalue, &bl.score);
  if (strlen(bl.query) >= 128 || strlen(bl.subject) >= 128) {
    g_err = "BLAST name of 128+ characters overflows cblast.pyx's char[128] (undefined behaviour in the reference)";
    return false;
  }
  if (bl.qstart > bl.qstop) std::swap(bl.qstart, bl.qstop);  // cblast.pyx:115-120
  if (bl.sstart > bl.sstop) std::swap(bl.sstart, bl.sstop);
  return true;
}

// formats/blast.py:90-95: every line whose first character is not '#'
bool load_blast(const char* path, std::vector<BlastLine>& out) {
  std::string txt;
  if (!read_file(path, txt)) return false;
  std::string scratch;
  bool ok = true;
  long n = 0;
  for_each_line(txt.data(), txt.size(), [&](const char* b, const char* e) {
    if (!ok) return;
    if (b < e && *b == '#') return;
    BlastLine bl;
    if (!parse_blast_line(b, e, bl, scratch)) { ok = false; return; }
    bl.line = n++;
    out.push_back(bl);
  });
  return ok;
}

// ---------------------------------------------------------------------------------
// Grouper  (src/jcvi/utils/grouper.py:40-81) -- list-merging disjoint sets whose
// iteration order is dict insertion order.
// ---------------------------------------------------------------------------------
struct Grouper {
  std::vector<long> keys;                       // insertion order of _mapping
  std::unordered_map<long, int> group_of;       // key -> group slot
  std::vector<std::vector<long>> groups;        // slot -> member list (may be emptied)

  int setdefault(long a) {
    auto it = group_of.find(a);
    if (it != group_of.end()) return it->second;
    groups.push_back({a});
    keys.push_back(a);
    group_of[a] = (int)groups.size() - 1;
    return (int)groups.size() - 1;
  }
  void join(long a, const std::vector<long>& args) {   // grouper.py:44-61
    int set_a = setdefault(a);
    for (long arg : args) {
      auto it = group_of.find(arg);
      if (it == group_of.end()) {
        groups[set_a].push_back(arg);
        keys.push_back(arg);
        group_of[arg] = set_a;
      } else if (it->second != set_a) {
        int set_b = it->second;
        if (groups[set_b].size() > groups[set_a].size()) std::swap(set_a, set_b);
        for (long el : groups[set_b]) { groups[set_a].push_back(el); group_of[el] = set_a; }
        groups[set_b].clear();
      }
    }
  }
  void join(long a, long b) { join(a, std::vector<long>{b}); }
  std::vector<std::vector<long>> iter() const {       // grouper.py:73-81
    std::vector<std::vector<long>> out;
    std::unordered_set<int> seen;
    for (long k : keys) {
      int g = group_of.at(k);
      if (seen.insert(g).second) out.push_back(groups[g]);
    }
    return out;
  }
};

// ---------------------------------------------------------------------------------
// AnchorFile  (src/jcvi/compara/base.py:9-27; read_block formats/base.py:610-639)
// ---------------------------------------------------------------------------------
struct AnchorBlocks {
  std::vector<std::vector<std::vector<std::string>>> blocks;  // block -> line -> tokens
};

bool load_anchors(const char* path, AnchorBlocks& ac) {
  std::string txt;
  if (!read_file(path, txt)) return false;
  std::vector<std::string> lines;
  for_each_line(txt.data(), txt.size(), [&](const char* b, const char* e) { lines.emplace_back(b, e); });
  // read_block(handle, "#"): groupby on (row.strip()[:1] == "#")
  auto is_sig = [](const std::string& s) { std::string t = strip(s); return !t.empty() && t[0] == '#'; };
  bool found = false;
  size_t i = 0, n = lines.size();
  while (i < n) {
    if (!is_sig(lines[i])) { ++i; continue; }  // contents before the first header are skipped
    size_t j = i;
    while (j < n && is_sig(lines[j])) ++j;  // run of headers: all but the last yield empty blocks
    for (size_t k = i; k + 1 < j; ++k) ac.blocks.emplace_back();
    found = true;
    if (j >= n) {
      // header at EOF: `next(it)` raises StopIteration inside the generator -> RuntimeError
      g_err = "anchors file ends with a header line (reference raises RuntimeError)";
      return false;
    }
    size_t k = j;
    while (k < n && !is_sig(lines[k])) ++k;
    std::vector<std::vector<std::string>> blk;
    for (size_t t = j; t < k; ++t) blk.push_back(split_ws(strip(lines[t])));
    ac.blocks.push_back(std::move(blk));
    i = k;
  }
  if (!found) {  // formats/base.py:635-638: whole file is one block
    std::vector<std::vector<std::string>> blk;
    for (auto& l : lines) blk.push_back(split_ws(strip(l)));
    ac.blocks.push_back(std::move(blk));
  }
  return true;
}

std::string int_str(float score) {  // str(int(b.score)): truncation toward zero
  char buf[64];
  double d = (double)score;
  if (std::fabs(d) < 9e18) snprintf(buf, sizeof buf, "%lld", (long long)d);
  else snprintf(buf, sizeof buf, "%.0f", std::trunc(d));
  return buf;
}

// ---------------------------------------------------------------------------------
// cKDTree (third-party, restated; scipy ckdtree build.cxx / query.cxx semantics,
// balanced_tree=True, compact_nodes=True, leafsize=16, k=1, p=1, eps=0)
// ---------------------------------------------------------------------------------
struct KDNode { int split_dim; double split; int lesser, greater; int start, end; };

struct KDTree {
  int n = 0;
  std::vector<double> data;        // n x 2
  std::vector<int> indices;
  std::vector<KDNode> nodes;
  double mins[2], maxes[2];
  static constexpr int kLeaf = 16;

  int build(int start, int end) {
    int me = (int)nodes.size();
    nodes.push_back(KDNode{-1, 0.0, -1, -1, start, end});
    if (end - start <= kLeaf) return me;
    double mx[2], mn[2];
    for (int d = 0; d < 2; ++d) { mx[d] = mn[d] = data[(size_t)indices[start] * 2 + d]; }
    for (int i = start + 1; i < end; ++i)
      for (int d = 0; d < 2; ++d) {
        double v = data[(size_t)indices[i] * 2 + d];
        if (v > mx[d]) mx[d] = v;
        if (v < mn[d]) mn[d] = v;
      }
    int d = 0; double size = 0;
    for (int i = 0; i < 2; ++i) if (mx[i] - mn[i] > size) { d = i; size = mx[i] - mn[i]; }
    if (mx[d] == mn[d]) return me;  // all points identical
    int npts = end - start;
    const double* dat = data.data();
    std::nth_element(indices.begin() + start, indices.begin() + start + npts / 2, indices.begin() + end,
                     [dat, d](int a, int b) { return dat[(size_t)a * 2 + d] < dat[(size_t)b * 2 + d]; });
    double split = dat[(size_t)indices[start + npts / 2] * 2 + d];
    int p = start, q = end - 1;
    while (p <= q) {
      if (dat[(size_t)indices[p] * 2 + d] < split) ++p;
      else if (dat[(size_t)indices[q] * 2 + d] >= split) --q;
      else { std::swap(indices[p], indices[q]); ++p; --q; }
    }
    if (p == start) {  // nothing < split: slide
      int j = start; split = dat[(size_t)indices[j] * 2 + d];
      for (int i = start + 1; i < end; ++i)
        if (dat[(size_t)indices[i] * 2 + d] < split) { j = i; split = dat[(size_t)indices[j] * 2 + d]; }
      std::swap(indices[start], indices[j]);
      p = start + 1;
    } else if (p == end) {  // nothing >= split
      int j = end - 1; split = dat[(size_t)indices[j] * 2 + d];
      for (int i = start; i < end - 1; ++i)
        if (dat[(size_t)indices[i] * 2 + d] > split) { j = i; split = dat[(size_t)indices[j] * 2 + d]; }
      std::swap(indices[end - 1], indices[j]);
      p = end - 1;
    }
    nodes[me].split_dim = d;
    nodes[me].split = split;
    int l = build(start, p);
    int g = build(p, end);
    nodes[me].lesser = l;
    nodes[me].greater = g;
    return me;
  }

  void init(const std::vector<std::pair<long, long>>& pts) {
    n = (int)pts.size();
    data.resize((size_t)n * 2);
    indices.resize(n);
    for (int i = 0; i < n; ++i) { data[2 * i] = (double)pts[i].first; data[2 * i + 1] = (double)pts[i].second; indices[i] = i; }
    for (int d = 0; d < 2; ++d) { mins[d] = maxes[d] = n ? data[d] : 0; }
    for (int i = 1; i < n; ++i) for (int d = 0; d < 2; ++d) {
      mins[d] = std::min(mins[d], data[2 * i + d]);
      maxes[d] = std::max(maxes[d], data[2 * i + d]);
    }
    nodes.clear();
    if (n) build(0, n);
  }

  struct HeapItem { double prio; int node; double sd[2]; };
  // returns index of the nearest point with L1 distance < bound (strict), or n; *dist_out set.
  int query(double x0, double x1, double bound, double* dist_out) const {
    if (!n) { *dist_out = INFINITY; return n; }
    std::vector<HeapItem> heap;
    auto push = [&](const HeapItem& it) {
      heap.push_back(it);
      size_t i = heap.size() - 1;
      while (i > 0) {
        size_t par = (i - 1) / 2;
        if (heap[i].prio < heap[par].prio) { std::swap(heap[i], heap[par]); i = par; } else break;
      }
    };
    auto pop = [&]() {
      HeapItem top = heap[0];
      heap[0] = heap.back();
      heap.pop_back();
      size_t nn = heap.size(), i = 0;
      // scipy heap::remove(): sift-down picking the smaller child, swap only if parent > child
      for (;;) {
        size_t j = 2 * i + 1, k = 2 * i + 2, l;
        if (j < nn && heap[i].prio > heap[j].prio) l = j; else l = i;
        if (k < nn && heap[l].prio > heap[k].prio) l = k;
        if (l == i) break;
        std::swap(heap[i], heap[l]);
        i = l;
      }
      return top;
    };
    const double x[2] = {x0, x1};
    HeapItem cur;
    cur.node = 0; cur.prio = 0;
    for (int d = 0; d < 2; ++d) {
      double s = std::max(0.0, std::max(mins[d] - x[d], x[d] - maxes[d]));
      cur.sd[d] = s; cur.prio += s;
    }
    double best = bound; int besti = n;
    for (;;) {
      const KDNode& nd = nodes[cur.node];
      if (nd.split_dim == -1) {
        for (int i = nd.start; i < nd.end; ++i) {
          int id = indices[i];
          double dd = std::fabs(data[2 * (size_t)id] - x[0]) + std::fabs(data[2 * (size_t)id + 1] - x[1]);
          if (dd < best) { best = dd; besti = id; }
        }
        if (heap.empty()) break;
        cur = pop();
      } else {
        if (cur.prio > best) break;  // eps = 0 -> epsfac 1
        int d = nd.split_dim;
        HeapItem nearI = cur, farI = cur;
        if (x[d] < nd.split) { nearI.node = nd.lesser; farI.node = nd.greater; }
        else { nearI.node = nd.greater; farI.node = nd.lesser; }
        double nsd = std::fabs(nd.split - x[d]);
        farI.prio = cur.prio + nsd - cur.sd[d];
        farI.sd[d] = nsd;
        if (farI.prio <= best) push(farI);
        cur = nearI;
      }
    }
    *dist_out = (besti == n) ? INFINITY : best;
    return besti;
  }
};

// ---------------------------------------------------------------------------------
// shared: BED pair + self detection (src/jcvi/compara/synteny.py:491-503)
// ---------------------------------------------------------------------------------
struct Beds { Bed q, s; bool is_self = false; };

bool load_beds(const char* qbed, const char* sbed, Beds& b) {
  b.is_self = std::string(qbed) == std::string(sbed);  // string equality of the paths
  if (!load_bed(qbed, b.q)) return false;
  if (!load_bed(sbed, b.s)) return false;
  return true;
}

std::string pair_key(const std::string& a, const std::string& b) { return a + '\t' + b; }

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// -----------------------------------------------------------------------------------
// blastfilter  (src/jcvi/compara/blastfilter.py:35-161 + helpers 164-284)
// stats[0..5] = total parsed lines, after join/dedupe, after exclude, after cscore,
//               after tandem, number of "not in bed" warnings
// qdups_path / sdups_path / qnewbed / snewbed: the --tandems_only side outputs (may be NULL)
// -----------------------------------------------------------------------------------
int orc_blastfilter(const char* blast_file, const char* qbed, const char* sbed, int strip_names,
                    double cscore, int tandem_Nmax, const char* exclude, const char* out_path,
                    const char* qdups_path, const char* sdups_path, const char* qnewbed,
                    const char* snewbed, long* stats) {
  Beds beds;
  if (!load_beds(qbed, sbed, beds)) return -1;
  std::vector<BlastLine> blasts;
  if (!load_blast(blast_file, blasts)) return -1;
  if (stats) stats[0] = (long)blasts.size();
  // blastfilter.py:49  sorted(list(bl), key=score, reverse=True): stable, ties keep file order
  {
    std::vector<uint32_t> ord(blasts.size());
    for (size_t i = 0; i < ord.size(); ++i) ord[i] = (uint32_t)i;
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return blasts[a].score > blasts[b].score; });
    std::vector<BlastLine> sorted_b;
    sorted_b.reserve(blasts.size());
    for (uint32_t i : ord) sorted_b.push_back(blasts[i]);
    blasts.swap(sorted_b);
  }

  const bool is_self = beds.is_self;
  std::vector<BlastLine> filtered;
  std::vector<std::string> fq, fs;      // current b.query / b.subject names
  std::vector<std::string> fqseq, fsseq;
  std::unordered_set<std::string> seen;
  long nwarn = 0;
  for (auto& b : blasts) {                                   // blastfilter.py:55-95
    std::string query = b.query, subject = b.subject;
    if (is_self && query == subject) continue;
    if (strip_names) { query = gene_name(query); subject = gene_name(subject); }
    auto qit = beds.q.order.find(query);
    if (qit == beds.q.order.end()) { ++nwarn; continue; }
    auto sit = beds.s.order.find(subject);
    if (sit == beds.s.order.end()) { ++nwarn; continue; }
    int qi = qit->second, si = sit->second;
    const Bed* qb = &beds.q; const Bed* sb = &beds.s;
    if (is_self && qi > si) { std::swap(query, subject); std::swap(qi, si); std::swap(qb, sb); }
    std::string key = pair_key(query, subject);
    if (!seen.insert(key).second) continue;
    b.qi = qi; b.si = si;
    filtered.push_back(b);
    fq.push_back(query); fs.push_back(subject);
    fqseq.push_back(qb->rows[qi].seqid); fsseq.push_back(sb->rows[si].seqid);
  }
  if (stats) { stats[1] = (long)filtered.size(); stats[5] = nwarn; }

  auto compact = [&](const std::vector<char>& keep) {
    size_t w = 0;
    for (size_t i = 0; i < filtered.size(); ++i)
      if (keep[i]) {
        if (w != i) { filtered[w] = filtered[i]; fq[w] = fq[i]; fs[w] = fs[i]; fqseq[w] = fqseq[i]; fsseq[w] = fsseq[i]; }
        ++w;
      }
    filtered.resize(w); fq.resize(w); fs.resize(w); fqseq.resize(w); fsseq.resize(w);
  };

  if (exclude && *exclude) {                                 // blastfilter.py:205-222
    AnchorBlocks ac;
    if (!load_anchors(exclude, ac)) return -1;
    std::unordered_set<std::string> ex;
    for (auto& blk : ac.blocks) for (auto& row : blk) {
      if (row.size() < 2) { g_err = "anchors row with <2 columns"; return -1; }
      ex.insert(pair_key(row[0], row[1])); ex.insert(pair_key(row[1], row[0]));
    }
    std::vector<char> keep(filtered.size());
    for (size_t i = 0; i < filtered.size(); ++i) keep[i] = !ex.count(pair_key(fq[i], fs[i]));
    compact(keep);
  }
  if (stats) stats[2] = (long)filtered.size();

  if (cscore != 0.0) {                                       // blastfilter.py:225-237
    std::unordered_map<std::string, double> best;            // defaultdict(float)
    for (size_t i = 0; i < filtered.size(); ++i) {
      double sc = (double)filtered[i].score;
      double& bq = best[fq[i]]; if (sc > bq) bq = sc;
      double& bs = best[fs[i]]; if (sc > bs) bs = sc;
    }
    std::vector<char> keep(filtered.size());
    for (size_t i = 0; i < filtered.size(); ++i) {
      double den = std::max(best[fq[i]], best[fs[i]]);
      if (den == 0.0) { g_err = "ZeroDivisionError: float division by zero (score <= 0 with cscore on)"; return -2; }
      double cur = (double)filtered[i].score / den;
      keep[i] = cur > cscore;
    }
    compact(keep);
  }
  if (stats) stats[3] = (long)filtered.size();

  if (tandem_Nmax) {                                         // blastfilter.py:113-156
    auto tandem_grouper = [&](bool flip) {                   // blastfilter.py:262-284
      struct T { const std::string* name; const std::string* seqid; int rank; };
      std::vector<T> simple;
      for (size_t i = 0; i < filtered.size(); ++i) {
        if (!(filtered[i].evalue < 1e-10)) continue;
        if (!flip) simple.push_back({&fq[i], &fsseq[i], filtered[i].si});
        else simple.push_back({&fs[i], &fqseq[i], filtered[i].qi});
      }
      std::sort(simple.begin(), simple.end(), [](const T& a, const T& b) {
        int c = a.name->compare(*b.name); if (c) return c < 0;
        c = a.seqid->compare(*b.seqid); if (c) return c < 0;
        return a.rank < b.rank;
      });
      Grouper g;
      for (size_t i = 0; i + 1 < simple.size(); ++i) {
        const T& a = simple[i]; const T& b = simple[i + 1];
        if (*a.name != *b.name) continue;
        if (b.rank - a.rank <= tandem_Nmax && *b.seqid == *a.seqid) g.join(a.rank, b.rank);
      }
      return g;
    };
    Grouper qt = tandem_grouper(true), st = tandem_grouper(false);
    auto write_localdups = [&](const Grouper& g, const Bed& bed, const char* path,
                               std::unordered_map<std::string, std::string>& d2m) -> bool {  // 164-185
      std::vector<std::vector<std::string>> tgroups;
      for (auto& grp : g.iter()) {
        std::vector<const BedLine*> rows;
        for (long r : grp) rows.push_back(&bed.rows[(size_t)r]);
        std::stable_sort(rows.begin(), rows.end(), [](const BedLine* a, const BedLine* b) {
          long la = -std::labs(a->end - a->start), lb = -std::labs(b->end - b->start);
          if (la != lb) return la < lb;
          return a->accn < b->accn;
        });
        std::vector<std::string> accns;
        for (auto* r : rows) accns.push_back(r->accn);
        tgroups.push_back(std::move(accns));
      }
      std::sort(tgroups.begin(), tgroups.end());
      FILE* fh = nullptr;
      if (path && *path) { fh = fopen(path, "w"); if (!fh) { g_err = "cannot write localdups"; return false; } }
      for (auto& accns : tgroups) {
        if (fh) {
          for (size_t i = 0; i < accns.size(); ++i) { if (i) fputc('\t', fh); fputs(accns[i].c_str(), fh); }
          fputc('\n', fh);
        }
        for (size_t i = 1; i < accns.size(); ++i) d2m[accns[i]] = accns[0];
      }
      if (fh) fclose(fh);
      return true;
    };
    auto write_new_bed = [&](const Bed& bed, const std::unordered_map<std::string, std::string>& children,
                             const char* path) -> bool {      // blastfilter.py:188-197
      if (!path || !*path) return true;
      FILE* fh = fopen(path, "w");
      if (!fh) { g_err = "cannot write nolocaldups bed"; return false; }
      for (auto& r : bed.rows) {
        if (children.count(r.accn)) continue;
        // BedLine.__str__ (bed.py:75-89): seqid, start-1, end, accn, [score, strand, extra...]
        fprintf(fh, "%s\t%ld\t%ld\t%s", r.seqid.c_str(), r.start - 1, r.end, r.accn.c_str());
        for (size_t k = 4; k < r.args.size(); ++k) fprintf(fh, "\t%s", r.args[k].c_str());
        fputc('\n', fh);
      }
      fclose(fh);
      return true;
    };
    std::unordered_map<std::string, std::string> qd2m, sd2m;
    if (is_self) {                                           // blastfilter.py:127-131
      for (auto& s : st.iter()) {
        std::vector<long> rest(s.begin() + 1, s.end());
        qt.join(s[0], rest);
      }
      if (!write_localdups(qt, beds.q, qdups_path, qd2m)) return -1;
      sd2m = qd2m;
      if (!write_new_bed(beds.q, qd2m, qnewbed)) return -1;
    } else {
      if (!write_localdups(qt, beds.q, qdups_path, qd2m)) return -1;
      if (!write_localdups(st, beds.s, sdups_path, sd2m)) return -1;
      if (!write_new_bed(beds.q, qd2m, qnewbed)) return -1;
      if (!write_new_bed(beds.s, sd2m, snewbed)) return -1;
    }
    // filter_tandem (blastfilter.py:240-259)
    for (size_t i = 0; i < filtered.size(); ++i) {
      auto a = qd2m.find(fq[i]); if (a != qd2m.end()) fq[i] = a->second;
      auto b = sd2m.find(fs[i]); if (b != sd2m.end()) fs[i] = b->second;
    }
    // mother_blast.sort(key=score, reverse=True) is a stable no-op on an already sorted list
    std::unordered_set<std::string> seen2;
    std::vector<char> keep(filtered.size());
    for (size_t i = 0; i < filtered.size(); ++i) {
      keep[i] = 0;
      if (fq[i] == fs[i]) continue;
      if (!seen2.insert(pair_key(fq[i], fs[i])).second) continue;
      keep[i] = 1;
    }
    compact(keep);
  }
  if (stats) stats[4] = (long)filtered.size();

  FILE* fw = fopen(out_path, "w");                            // blastfilter.py:158-161, 200-202
  if (!fw) { g_err = "cannot write filtered file"; return -1; }
  char buf[1024];
  for (size_t i = 0; i < filtered.size(); ++i) {
    const BlastLine& b = filtered[i];
    snprintf(buf, sizeof buf, kBlastOutput, fq[i].c_str(), fs[i].c_str(), b.pctid, b.hitlen, b.nmismatch, b.ngaps,
             b.qstart, b.qstop, b.sstart, b.sstop, b.evalue, b.score);
    fputs(buf, fw); fputc('\n', fw);
  }
  fclose(fw);
  return 0;
}

// -----------------------------------------------------------------------------------
// read_blast semantics shared by scan and liftover
// (src/jcvi/compara/synteny.py:255-297 and 300-374)
// -----------------------------------------------------------------------------------
}  // extern "C"

namespace {
struct Hit { int qi, si; float score; };

bool read_blast_ordered(const char* blast_file, const Beds& beds, bool ostrip, std::vector<Hit>& hits, long* nparsed) {
  std::vector<BlastLine> blasts;
  if (!load_blast(blast_file, blasts)) return false;
  if (nparsed) *nparsed = (long)blasts.size();
  std::unordered_set<std::string> seen;
  for (auto& b : blasts) {
    std::string query = b.query, subject = b.subject;
    if (beds.is_self && query == subject) continue;
    if (ostrip) { query = gene_name(query); subject = gene_name(subject); }
    auto qit = beds.q.order.find(query);
    auto sit = beds.s.order.find(subject);
    if (qit == beds.q.order.end() || sit == beds.s.order.end()) continue;
    int qi = qit->second, si = sit->second;
    if (beds.is_self) {
      if (qi > si) { std::swap(query, subject); std::swap(qi, si); }
      // after the swap q/s BedLines are swapped too; same BED in self mode
      if (beds.q.rows[qi].seqid == beds.s.rows[si].seqid && si - qi < 40) continue;
    }
    if (!seen.insert(pair_key(query, subject)).second) continue;
    hits.push_back({qi, si, b.score});
  }
  return true;
}
}  // namespace

extern "C" {

// -----------------------------------------------------------------------------------
// scan  (src/jcvi/compara/synteny.py:1849-1904; batch_scan 437-452; synteny_scan 402-434)
// returns 0 ok, 1 = zero anchors (reference raises ValueError after writing the file), <0 error
// stats[0..2] = hits imported, clusters, anchors
// -----------------------------------------------------------------------------------
int orc_scan(const char* blast_file, const char* qbed, const char* sbed, int dist, int min_size,
             int intrabound, const char* anchor_path, long* stats) {
  Beds beds;
  if (!load_beds(qbed, sbed, beds)) return -1;
  std::vector<Hit> hits;
  if (!read_blast_ordered(blast_file, beds, /*ostrip=*/false, hits, nullptr)) return -1;
  if (stats) stats[0] = (long)hits.size();
  // group_hits (synteny.py:239-252)
  std::map<std::pair<std::string, std::string>, std::vector<Hit>> groups;
  for (auto& h : hits) groups[{beds.q.rows[h.qi].seqid, beds.s.rows[h.si].seqid}].push_back(h);
  FILE* fw = fopen(anchor_path, "w");
  if (!fw) { g_err = "cannot write anchors"; return -1; }
  long nclusters = 0, nanchors = 0;
  const bool is_self = beds.is_self;
  for (auto& kv : groups) {  // sorted(chr_pair_points.keys())
    std::vector<Hit>& pts = kv.second;
    std::sort(pts.begin(), pts.end(), [](const Hit& a, const Hit& b) {
      if (a.qi != b.qi) return a.qi < b.qi;
      if (a.si != b.si) return a.si < b.si;
      return a.score < b.score;
    });
    Grouper clusters;
    int n = (int)pts.size();
    for (int i = 0; i < n; ++i) {
      for (int j = i - 1; j >= 0; --j) {
        int del_x = pts[i].qi - pts[j].qi;
        if (del_x > dist) break;
        int del_y = pts[i].si - pts[j].si;
        if (std::abs(del_y) > dist) continue;
        if (is_self) {
          int intradist = std::min(std::abs(pts[i].qi - pts[i].si), std::abs(pts[j].qi - pts[j].si));
          if (intradist < intrabound) continue;
        }
        clusters.join(i, j);
      }
    }
    for (auto& c : clusters.iter()) {
      std::unordered_set<int> xs, ys;
      for (long m : c) { xs.insert(pts[m].qi); ys.insert(pts[m].si); }
      if ((int)std::min(xs.size(), ys.size()) < min_size) continue;   // _score (synteny.py:214-219)
      std::vector<long> members = c;
      std::sort(members.begin(), members.end());
      fputs("###\n", fw);
      for (long m : members) {
        fprintf(fw, "%s\t%s\t%s\n", beds.q.rows[pts[m].qi].accn.c_str(), beds.s.rows[pts[m].si].accn.c_str(),
                int_str(pts[m].score).c_str());
        ++nanchors;
      }
      ++nclusters;
    }
  }
  fclose(fw);
  if (stats) { stats[1] = nclusters; stats[2] = nanchors; }
  return nclusters == 0 ? 1 : 0;
}

// -----------------------------------------------------------------------------------
// liftover  (src/jcvi/compara/synteny.py:1919-1971; synteny_liftover 455-473;
//            read_anchors 377-399; AnchorFile.filter_blocks/print_to_file base.py:52-98)
// returns 0 ok, 1 = zero anchors after filtering (ValueError in summary), <0 error
// stats[0..4] = hits imported, anchors imported, lifted, removed, corrected
// -----------------------------------------------------------------------------------
int orc_liftover(const char* blast_file, const char* anchor_file, const char* qbed, const char* sbed,
                 int strip_names, int dist, const char* out_path, long* stats) {
  Beds beds;
  if (!load_beds(qbed, sbed, beds)) return -1;
  std::vector<Hit> hits;
  if (!read_blast_ordered(blast_file, beds, strip_names != 0, hits, nullptr)) return -1;
  if (stats) stats[0] = (long)hits.size();
  std::unordered_map<std::string, std::string> accepted;   // (query,subject) -> str(int(score))
  std::unordered_map<uint64_t, std::string> blast_to_score;
  std::map<std::pair<std::string, std::string>, std::vector<Hit>> all_hits;
  for (auto& h : hits) {
    std::string sc = int_str(h.score);
    // accepted is keyed by the (possibly swapped) names, which are the BED accns of the ranks
    accepted[pair_key(beds.q.rows[h.qi].accn, beds.s.rows[h.si].accn)] = sc;
    blast_to_score[((uint64_t)(uint32_t)h.qi << 32) | (uint32_t)h.si] = sc;
    all_hits[{beds.q.rows[h.qi].seqid, beds.s.rows[h.si].seqid}].push_back(h);
  }
  AnchorBlocks ac;
  if (!load_anchors(anchor_file, ac)) return -1;
  // read_anchors
  std::map<std::pair<std::string, std::string>, std::vector<std::pair<long, long>>> all_anchors;
  std::unordered_map<uint64_t, int> anchor_to_block;
  long nanch = 0;
  for (size_t bi = 0; bi < ac.blocks.size(); ++bi)
    for (auto& row : ac.blocks[bi]) {
      if (row.size() < 2) { g_err = "anchors row with <2 columns"; return -1; }
      auto qit = beds.q.order.find(row[0]);
      auto sit = beds.s.order.find(row[1]);
      if (qit == beds.q.order.end() || sit == beds.s.order.end()) continue;
      int qi = qit->second, si = sit->second;
      all_anchors[{beds.q.rows[qi].seqid, beds.s.rows[si].seqid}].push_back({qi, si});
      anchor_to_block[((uint64_t)(uint32_t)qi << 32) | (uint32_t)si] = (int)bi;
      ++nanch;
    }
  if ((size_t)nanch != anchor_to_block.size()) { g_err = "AssertionError: duplicate anchors"; return -3; }
  if (stats) stats[1] = nanch;
  long lifted = 0;
  for (auto& kv : all_anchors) {   // sorted(all_anchors.keys())
    auto hit_it = all_hits.find(kv.first);
    if (hit_it == all_hits.end() || hit_it->second.empty()) continue;
    KDTree tree;
    tree.init(kv.second);
    for (auto& h : hit_it->second) {
      double d;
      int idx = tree.query((double)h.qi, (double)h.si, (double)dist, &d);
      if (idx == tree.n) continue;
      if (d == 0) continue;
      auto& near = kv.second[(size_t)idx];
      int block = anchor_to_block[((uint64_t)(uint32_t)near.first << 32) | (uint32_t)near.second];
      std::string sc = blast_to_score[((uint64_t)(uint32_t)h.qi << 32) | (uint32_t)h.si] + "L";
      ac.blocks[(size_t)block].push_back({beds.q.rows[h.qi].accn, beds.s.rows[h.si].accn, sc});
      ++lifted;
    }
  }
  if (stats) stats[2] = lifted;
  long nremoved = 0, ncorrected = 0;
  if (!accepted.empty()) {   // filter_blocks (base.py:52-83)
    std::vector<std::vector<std::vector<std::string>>> nb;
    for (auto& blk : ac.blocks) {
      std::vector<std::vector<std::string>> keep;
      for (auto& line : blk) {
        if (line.size() != 3) { g_err = "ValueError: anchors line does not have 3 columns"; return -4; }
        auto it = accepted.find(pair_key(line[0], line[1]));
        if (it == accepted.end()) { ++nremoved; continue; }
        std::string score = line[2];
        if (score != it->second && score != it->second + "L") { score = it->second; ++ncorrected; }
        keep.push_back({line[0], line[1], score});
      }
      if (!keep.empty()) nb.push_back(std::move(keep));
    }
    ac.blocks.swap(nb);
  }
  if (stats) { stats[3] = nremoved; stats[4] = ncorrected; }
  FILE* fw = fopen(out_path, "w");
  if (!fw) { g_err = "cannot write lifted anchors"; return -1; }
  long total = 0;
  for (auto& blk : ac.blocks) {   // print_to_file (base.py:85-98)
    fputs("###\n", fw);
    for (auto& line : blk) {
      if (line.size() != 3) { fclose(fw); g_err = "ValueError: anchors line does not have 3 columns"; return -4; }
      fprintf(fw, "%s\t%s\t%s\n", line[0].c_str(), line[1].c_str(), line[2].c_str());
      ++total;
    }
  }
  fclose(fw);
  return total == 0 ? 1 : 0;
}

// -----------------------------------------------------------------------------------
// Small probes used by unit tests
// -----------------------------------------------------------------------------------
// cKDTree restatement: build on n integer points, query m points; outputs tree.indices,
// nearest index (n when none) per query.
int orc_kdtree(const long* anchors, int n, const long* queries, int m, double bound, int* indices_out,
               int* nearest_out, double* dist_out) {
  std::vector<std::pair<long, long>> pts((size_t)n);
  for (int i = 0; i < n; ++i) pts[(size_t)i] = {anchors[2 * i], anchors[2 * i + 1]};
  KDTree t;
  t.init(pts);
  for (int i = 0; i < n; ++i) indices_out[i] = t.indices[(size_t)i];
  for (int i = 0; i < m; ++i) {
    double d;
    nearest_out[i] = t.query((double)queries[2 * i], (double)queries[2 * i + 1], bound, &d);
    dist_out[i] = d;
  }
  return 0;
}

// gene_name probe
int orc_gene_name(const char* in, char* out, int cap) {
  std::string r = gene_name(in);
  if ((int)r.size() + 1 > cap) return -1;
  memcpy(out, r.c_str(), r.size() + 1);
  return 0;
}

// natsort compare probe
int orc_natcmp(const char* a, const char* b) { return natcmp(a, b); }

// BED order probe: writes the sorted accns, one per line
int orc_bed_order(const char* bed_path, const char* out_path) {
  Bed bed;
  if (!load_bed(bed_path, bed)) return -1;
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  for (auto& r : bed.rows) fprintf(fw, "%s\n", r.accn.c_str());
  fclose(fw);
  return 0;
}

// BlastLine parse -> __str__ round trip probe (cblast.pyx:106-120,144-156)
int orc_blastline_str(const char* line, char* out, int cap) {
  BlastLine bl; std::string scratch;
  if (!parse_blast_line(line, line + strlen(line), bl, scratch)) return -1;
  snprintf(out, (size_t)cap, kBlastOutput, bl.query, bl.subject, bl.pctid, bl.hitlen, bl.nmismatch, bl.ngaps,
           bl.qstart, bl.qstop, bl.sstart, bl.sstop, bl.evalue, bl.score);
  return 0;
}

// ---------------------------------------------------------------------------------
// formats.blast cscore  (src/jcvi/formats/blast.py:1098-1189): two passes over the table with
// dictionaries keyed by the (gene_name-stripped) names themselves; no BED.
//   pass 1 (1146-1157): best_score[name] = max score over every line naming it (starting at 0.0)
//   pass 2 (1159-1172): s = score / max(best[q], best[s]); keep s > cutoff; per (q, s) the largest s,
//                       the first line on ties
//   output (1174-1186): sorted by (query, subject); "{:.2f}" of s, optionally "{:.1f}" of pctid;
//                       --writeblast prints str(BlastLine) of the chosen line
// returns 0, -1 on I/O error, -2 on an unparsable line (128+ byte name), -3 ZeroDivisionError
// ---------------------------------------------------------------------------------
int orc_cscore(const char* blast_file, const char* out_path, const char* blast_out_path, double cutoff, int pct,
               int strip_names) {
  std::vector<BlastLine> lines;
  if (!load_blast(blast_file, lines)) return g_err.empty() ? -1 : -2;
  auto nm = [&](const char* raw) { return strip_names ? gene_name(raw) : std::string(raw); };
  std::map<std::string, double> best;
  for (const BlastLine& b : lines) {
    const std::string q = nm(b.query), s = nm(b.subject);
    const double score = (double)b.score;
    double& bq = best[q];                 // defaultdict(float): the lookup creates the key
    if (score > bq) bq = score;
    double& bs = best[s];
    if (score > bs) bs = score;
  }
  struct Pick { double s; float pctid; const BlastLine* b; };
  std::map<std::pair<std::string, std::string>, Pick> pairs;   // std::map iterates in sorted(tuple) order
  for (const BlastLine& b : lines) {
    const std::string q = nm(b.query), s = nm(b.subject);
    const double den = std::max(best[q], best[s]);
    if (den == 0.0) { g_err = "float division by zero"; return -3; }
    const double c = (double)b.score / den;
    if (c > cutoff) {
      auto key = std::make_pair(q, s);
      auto it = pairs.find(key);
      if (it == pairs.end() || c > it->second.s) pairs[key] = Pick{c, b.pctid, &b};
    }
  }
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  FILE* fwb = blast_out_path && *blast_out_path ? fopen(blast_out_path, "w") : nullptr;
  char buf[1024];
  for (const auto& kv : pairs) {
    fprintf(fw, "%s\t%s\t%.2f", kv.first.first.c_str(), kv.first.second.c_str(), kv.second.s);
    if (pct) fprintf(fw, "\t%.1f", (double)kv.second.pctid);
    fputc('\n', fw);
    if (fwb) {
      const BlastLine& bl = *kv.second.b;
      snprintf(buf, sizeof buf, kBlastOutput, bl.query, bl.subject, bl.pctid, bl.hitlen, bl.nmismatch, bl.ngaps, bl.qstart,
               bl.qstop, bl.sstart, bl.sstop, bl.evalue, bl.score);
      fprintf(fwb, "%s\n", buf);
    }
  }
  fclose(fw);
  if (fwb) fclose(fwb);
  return 0;
}

// ---------------------------------------------------------------------------------
// formats.blast filter  (src/jcvi/formats/blast.py:620-709): one pass, one predicate per row.
//   ids_path: file of ids to retain ('#' lines skipped, commas are separators, blast.py:655-662) or NULL
// Output rows are row.rstrip() + "\n" (ASCII white space and 0x1c-0x1f; non-ASCII Unicode spaces are
// not reproduced).  returns 0, -1 I/O error, -2 unparsable row (128+ byte name)
// ---------------------------------------------------------------------------------
int orc_blast_filter(const char* blast_file, const char* out_path, const char* ids_path, double score, double pctid,
                     long hitlen, double evalue, int noself, int inverse) {
  std::string txt;
  if (!read_file(blast_file, txt)) return -1;
  std::unordered_set<std::string> ids;
  const bool use_ids = ids_path && *ids_path;
  if (use_ids) {
    std::string itxt;
    if (!read_file(ids_path, itxt)) return -1;
    for_each_line(itxt.data(), itxt.size(), [&](const char* b, const char* e) {
      if (b < e && *b == '#') return;
      std::string row(b, e);
      for (char& c : row) if (c == ',') c = '\t';
      for (const std::string& t : split_ws(row)) ids.insert(t);
    });
  }
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  std::string scratch;
  bool ok = true;
  for_each_line(txt.data(), txt.size(), [&](const char* b, const char* e) {
    if (!ok) return;
    if (b < e && *b == '#') return;
    BlastLine c;
    if (!parse_blast_line(b, e, c, scratch)) { ok = false; return; }
    bool noids = false;
    if (use_ids && ids.size()) noids = !(ids.count(c.query) && ids.count(c.subject));
    bool remove = (double)c.score < score || (double)c.pctid < pctid || (long)c.hitlen < hitlen || c.evalue > evalue || noids;
    if (inverse) remove = !remove;
    remove = remove || (noself && strcmp(c.query, c.subject) == 0);
    if (!remove) {
      const char* r = e;
      while (r > b && (isspace((unsigned char)r[-1]) || ((unsigned char)r[-1] >= 0x1c && (unsigned char)r[-1] <= 0x1f))) --r;
      fwrite(b, 1, (size_t)(r - b), fw);
      fputc('\n', fw);
    }
  });
  fclose(fw);
  return ok ? 0 : -2;
}


// ---------------------------------------------------------------------------------
// synteny mcscan  (src/jcvi/compara/synteny.py:1538-1652): stack the anchor blocks on a reference BED.
//   AnchorFile.make_ranges / make_range / get_best_pair   src/jcvi/compara/base.py:29-50,139-164
//   range_chain / _make_endpoints                           src/jcvi/utils/range.py:347-355,412-463
// tracks_path: --trackids log (NULL: none); tandem_path: --mergetandem file (NULL: none).
// returns 0, -1 I/O, -4 KeyError (gene of a block not in the BED), -5 ValueError (score not an integer, empty block)
// ---------------------------------------------------------------------------------
struct McRange { double start, end; long score; int id; };

static bool mc_int(const std::string& t0, long& v) {          // int(t[:-1]) if t[-1] == "L" else int(t)
  std::string t = t0;
  if (!t.empty() && t.back() == 'L') t.pop_back();
  if (t.empty()) return false;
  char* endp = nullptr;
  v = strtol(t.c_str(), &endp, 10);
  return endp && *endp == 0 && (isdigit((unsigned char)t[0]) || ((t[0] == '-' || t[0] == '+') && t.size() > 1));
}

// range_chain on the current list; fills `chain` with indices into `ranges` in chain order
static long mc_range_chain(const std::vector<McRange>& ranges, std::vector<int>& chain) {
  struct End { double pos; int lr; int i; long score; };
  std::vector<End> ends;
  ends.reserve(2 * ranges.size());
  for (size_t i = 0; i < ranges.size(); ++i) {
    ends.push_back({ranges[i].start, 0, (int)i, ranges[i].score});
    ends.push_back({ranges[i].end, 1, (int)i, ranges[i].score});
  }
  std::sort(ends.begin(), ends.end(), [](const End& a, const End& b) {       // sorted(tuples); seqid is "0" throughout
    if (a.pos != b.pos) return a.pos < b.pos;
    if (a.lr != b.lr) return a.lr < b.lr;
    if (a.i != b.i) return a.i < b.i;
    return a.score < b.score;
  });
  struct Cell { long score; int from; int which; };
  std::vector<Cell> sc;
  sc.reserve(ends.size());
  std::vector<int> left_index(ranges.size(), -1);
  for (size_t i = 0; i < ends.size(); ++i) {
    Cell cur = i == 0 ? Cell{0, -1, -1} : sc.back();
    const End& e = ends[i];
    if (e.lr == 0) left_index[(size_t)e.i] = (int)i;
    else {
      const int lj = left_index[(size_t)e.i];
      const long cs = sc[(size_t)lj].score + e.score;
      if (cs > cur.score) cur = Cell{cs, lj, e.i};
    }
    sc.push_back(cur);
  }
  chain.clear();
  long score = sc.back().score;
  int last = sc.back().from, which = sc.back().which;
  while (last != -1) {
    if (which != -1) chain.push_back(which);
    which = sc[(size_t)last].which;
    last = sc[(size_t)last].from;
  }
  std::reverse(chain.begin(), chain.end());
  return score;
}

int orc_mcscan(const char* bed_path, const char* anchor_path, const char* out_path, int max_iter, int ascii, int clip,
               const char* tracks_path, const char* tandem_path) {
  Bed bed;
  if (!load_bed(bed_path, bed)) return -1;
  AnchorBlocks ac;
  if (!load_anchors(anchor_path, ac)) return -1;
  std::unordered_map<std::string, std::string> tandems;
  if (tandem_path && *tandem_path) {
    std::string txt;
    if (!read_file(tandem_path, txt)) return -1;
    for_each_line(txt.data(), txt.size(), [&](const char* b, const char* e) {
      std::vector<std::string> row = split_ws(std::string(b, e));
      std::string j;
      for (size_t k = 0; k < row.size(); ++k) { if (k) j += ";"; j += row[k]; }
      for (const std::string& a : row) tandems[a] = j;
    });
  }
  // make_ranges
  std::vector<McRange> ranges;
  std::vector<std::unordered_map<std::string, std::string>> block_pairs(ac.blocks.size());
  int rc = 0;
  auto make_range = [&](const std::vector<const std::string*>& q, const std::vector<const std::string*>& sv,
                        const std::vector<const std::string*>& t, int i) -> bool {
    std::unordered_map<std::string, std::pair<std::string, long>> pairs;      // get_best_pair: larger score wins, first on ties
    for (size_t k = 0; k < q.size(); ++k) {
      long tv;
      if (!mc_int(*t[k], tv)) { rc = -5; return false; }
      auto it = pairs.find(*q[k]);
      if (it == pairs.end()) pairs.emplace(*q[k], std::make_pair(*sv[k], tv));
      else if (it->second.second < tv) it->second = std::make_pair(*sv[k], tv);
    }
    for (auto& kv : pairs) block_pairs[(size_t)i][kv.first] = kv.second.first;   // dict.update
    std::vector<long> ranks;
    for (const std::string* x : q) {
      auto it = bed.order.find(*x);
      if (it == bed.order.end()) { rc = -4; return false; }
      ranks.push_back(it->second);
    }
    std::sort(ranks.begin(), ranks.end());
    double qmin = (double)ranks.front(), qmax = (double)ranks.back();
    if (ranks.back() - ranks.front() >= 2L * clip) { qmin += clip / 2.0; qmax -= clip / 2.0; }
    ranges.push_back({qmin, qmax, (long)pairs.size(), i});
    return true;
  };
  for (size_t i = 0; i < ac.blocks.size(); ++i) {
    const auto& ib = ac.blocks[i];
    if (ib.empty()) { g_err = "empty block (zip(*ib) fails in the reference)"; return -5; }
    std::vector<const std::string*> q, sv, t;
    for (const auto& row : ib) {
      if (row.size() != 3) { g_err = "anchor line without three columns"; return -5; }
      q.push_back(&row[0]); sv.push_back(&row[1]); t.push_back(&row[2]);
    }
    if (!bed.order.count(*q[0])) std::swap(q, sv);
    if (!make_range(q, sv, t, (int)i)) return rc;
    if (!bed.order.count(*sv[0])) continue;
    if (!make_range(sv, q, t, (int)i)) return rc;            // self comparison: the block seen from the other side
  }
  FILE* fl = nullptr;
  if (tracks_path && *tracks_path) { fl = fopen(tracks_path, "w"); if (!fl) return -1; }
  std::vector<std::vector<int>> tracks;                      // block ids in chain order
  int iteration = 0;
  while (!ranges.empty()) {
    if (iteration >= max_iter) break;
    std::vector<int> chain;
    mc_range_chain(ranges, chain);
    std::vector<int> ids;
    std::set<int> sel;
    for (int x : chain) { ids.push_back(ranges[(size_t)x].id); sel.insert(ranges[(size_t)x].id); }
    tracks.push_back(ids);
    if (fl) {
      bool first = true;
      for (int x : sel) { fprintf(fl, first ? "%d" : ",%d", x); first = false; }
      fputc('\n', fl);
    }
    std::vector<McRange> rest;
    for (const McRange& r : ranges) if (!sel.count(r.id)) rest.push_back(r);
    ranges.swap(rest);
    ++iteration;
  }
  if (fl) fclose(fl);
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  for (const BedLine& b : bed.rows) {
    std::string line = b.accn + "\t";
    for (size_t t = 0; t < tracks.size(); ++t) {
      std::string anchor = ".";
      for (int tid : tracks[t]) {
        auto it = block_pairs[(size_t)tid].find(b.accn);
        if (it != block_pairs[(size_t)tid].end()) { anchor = it->second; }
        if (anchor != ".") break;
      }
      if (ascii && anchor != ".") anchor = "x";
      if (tandem_path && *tandem_path) { auto it = tandems.find(anchor); if (it != tandems.end()) anchor = it->second; }
      if (t && !ascii) line += "\t";
      line += anchor;
    }
    line += "\n";
    fwrite(line.data(), 1, line.size(), fw);
  }
  fclose(fw);
  return 0;
}


// ---------------------------------------------------------------------------------
// synteny simple  (src/jcvi/compara/synteny.py:1137-1316): the block ends of every block of an anchors file.
//   get_orientation 222-236 (np.polyfit slope: restated as the sign of the exact integer covariance
//   n*sum(xy) - sum(x)*sum(y); a covariance of exactly 0 gives a slope of +-rounding noise in the reference --
//   "+" here, parity unpinned for such flat blocks), get_boundary_bases 1124-1134, range_minmax utils/range.py:157-168.
// mode bits: 1 --rich, 2 --coords, 4 --noheader, 8 --bed (implies --coords; bed_path receives <simple>.bed)
// returns 0, 1 (no blocks: the reference logs an error and returns), -1 I/O, -4 KeyError, -5 ValueError,
// -7 AssertionError (a block end on two seqids)
// ---------------------------------------------------------------------------------
int orc_simple(const char* anchor_file, const char* qbed_path, const char* sbed_path, const char* out_path, int mode,
               const char* bed_path) {
  const bool rich = mode & 1, noheader = mode & 4, bedout = mode & 8, coords = (mode & 2) || bedout;
  Beds beds;
  if (!load_beds(qbed_path, sbed_path, beds)) return -1;
  AnchorBlocks ac;
  if (!load_anchors(anchor_file, ac)) return -1;
  if (ac.blocks.empty() || ac.blocks[0].empty()) return 1;          // AnchorFile.is_empty
  std::string af(anchor_file);
  std::string pf;                                                   // "-".join(anchorfile.split(".", 2)[:2])
  {
    size_t d1 = af.find('.');
    if (d1 == std::string::npos) pf = af;
    else {
      size_t d2 = af.find('.', d1 + 1);
      pf = af.substr(0, d1) + "-" + af.substr(d1 + 1, d2 == std::string::npos ? std::string::npos : d2 - d1 - 1);
    }
  }
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  if (!noheader) {
    if (coords) fputs("Block\tChr\tStart\tEnd\tSpan\tStartGene\tEndGene\tGeneSpan\tOrientation\n", fw);
    else {
      fputs("StartGeneA\tEndGeneA\tStartGeneB\tEndGeneB\tOrientation\tScore", fw);
      if (rich) fputs("\tStartOrderA\tEndOrderA\tStartOrderB\tEndOrderB\tSizeA\tSizeB\tSize\tBlock", fw);
      fputc('\n', fw);
    }
  }
  struct BB { std::string seqid; long start, end; std::string accn; long size; char orient; };
  std::vector<BB> bbed;
  int rc = 0;
  for (size_t i = 0; i < ac.blocks.size() && !rc; ++i) {
    const auto& block = ac.blocks[i];
    if (block.empty()) { rc = -5; break; }                           // zip(*block) of nothing
    std::vector<long> ia, ib;
    for (const auto& row : block) {
      if (row.size() != 3) { rc = -5; break; }
      auto qa = beds.q.order.find(row[0]);
      auto sa = beds.s.order.find(row[1]);
      if (qa == beds.q.order.end() || sa == beds.s.order.end()) { rc = -4; break; }
      ia.push_back(qa->second); ib.push_back(sa->second);
    }
    if (rc) break;
    const long astarti = *std::min_element(ia.begin(), ia.end()), aendi = *std::max_element(ia.begin(), ia.end());
    const long bstarti = *std::min_element(ib.begin(), ib.end()), bendi = *std::max_element(ib.begin(), ib.end());
    const BedLine &as = beds.q.rows[(size_t)astarti], &ae = beds.q.rows[(size_t)aendi];
    const BedLine &bs = beds.s.rows[(size_t)bstarti], &be = beds.s.rows[(size_t)bendi];
    const size_t sizeA = std::set<long>(ia.begin(), ia.end()).size(), sizeB = std::set<long>(ib.begin(), ib.end()).size();
    const size_t size = block.size();
    char orient = '+';
    if (ia.size() >= 2) {
      __int128 n = (long long)ia.size(), sx = 0, sy = 0, sxy = 0;
      for (size_t k = 0; k < ia.size(); ++k) { sx += ia[k]; sy += ib[k]; sxy += (__int128)ia[k] * ib[k]; }
      const __int128 cov = n * sxy - sx * sy;
      orient = cov >= 0 ? '+' : '-';
    }
    const long aspan = aendi - astarti + 1, bspan = bendi - bstarti + 1;
    const long score = (long)std::pow((double)(aspan * bspan), 0.5);   // int((aspan * bspan) ** 0.5)
    const std::string block_id = pf + "-block-" + std::to_string(i);
    if (coords) {
      if (as.seqid != ae.seqid || bs.seqid != be.seqid) { rc = -7; break; }
      auto minmax = [](const BedLine& s, const BedLine& e, long& lo, long& hi) {   // range_minmax of the two genes
        lo = std::min(std::make_pair(s.start, s.end), std::make_pair(e.start, e.end)).first;
        hi = e.end > s.end ? e.end : s.end;            // max(..., key=end): the first of equal ends, same value
      };
      long alo, ahi, blo, bhi;
      minmax(as, ae, alo, ahi); minmax(bs, be, blo, bhi);
      fprintf(fw, "%s\t%s\t%ld\t%ld\t%ld\t%s\t%s\t%ld\t+\n", block_id.c_str(), as.seqid.c_str(), alo, ahi, ahi - alo + 1,
              as.accn.c_str(), ae.accn.c_str(), aspan);
      fprintf(fw, "%s\t%s\t%ld\t%ld\t%ld\t%s\t%s\t%ld\t%c\n", block_id.c_str(), bs.seqid.c_str(), blo, bhi, bhi - blo + 1,
              bs.accn.c_str(), be.accn.c_str(), bspan, orient);
      if (bedout) {
        char nm[512];
        snprintf(nm, sizeof nm, "%s:%ld-%ld", as.seqid.c_str(), alo, ahi);
        bbed.push_back({bs.seqid, blo, bhi, nm, (long)size, orient});   // BedLine: start = (blo - 1) + 1
      }
      continue;
    }
    fprintf(fw, "%s\t%s\t%s\t%s\t%ld\t%c", as.accn.c_str(), ae.accn.c_str(), bs.accn.c_str(), be.accn.c_str(), score, orient);
    if (rich) fprintf(fw, "\t%ld\t%ld\t%ld\t%ld\t%zu\t%zu\t%zu\t%s", astarti, aendi, bstarti, bendi, sizeA, sizeB, size, block_id.c_str());
    fputc('\n', fw);
  }
  fclose(fw);
  if (rc) return rc;
  if (bedout) {
    if (!bed_path || !*bed_path) return -1;
    // Bed.print_to_file(sorted=True): key (natsort_key(seqid), start, accn); BedLine.__str__ = seqid, start-1, end, accn, score, strand
    std::stable_sort(bbed.begin(), bbed.end(), [](const BB& a, const BB& b) {
      int c = natcmp(a.seqid, b.seqid);
      if (c) return c < 0;
      if (a.start != b.start) return a.start < b.start;
      return a.accn < b.accn;
    });
    FILE* fb = fopen(bed_path, "w");
    if (!fb) return -1;
    for (const BB& b : bbed) {
      long start = b.start < 1 ? 1 : b.start;                        // "Start < 1. Reset start"
      fprintf(fb, "%s\t%ld\t%ld\t%s\t%ld\t%c\n", b.seqid.c_str(), start - 1, b.end, b.accn.c_str(), b.size, b.orient);
    }
    fclose(fb);
  }
  return 0;
}


// ---------------------------------------------------------------------------------
// compara.synfind  (src/jcvi/compara/synfind.py:45-240; LIS: src/jcvi/algorithms/lis.py:20-116)
// For every gene of the query BED: the hits of the +-window//2 genes around it on the same seqid, single-linked
// along the subject when consecutive (subject order) hits are < window//2 ranks apart on one seqid; per linked
// group the closest flanker, the longer of LIS / LDS over the subject ranks ("collinear" scoring), the score
// min(#distinct x, #distinct y) (- 1 without a syntelog), kept when >= int(cutoff * window); then the mirror pass.
// `--scoring=density` leaves `orientation` unbound in the reference (UnboundLocalError): not restated.
// returns 0, -1 I/O, -5 ValueError (no hit maps onto the BEDs: `transposed([])` in the mirror pass)
// ---------------------------------------------------------------------------------
namespace {
typedef std::pair<long, long> SfPt;                         // (x, y) = (query rank, subject rank)

// lis.py:81-107 on (y, i) pairs; returns the chosen elements
std::vector<SfPt> sf_lis(const std::vector<SfPt>& xs) {
  std::vector<SfPt> tops;                                   // patience_sort: bisect_left on the pile tops
  std::vector<std::vector<std::pair<SfPt, long>>> piles(1); // dummy pile 0; entries (x, index into the previous pile)
  for (const SfPt& x : xs) {
    const size_t p = (size_t)(std::lower_bound(tops.begin(), tops.end(), x) - tops.begin());
    if (p == tops.size()) tops.push_back(x); else tops[p] = x;
    if (p + 1 == piles.size()) piles.emplace_back();
    piles[p + 1].push_back({x, (long)piles[p].size() - 1});
  }
  std::vector<SfPt> lis;
  long prev = 0;
  for (size_t pile = piles.size() - 1; pile >= 1; --pile) {
    const auto& e = piles[pile][(size_t)prev];
    lis.push_back(e.first);
    prev = e.second;
  }
  std::reverse(lis.begin(), lis.end());
  return lis;
}
std::vector<SfPt> sf_lds(const std::vector<SfPt>& xs) {     // list(reversed(LIS(reversed(xs))))
  std::vector<SfPt> r(xs.rbegin(), xs.rend());
  std::vector<SfPt> l = sf_lis(r);
  std::reverse(l.begin(), l.end());
  return l;
}
struct SfRegion { long syntelog, far_syntelog, left, right; char gray, orient; long score; };

void sf_regions(long query, const Bed& sbed, const std::vector<SfPt>& data, long window, long cutoff, std::vector<SfRegion>& regions) {
  regions.clear();
  std::vector<SfPt> ys(data);
  std::stable_sort(ys.begin(), ys.end(), [](const SfPt& a, const SfPt& b) { return a.second < b.second; });
  // consecutive links -> runs; a point without any link is in no group
  std::vector<std::vector<SfPt>> groups;
  size_t i = 0;
  while (i + 1 < ys.size()) {
    auto linked = [&](size_t k) {
      return ys[k + 1].second - ys[k].second < window && sbed.rows[(size_t)ys[k].second].seqid == sbed.rows[(size_t)ys[k + 1].second].seqid;
    };
    if (!linked(i)) { ++i; continue; }
    size_t j = i;
    while (j + 1 < ys.size() && linked(j)) ++j;
    groups.emplace_back(ys.begin() + (long)i, ys.begin() + (long)j + 1);
    i = j + 1;
  }
  std::sort(groups.begin(), groups.end());                  // sorted(g): lists of tuples, lexicographic
  for (auto& group : groups) {
    std::sort(group.begin(), group.end());                  // get_flanker: group.sort()
    const size_t pos = (size_t)(std::lower_bound(group.begin(), group.end(), SfPt(query, 0)) - group.begin());
    const SfPt left = pos == 0 ? group.front() : group[pos - 1];
    const SfPt right = pos == group.size() ? group.back() : group[pos];
    SfPt flanker, other;
    if (std::labs(query - left.first) < std::labs(query - right.first)) { flanker = left; other = right; }
    else { flanker = right; other = left; }
    const bool flanked = !(pos == 0 || pos == group.size());   // `flanker == query` compares a tuple with an int: never true
    // collinear scoring: the longer of LIS / LDS over (y, i)
    std::vector<SfPt> yi;
    for (size_t k = 0; k < group.size(); ++k) yi.push_back(SfPt(group[k].second, (long)k));
    const std::vector<SfPt> lis = sf_lis(yi), lds = sf_lds(yi);
    const bool plus = lis.size() >= lds.size();
    const std::vector<SfPt>& track = plus ? lis : lds;
    std::vector<SfPt> g2;
    for (const SfPt& t : track) g2.push_back(group[(size_t)t.second]);
    std::set<long> xs, ysset;
    for (const SfPt& q : g2) { xs.insert(q.first); ysset.insert(q.second); }
    long score = (long)std::min(xs.size(), ysset.size());
    char gray;
    if (flanker.first == query) gray = 'S';
    else { gray = flanked ? 'F' : 'G'; score -= 1; }
    if (score < cutoff) continue;
    regions.push_back({flanker.second, other.second, g2.front().second, g2.back().second, gray, plus ? '+' : '-', score});
  }
  std::stable_sort(regions.begin(), regions.end(), [](const SfRegion& a, const SfRegion& b) { return a.score > b.score; });
}

std::string sf_note(const std::string& filename) {          // op.basename(filename).split(".")[0]
  size_t sl = filename.rfind('/');
  std::string b = sl == std::string::npos ? filename : filename.substr(sl + 1);
  return b.substr(0, b.find('.'));
}

void sf_batch(const Bed& qbed, const Bed& sbed, std::vector<SfPt> all, long window_opt, double cutoff_opt, const std::string& qnote,
              const std::string& snote, FILE* fw) {
  const long cutoff = (long)(cutoff_opt * (double)window_opt);   // int(opts.cutoff * opts.window)
  const long window = window_opt / 2;
  std::sort(all.begin(), all.end());
  std::vector<SfRegion> regions;
  size_t a = 0;
  const size_t G = qbed.rows.size();
  while (a < G) {                                           // groupby(qsimplebed, seqid): consecutive equal seqids
    size_t b = a;
    while (b < G && qbed.rows[b].seqid == qbed.rows[a].seqid) ++b;
    const long first = (long)a, last = (long)b - 1;
    for (long r = first; r <= last; ++r) {
      const long rmin = std::max(r - window, first), rmax = std::min(r + window + 1, last);
      const auto lo = std::lower_bound(all.begin(), all.end(), SfPt(rmin, 0));
      const auto hi = std::lower_bound(all.begin(), all.end(), SfPt(rmax, 0));
      std::vector<SfPt> data;
      if (lo < hi) data.assign(lo, hi);
      sf_regions(r, sbed, data, window, cutoff, regions);
      for (const SfRegion& g : regions) {
        const BedLine& L = sbed.rows[(size_t)g.left]; const BedLine& R = sbed.rows[(size_t)g.right];
        const BedLine& A = sbed.rows[(size_t)g.syntelog];
        const long left_dist = A.seqid == L.seqid ? std::labs(A.start - L.start) : 0;
        const long right_dist = A.seqid == R.seqid ? std::labs(A.start - R.start) : 0;
        const long flank = (std::max(left_dist, right_dist) / 10000 + 1) * 10000;
        fprintf(fw, "%s\t%s\t%c\t%ld\t%ld\t%c\t%s\t%s\n", qbed.rows[(size_t)r].accn.c_str(), A.accn.c_str(), g.gray, g.score, flank,
                g.orient, qnote.c_str(), snote.c_str());
      }
    }
    a = b;
  }
}
}  // namespace

int orc_synfind(const char* blast_file, const char* qbed_path, const char* sbed_path, const char* out_path, int strip_names,
                int window, double cutoff) {
  Beds beds;
  if (!load_beds(qbed_path, sbed_path, beds)) return -1;
  std::vector<Hit> hits;
  if (!read_blast_ordered(blast_file, beds, strip_names != 0, hits, nullptr)) return -1;
  std::vector<SfPt> all;
  for (const Hit& h : hits) all.push_back(SfPt(h.qi, h.si));
  FILE* fw = fopen(out_path, "w");
  if (!fw) return -1;
  const std::string qnote = sf_note(beds.q.filename), snote = sf_note(beds.s.filename);
  sf_batch(beds.q, beds.s, all, window, cutoff, qnote, snote, fw);
  if (!beds.is_self && all.empty()) { fclose(fw); g_err = "not enough values to unpack (expected 2, got 0)"; return -5; }   // transposed([])
  if (!beds.is_self) {                                      // the mirror pass: transposed data, BEDs and notes swapped
    std::vector<SfPt> tr;
    for (const SfPt& p : all) tr.push_back(SfPt(p.second, p.first));
    sf_batch(beds.s, beds.q, tr, window, cutoff, snote, qnote, fw);
  }
  fclose(fw);
  return 0;
}

}  // extern "C"

"""`jcvi.formats.blast filter` (src/jcvi/formats/blast.py:620-709).

CPU: the oracle's restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden_filter.py).  GPU: the drop-in (`jcvi_b200.formats.blast.filter`) against the
same vectors and against the oracle on a larger synthetic table."""
import os

import numpy as np
import pytest

from conftest import read


def _argv_to_kwargs(argv, d):
    kw = dict(score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, inverse=False, ids=None)
    for a in argv:
        k, _, v = a.lstrip("-").partition("=")
        if k in ("noself", "inverse"):
            kw[k] = True
        elif k == "ids":
            kw[k] = os.path.join(d, v)
        elif k in ("score", "hitlen"):
            kw[k] = int(v)
        else:
            kw[k] = float(v)
    return kw


def test_oracle_filter_matches_reference_golden(golden):
    import oracle
    d, cases = golden("blast_filter")
    assert len(cases) >= 24
    for c in cases:
        out = os.path.join(d, c["out"] + ".orc")
        oracle.blast_filter(os.path.join(d, c["input"]), out, **_argv_to_kwargs(c["argv"], d))
        assert read(out) == read(os.path.join(d, c["out"])), c


def test_filtered_blastfile_name():
    """blast.py:604-617; tests/formats/test_blast.py of the reference pins the same strings."""
    from jcvi_b200.formats.blast import filtered_blastfile_name as f
    assert f("a.last", 95.0, 100) == "a.last.P95L100"
    assert f("a.last", 98.5, 0, True) == "a.last.P98_5L0.inverse"
    assert f("a.last", 0.0, -10) == "a.last.P0L-10"


@pytest.mark.gpu
def test_gpu_filter_matches_reference_golden(golden):
    from jcvi_b200.formats import blast as gblast
    d, cases = golden("blast_filter")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for c in cases:
            got = gblast.filter([c["input"], "-o", c["outfile"] + ".gpu"] + c["argv"])
            assert os.path.basename(got) == c["out"].replace(c["outfile"], c["outfile"] + ".gpu"), c
            assert read(got) == read(c["out"]), c
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_gpu_filter_matches_oracle_on_a_larger_table(tmp_path):
    import oracle
    from jcvi_b200 import synth
    from jcvi_b200.formats import blast as gblast
    d = str(tmp_path)
    _, _, text = synth.make_pair(35, 6000, 6000, 500000, Kq=5, Ks=5, self_mode=True)
    last = os.path.join(d, "A.A.last")
    open(last, "wb").write(bytes(text))
    names = sorted({ln.split(b"\t")[0] for ln in bytes(text).split(b"\n")[3:200000:7] if ln})
    ids = os.path.join(d, "keep.ids")
    open(ids, "wb").write(b"\n".join(names) + b"\n")
    for tag, argv in (("orth", ["--pctid=98", "--inverse", "--noself"]),
                      ("cut", ["--pctid=50", "--hitlen=200", "--score=300", "--evalue=1e-40"]),
                      ("ids", ["--pctid=30", "--hitlen=0", "--ids=" + ids, "--noself"])):
        out = last + "." + tag + ".orc"
        oracle.blast_filter(last, out, **_argv_to_kwargs(argv, d))
        got = gblast.filter([last, "-o", last + "." + tag] + argv)
        assert read(got) == read(out), tag
        assert 1000 < len(read(got)) < len(bytes(text))


def _rows(text):
    """Row spans as Python's universal-newlines text mode yields them ('\\n', '\\r\\n', lone '\\r')."""
    import re
    spans, p = [], 0
    for m in re.finditer(rb"\r\n|\n|\r", text):
        spans.append((p, m.start()))
        p = m.end()
    if p < len(text):
        spans.append((p, len(text)))
    return spans


def test_kernel_row_decision_on_the_host_equals_glibc(golden):
    """The kernel's per-row code (jx_fast5.cuh filter_parse / filter_remove, host build, both the fast and the
    generic route) against the reference's arithmetic (glibc sscanf + C compares) on every row of the golden
    inputs, for every golden option set (without --ids: that is the name table's business)."""
    import ctypes
    import subprocess
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "jcvi_b200", "csrc")
    so = os.path.join(root, "jcvi_b200", "libjx_hostprobe.so")
    srcs = [os.path.join(csrc, f) for f in ("jx_hostprobe.cpp", "jx_parse.cuh", "jx_fast5.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", srcs[0], "-o", so])
    L = ctypes.CDLL(so)
    d, cases = golden("blast_filter")
    n_fast = n_rows = 0
    for c in cases:
        kw = _argv_to_kwargs(c["argv"], d)
        text = read(os.path.join(d, c["input"]))
        spans = [(s, e) for s, e in _rows(text) if text[s:s + 1] != b"#"]
        rs = np.array([s for s, _ in spans], dtype=np.int32)
        re_ = np.array([e for _, e in spans], dtype=np.int32)
        out = np.zeros(3 * len(spans), dtype=np.int8)
        route = np.zeros(len(spans), dtype=np.int8)
        L.probe_filter_rows(text, len(text), rs.ctypes.data_as(ctypes.c_void_p), re_.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_long(len(spans)), ctypes.c_double(kw["score"]), ctypes.c_double(kw["pctid"]),
                            ctypes.c_longlong(kw["hitlen"]), ctypes.c_double(kw["evalue"]), int(kw["noself"]),
                            int(kw["inverse"]), out.ctypes.data_as(ctypes.c_void_p), route.ctypes.data_as(ctypes.c_void_p))
        out = out.reshape(-1, 3)
        bad = np.nonzero((out[:, 0] != out[:, 2]) | (out[:, 1] != out[:, 2]))[0]
        assert len(bad) == 0, (c["input"], c["argv"], [(text[spans[i][0]:spans[i][1]], out[i].tolist(), int(route[i])) for i in bad[:5]])
        n_fast += int(route.sum()); n_rows += len(spans)
        if not kw["ids"]:                  # and the kept rows are the reference's output
            kept = b"".join(text[s:e].rstrip() + b"\n" for (s, e), o in zip(spans, out[:, 0]) if o == 1)
            assert kept == read(os.path.join(d, c["out"])), c
    assert n_fast > 0.8 * n_rows           # the fast route is the one being exercised


def test_exact_strtod_greater_than_cutoff_on_the_host():
    """jx_parse.cuh dec_gt_double (host build): `strtod(text) > c` for decimals at, next to and half-way between
    the doubles around c -- the decision `c.evalue > evalue` (blast.py:690) needs.  Python's float() is the
    correctly rounded conversion glibc's strtod is.  Refusals (-1) are allowed only beyond 19 significant digits."""
    import ctypes
    import math
    import random
    import struct
    from decimal import Decimal, getcontext
    from fractions import Fraction
    getcontext().prec = 1200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = ctypes.CDLL(os.path.join(root, "jcvi_b200", "libjx_hostprobe.so"))
    L.probe_gt_double.argtypes = [ctypes.c_char_p, ctypes.c_double]
    rng = random.Random(11)
    fixed = [0.0, 0.01, 1e-30, 1e-300, 5e-324, 1.7976931348623157e308, 1e-10, 1.0, 10.0, 1e22, 1e23, 2.0 ** 53,
             2.2250738585072014e-308]

    def cutoff():
        k = rng.random()
        if k < 0.15:
            return rng.choice(fixed)
        bits = rng.getrandbits(63)
        if (bits >> 52) & 0x7FF == 0x7FF:
            bits &= ~(1 << 62)
        if k < 0.6:
            bits = (bits & ((1 << 52) - 1)) | ((1023 + rng.randint(-120, 120)) << 52)
        return struct.unpack("<d", struct.pack("<Q", bits))[0]

    def text(fr, digits, bump):
        d = Decimal(fr.numerator) / Decimal(fr.denominator)
        if d == 0:
            return "0"
        _, dig, ex = d.as_tuple()
        keep = dig[:digits]
        m = max(int("".join(map(str, keep))) + bump, 0)
        e = ex + len(dig) - len(keep)
        return "%de%d" % (m, e) if rng.random() < 0.5 else "%s.%se%d" % (str(m)[0], str(m)[1:] or "0", e + len(str(m)) - 1)

    n = 0
    for _ in range(400):
        c = cutoff() * (-1 if rng.random() < 0.2 else 1)
        ac = abs(c)
        up = math.nextafter(ac, math.inf)
        dn = math.nextafter(ac, -math.inf) if ac > 0 else 0.0
        points = [Fraction(ac), Fraction(dn), (Fraction(ac) + Fraction(dn)) / 2]
        points += [Fraction(up), (Fraction(ac) + Fraction(up)) / 2] if up != math.inf else [Fraction(ac) + Fraction(2) ** 970]
        cands = [repr(ac), repr(dn), "1e-400", "2.4703282292062327e-324", "2.4703282292062328e-324", "1e400", "0e10"]
        for fr in points:
            for digits in (rng.randint(1, 16), 17, 18, 19, rng.randint(20, 40)):
                cands += [text(fr, digits, b) for b in (-1, 0, 1)]
        for s in cands:
            for t in (s, "-" + s):
                got = L.probe_gt_double(t.encode(), c)
                digs = t.lstrip("-").split("e")[0].replace(".", "").lstrip("0")
                if got == -1:
                    assert len(digs) > 19, (t, c)
                else:
                    assert got == (1 if float(t) > c else 0), (t, c, got)
                n += 1
    assert n > 50000


@pytest.mark.gpu
def test_gpu_filter_edge_inputs(tmp_path):
    """Empty and comment-only tables, a table without a final newline, rows longer than the tokenizer's window,
    rows cut by tile boundaries, an --ids file whose names are absent, a 128-byte name (refused loudly like the
    rest of the path: the reference overflows its char[128])."""
    import oracle
    from jcvi_b200.engine import get_engine
    from jcvi_b200.formats import blast as gblast
    eng = get_engine()
    d = str(tmp_path)
    row = "Ag%06d.1\tAg%06d.1\t%.2f\t%d\t3\t1\t10\t%d\t20\t%d\t%s\t%.1f"

    def table(n, tail=""):
        out = []
        for i in range(n):
            out.append(row % (i % 977, (i * 7) % 977 if i % 13 else i % 977, 90 + (i % 100) / 10.0, 50 + i % 300,
                              10 + i % 300, 20 + i % 300, "%.1e" % (10.0 ** -(i % 40)), 30 + (i % 2000) / 3.0) + tail)
        return out

    cases = {
        "empty": b"",
        "comments": b"# a\n# b\n",
        "one_no_newline": (table(1)[0]).encode(),
        "blank_only": b"\n\n\n",
        "long_rows": ("\n".join(r + "\t" + "x" * (1500 + 37 * (i % 50)) for i, r in enumerate(table(400))) + "\n").encode(),
        "many": ("\n".join(table(60000)) + "\n").encode(),           # ~4 MB: hundreds of tile boundaries
        "trailing_ws": ("\n".join(table(3000, tail=" \t ")) + "\n").encode(),
    }
    for name, text in cases.items():
        p = os.path.join(d, name + ".last")
        open(p, "wb").write(text)
        for tag, kw in (("dflt", {}), ("inv", dict(pctid=95.0, inverse=True, noself=True)), ("ev", dict(pctid=0.0, hitlen=0, evalue=1e-20))):
            okw = dict(score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, inverse=False)
            okw.update(kw)
            oracle.blast_filter(p, p + ".orc", **okw)
            got, st = eng.blast_filter(np.frombuffer(text, dtype=np.uint8), **okw)
            assert got == read(p + ".orc"), (name, tag)
            assert st["kept"] == got.count(b"\n")
    # --ids whose names never occur: nothing is kept, everything with --inverse
    p = os.path.join(d, "many.last")
    ids = os.path.join(d, "none.ids")
    open(ids, "w").write("nobody\nnothing,else\n")
    out = gblast.filter([p, "-o", os.path.join(d, "x"), "--pctid=0", "--hitlen=0", "--evalue=10", "--ids=" + ids])
    assert read(out) == b""
    out = gblast.filter([p, "-o", os.path.join(d, "y"), "--pctid=0", "--hitlen=0", "--evalue=10", "--ids=" + ids, "--inverse"])
    assert read(out) == cases["many"]
    # a 128-byte name
    bad = ("A" * 128 + "\tB\t99.0\t200\t0\t0\t1\t2\t3\t4\t1e-5\t50\n").encode()
    with pytest.raises(NotImplementedError):
        eng.blast_filter(np.frombuffer(bad, dtype=np.uint8))


def test_kernel_row_decision_fuzz_against_glibc():
    """Randomly perturbed rows (number formats, signs, leading zeros, exponents, separators, missing and extra
    fields, cut-off-exact values): the kernel's row code on both routes must agree with glibc sscanf + C compares
    for several cut-off sets, or refuse (-1) only where the path documents a refusal (20+ significant digits,
    10+ digit integers, dangling exponents, inf / nan / hex)."""
    import ctypes
    import random
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    L = ctypes.CDLL(os.path.join(root, "jcvi_b200", "libjx_hostprobe.so"))
    rng = random.Random(5)

    def num(kind):
        r = rng.random()
        if kind == "int":
            v = rng.choice([0, 1, 99, 100, 101, 150, rng.randint(0, 5000), rng.randint(0, 10 ** 8)])
            s = str(v)
            if r < 0.05:
                s = "0" * rng.randint(1, 3) + s
            elif r < 0.08:
                s = "+" + s
            elif r < 0.11:
                s = "-" + s
            return s
        if kind == "pct":
            v = rng.choice([95.0, 94.99, 95.01, 60.0, 100.0, rng.uniform(0, 100)])
            return rng.choice(["%.2f", "%.1f", "%.0f", "%.3f", "%g"]) % v
        if kind == "ev":
            if r < 0.15:
                return rng.choice(["0", "0.0", "0.01", "1e-2", "1e-30", "1.0e-30", "0.010000000000000002", "9.999999999999999e-31", "1e-400", "1e5"])
            return rng.choice(["%.1e", "%.2e", "%g", "%.0e"]) % (10.0 ** rng.uniform(-60, 2))
        if kind == "score":
            v = rng.choice([400.0, 399.99, 400.01, rng.uniform(20, 3000)])
            return rng.choice(["%.1f", "%.0f", "%g", "%.2e", "%.3g"]) % v
        raise ValueError(kind)

    def row():
        q = "g%05d.%d" % (rng.randint(0, 300), rng.randint(1, 3))
        s = q if rng.random() < 0.05 else "h%05d" % rng.randint(0, 300)
        f = [q, s, num("pct")] + [num("int") for _ in range(7)] + [num("ev"), num("score")]
        r = rng.random()
        sep = "\t"
        if r < 0.04:
            f = f[:rng.randint(0, 11)]
        elif r < 0.07:
            f += ["extra", "cols"]
        elif r < 0.10:
            sep = rng.choice([" ", "  ", " \t"])
        elif r < 0.12:
            f[rng.randint(2, 11)] = rng.choice(["abc", "", "1e", ".", "-", "1.2.3", "12abc"])
        line = sep.join(f)
        if rng.random() < 0.03:
            line = " " + line
        if rng.random() < 0.03:
            line += rng.choice([" ", "\t", " \t "])
        return line

    rows = [row() for _ in range(20000)]
    text = ("\n".join(rows) + "\n").encode()
    spans = _rows(text)
    rs = np.array([s for s, _ in spans], dtype=np.int32)
    re_ = np.array([e for _, e in spans], dtype=np.int32)
    n_refused = 0
    for kw in (dict(score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, inverse=False),
               dict(score=400, pctid=60.0, hitlen=150, evalue=1e-30, noself=True, inverse=False),
               dict(score=-5, pctid=0.0, hitlen=0, evalue=0.0, noself=True, inverse=True)):
        out = np.zeros(3 * len(spans), dtype=np.int8)
        route = np.zeros(len(spans), dtype=np.int8)
        L.probe_filter_rows(text, len(text), rs.ctypes.data_as(ctypes.c_void_p), re_.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_long(len(spans)), ctypes.c_double(kw["score"]), ctypes.c_double(kw["pctid"]),
                            ctypes.c_longlong(kw["hitlen"]), ctypes.c_double(kw["evalue"]), int(kw["noself"]),
                            int(kw["inverse"]), out.ctypes.data_as(ctypes.c_void_p), route.ctypes.data_as(ctypes.c_void_p))
        out = out.reshape(-1, 3)
        for col in (0, 1):
            decided = out[:, col] >= 0
            bad = np.nonzero(decided & (out[:, col] != out[:, 2]))[0]
            assert len(bad) == 0, (kw, col, [(rows[i], out[i].tolist(), int(route[i])) for i in bad[:5]])
        refused = np.nonzero(out[:, 0] < 0)[0]
        n_refused += len(refused)
        for i in refused[:50]:                      # refusals must be explicable
            toks = rows[i].split()
            assert (any(len(t.lstrip("+-").lstrip("0")) >= 10 for t in toks[3:10])
                    or any(len(t.split("e")[0].replace(".", "").lstrip("+-0")) > 19 for t in toks[2:])
                    or any(t[-1:] in "eE" or t[-2:] in ("e+", "e-", "E+", "E-") for t in toks[2:])), rows[i]     # dangling exponent
        assert route.mean() > 0.3            # both routes get their share of the rows
    assert n_refused < 0.02 * 3 * len(rows)

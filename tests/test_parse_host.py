"""CPU checks of the tokenizer arithmetic (jcvi_b200/csrc/jx_parse.cuh, host build) against
glibc -- the very functions the reference calls through cblast.pyx (sscanf "%f"/"%lf",
sprintf "%.3g")."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "jcvi_b200", "csrc")
SO = os.path.join(ROOT, "jcvi_b200", "libjx_hostprobe.so")
libc = ctypes.CDLL("libc.so.6")
libc.strtof.restype = ctypes.c_float
libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
libc.strtod.restype = ctypes.c_double
libc.strtod.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


@pytest.fixture(scope="module")
def probe():
    srcs = [os.path.join(CSRC, f) for f in ("jx_hostprobe.cpp", "jx_parse.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               srcs[0], "-o", SO])
    L = ctypes.CDLL(SO)
    L.probe_round3g_range.restype = ctypes.c_long
    L.probe_hash.restype = ctypes.c_ulonglong
    return L


def f32bits(x):
    return np.float32(x).view(np.uint32)


def _decimals(rng, n):
    out = []
    for _ in range(n):
        kind = rng.integers(0, 8)
        if kind == 0:
            out.append("%d" % rng.integers(0, 10**6))
        elif kind == 1:
            out.append("%.1f" % rng.uniform(0, 5000))
        elif kind == 2:
            out.append("%.2f" % rng.uniform(0, 100))
        elif kind == 3:
            out.append("%.3g" % rng.uniform(0.5, 2e6))
        elif kind == 4:
            out.append("%.*e" % (int(rng.integers(0, 10)), rng.uniform(1e-3, 1e7)))
        elif kind == 5:
            # near float32 midpoints: take a float, move half an ulp, print with few digits
            f = np.float32(rng.uniform(1, 4096))
            nxt = np.nextafter(f, np.float32(np.inf))
            mid = (float(f) + float(nxt)) / 2
            out.append("%.*f" % (int(rng.integers(4, 12)), mid))
        elif kind == 6:
            out.append("%.*f" % (int(rng.integers(0, 12)), rng.uniform(0, 3000)))
        else:
            out.append("%se%+d" % (rng.integers(1, 99999), rng.integers(-6, 6)))
    return out


def test_float32_parse_matches_strtof(probe):
    rng = np.random.default_rng(1)
    n_ok = n_hard = 0
    for s in _decimals(rng, 80000) + ["0", "0.0", "-0", "54.0", "1.06e+03", "1e+05", "100", "+7.5", ".5", "5.",
                                      "16777217", "33554433", "1e38", "3.5e38", "1234.5678", "0.1", "99999",
                                      "54.000000000000000"]:
        out = ctypes.c_float()
        used = ctypes.c_int()
        rc = probe.probe_f32(s.encode(), ctypes.byref(out), ctypes.byref(used))
        assert rc in (0, 1), s
        if rc == 0:
            n_hard += 1
            continue
        n_ok += 1
        assert used.value == len(s), s
        want = libc.strtof(s.encode(), None)
        assert f32bits(out.value) == f32bits(want), (s, out.value, want)
    assert n_hard < 0.01 * n_ok      # the exact routes cover (nearly) everything that occurs


def test_float_prefix_and_failures(probe):
    out = ctypes.c_float(); used = ctypes.c_int()
    assert probe.probe_f32(b"54.0abc", ctypes.byref(out), ctypes.byref(used)) == 1 and used.value == 4
    assert probe.probe_f32(b"abc", ctypes.byref(out), ctypes.byref(used)) == -1
    assert probe.probe_f32(b"-", ctypes.byref(out), ctypes.byref(used)) == -1
    assert probe.probe_f32(b".", ctypes.byref(out), ctypes.byref(used)) == -1
    assert probe.probe_f32(b"none", ctypes.byref(out), ctypes.byref(used)) == -1
    for bad in (b"nan", b"inf", b"-Infinity", b"0x1p3", b"1e", b"1e+"):
        assert probe.probe_f32(bad, ctypes.byref(out), ctypes.byref(used)) == 0, bad


def test_evalue_threshold_matches_strtod(probe):
    rng = np.random.default_rng(2)
    cases = ["1e-10", "1.0e-10", "9.9e-11", "9.99999e-11", "1.00001e-10", "0", "0.0", "1e-400", "1e-50", "2.5e-180",
             "1e400", "5", "0.001", "1e-9", "1e-11", "100e-12", "10e-11", "0.0000000001", "0.00000000009",
             "1234567890123456789012e-32", "1000000000000e-22", "999999999999e-22"]
    for _ in range(40000):
        cases.append("%.*e" % (int(rng.integers(0, 6)), 10.0 ** rng.uniform(-200, 3)))
    for _ in range(20000):
        cases.append("%.*e" % (int(rng.integers(0, 8)), 1e-10 * rng.uniform(0.9, 1.1)))
    hard = 0
    for s in cases:
        rc = probe.probe_lt1e10(s.encode())
        if rc == -1:
            hard += 1
            continue
        want = 1 if libc.strtod(s.encode(), None) < 1e-10 else 0
        assert rc == want, s
    assert hard == 0


def test_round3g_matches_printf_strtof(probe):
    bad = ctypes.c_uint(); unsup = ctypes.c_long()
    # every 97th float32 between 1e-3 and 1e12, plus dense sweeps where scores live
    lo, hi = int(f32bits(1e-3)), int(f32bits(1e12))
    assert probe.probe_round3g_range(lo, hi, 97, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    assert unsup.value == 0
    lo, hi = int(f32bits(30.0)), int(f32bits(3000.0))
    assert probe.probe_round3g_range(lo, hi, 1, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    assert unsup.value == 0


def test_strip_isoform_matches_oracle(probe):
    import oracle
    for s in ["Ag000123.1", "Ag000123.12", "evm.model.x.5", "a.b.c|x.1", "abc", "abc.", ".1", "a|b", "a.1|b.2",
              "ev.1", "e.1", "x..2", "AT1G01010.1", "gene|", "|x"]:
        b = probe.probe_strip(s.encode(), len(s))
        assert s[:b] == oracle.gene_name(s), s


def test_tokenize_line_matches_sscanf(probe):
    rng = np.random.default_rng(3)
    base = "Ag000791.3\tBg002441.1\t75.72\t1173\t284\t18\t1402\t2575\t1556\t2729\t6.5e-21\t161.5"
    lines = [base, base.replace("\t", " "), "  " + base, base + "\textra\tcols", "\t".join(base.split("\t")[:8]),
             "a b", "a", "", "   ", base.replace("1173", "x1173"), base + "abc", base.replace("6.5e-21", "1e-10"),
             base.replace("161.5", "1.06e+03"), base.replace("75.72", "abc"), base.replace("\t284\t", "\t-284\t"),
             base.replace("\t18\t", "\t+18\t"), base.replace("\t18\t", "\t1.8\t"), "a\tb\t1\t2"]
    for ln in lines:
        spans = (ctypes.c_int * 4)(); score = ctypes.c_float(); nconv = ctypes.c_int(); eflag = ctypes.c_int()
        fl = probe.probe_line(ln.encode(), len(ln), spans, ctypes.byref(score), ctypes.byref(nconv), ctypes.byref(eflag))
        q = ctypes.create_string_buffer(256); s = ctypes.create_string_buffer(256)
        gs = ctypes.c_float(); ge = ctypes.c_double(); ints = (ctypes.c_int * 7)()
        n = probe.probe_line_glibc((ln + "\n").encode(), q, s, ctypes.byref(gs), ctypes.byref(ge), ints)
        n = max(n, 0)
        assert fl == 0, ln
        assert nconv.value == n, (ln, nconv.value, n)
        assert ln[spans[0]:spans[1]] == q.value.decode()
        assert ln[spans[2]:spans[3]] == s.value.decode()
        assert f32bits(score.value) == f32bits(gs.value), ln
        if n >= 11:
            assert eflag.value == (1 if ge.value < 1e-10 else 0)
        else:
            assert eflag.value == 1      # evalue stays 0.0 in the Cython object


def test_hash_is_stable(probe):
    # host table build and device probe share this function; pin a few values against drift
    h1 = probe.probe_hash(b"Ag000123", 8)
    h2 = probe.probe_hash(b"Ag000124", 8)
    assert h1 != h2 and probe.probe_hash(b"Ag000123x", 8) == h1


@pytest.mark.parametrize("version", [2, 3, 4, 5])
def test_fast_line_agrees_with_generic_scanner(probe, version):
    """The SWAR fast path (jx::fast_line) must either refuse a line or produce exactly what the
    generic sscanf-faithful scanner produces (spans, score bits, evalue flag, stripped name,
    hash of the stripped name)."""
    rng = np.random.default_rng(4)
    base = "Ag000791.3\tBg002441.1\t75.72\t1173\t284\t18\t1402\t2575\t1556\t2729\t6.5e-21\t161.5"
    lines = [base]
    names = ["Ag1.1", "evm.model.000021F_np12.565", "mrna.AA565170.t1", "AT1G01010.1|x", "a|b.1", "g.12", "x", "ev.1",
             "Os09g11510", "a.b.c", "averyveryveryverylongnamethatkeepsgoing_0123456789.2", "n..1", "p.1.", ".1",
             "evx|y.1", "ab.1|cd", "abcd|efgh.2", "abc.1|", "|abc"]
    nums = ["161.5", "54.0", "1386", "1.06e+03", "1e+05", "0", "99999", "3.5", ".5", "5.", "12345678.125", "1e-3"]
    evs = ["6.5e-21", "0.0", "0", "1e-10", "9.9e-11", "2.5e-180", "0.001", "3", "1e-400", "1.0e-10", "5e-11", "10e-11",
           "100e-12", "0.5e-10", "12.5e-12", "9.9999999e-11", "99999999e-18", "1E-11", "1e+2", "1e-09", "0.00000000001",
           "0.31", "007e-12", "1.e-11", "5e-010"]
    pcts = ["75.72", "100", "100.00", "99.", ".5", "7"]
    for _ in range(4000):
        f = [names[rng.integers(len(names))], names[rng.integers(len(names))], pcts[rng.integers(len(pcts))]]
        f += [str(rng.integers(0, 10 ** int(rng.integers(1, 7)))) for _ in range(7)]
        f += [evs[rng.integers(len(evs))], nums[rng.integers(len(nums))]]
        ln = "\t".join(f)
        r = rng.integers(0, 28)             # half of the lines stay regular
        if r == 0: ln = ln.replace("\t", " ", 1)
        elif r == 1: ln = ln + "\textra"
        elif r == 2: ln = "\t".join(f[:9])
        elif r == 3: ln = ln.replace("\t", "\t\t", 1)
        elif r == 4: f[5] = "-" + f[5]; ln = "\t".join(f)
        elif r == 5: f[4] = "1.5"; ln = "\t".join(f)
        elif r == 6: ln = ln + "abc"
        elif r == 7: f[0] = "a b"; ln = "\t".join(f)
        elif r == 8: f[2] = "1.2.3"; ln = "\t".join(f)
        elif r == 9: f[10] = "1e"; ln = "\t".join(f)
        elif r == 10: f[11] = "nan"; ln = "\t".join(f)
        elif r == 11: ln = " " + ln
        elif r == 12: ln = ln + "\r"
        lines.append(ln)
    nfast = 0
    for ln in lines:
        for pad in (0, 1, 2, 3):          # every alignment of the line start
            raw = (b"#" * pad) + ln.encode()
            buf = raw + b"\n" + b"\0" * 24
            s, n = pad, len(raw)
            spans = (ctypes.c_int * 4)(); score = ctypes.c_float(); eflag = ctypes.c_int()
            ends = (ctypes.c_int * 2)(); hs = (ctypes.c_ulonglong * 2)()
            if version == 2:
                ok = probe.probe_fast_line(buf, s, n, spans, ctypes.byref(score), ctypes.byref(eflag), ends, hs)
            elif version == 3:
                ok = probe.probe_fast_line3(buf, s, len(buf), spans, ctypes.byref(score), ctypes.byref(eflag), ends, hs)
            elif version == 4:
                ok = probe.probe_fast_line4(buf, s, n, spans, ctypes.byref(score), ctypes.byref(eflag), ends, hs)
            else:
                ok = probe.probe_fast_line5(buf, len(buf), s, n, spans, ctypes.byref(score), ctypes.byref(eflag), ends, hs)
            gsp = (ctypes.c_int * 4)(); gsc = ctypes.c_float(); gnc = ctypes.c_int(); gef = ctypes.c_int()
            fl = probe.probe_line(ln.encode(), len(ln), gsp, ctypes.byref(gsc), ctypes.byref(gnc), ctypes.byref(gef))
            if not ok:
                continue
            nfast += 1
            assert fl == 0 and gnc.value == 12, ln
            assert [x - pad for x in spans] == list(gsp), ln
            assert f32bits(score.value) == f32bits(gsc.value), ln
            assert eflag.value == gef.value, ln
            for k in (0, 1):
                name = ln[gsp[2 * k]:gsp[2 * k + 1]]
                b = probe.probe_strip(name.encode(), len(name))
                assert ends[k] - pad - gsp[2 * k] == b, (ln, k)
                assert hs[k] == probe.probe_hash(name[:b].encode(), b), (ln, k)
    assert nfast > 0.2 * 4 * len(lines)


def test_byte_class_masks_match_definition(probe):
    """Phase A of the v5 tokenizer: W = byte <= 0x20, N = neither digit nor tab, '|' detector."""
    rng = np.random.default_rng(11)
    blocks = [bytes(range(k, k + 32)) for k in range(0, 256, 32)]
    blocks += [bytes(rng.integers(0, 256, 32, dtype=np.uint8)) for _ in range(3000)]
    blocks += [bytes(rng.choice(np.frombuffer(b"0123456789\t\n\r .|e-+Az", dtype=np.uint8), 32)) for _ in range(3000)]
    for b in blocks:
        wm, nm, hi = ctypes.c_uint(), ctypes.c_uint(), ctypes.c_uint()
        probe.probe_classify32(b, ctypes.byref(wm), ctypes.byref(nm), ctypes.byref(hi))
        want_w = sum(1 << i for i, c in enumerate(b) if c <= 0x20)
        want_n = sum(1 << i for i, c in enumerate(b) if not (0x30 <= c <= 0x39 or c == 9))
        assert wm.value == want_w and nm.value == want_n, b
        if any(c == 0x7C for c in b):
            assert hi.value != 0, b
        if all((c & 0x7F) < 0x7C for c in b):
            assert hi.value == 0, b


def test_fast_line5_alignments_and_formats(probe):
    """v5 reads its masks at a bit offset (line start mod 32) and converts numbers from words: sweep
    every alignment, name lengths up to the 128-byte line limit and the number formats the aligners
    and jcvi's own writers produce (%d, %.1f, %.2f, %.2g, %.3g, %g, %e)."""
    rng = np.random.default_rng(2025)
    alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_.-|"

    def name():
        n = int(rng.integers(1, 36))
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet) - (0 if rng.random() < 0.05 else 1), n))

    def num(kind):
        x = float(10 ** rng.uniform(-3, 6)) if kind == "score" else float(10 ** rng.uniform(-200, 2))
        if rng.random() < 0.1:
            x = 0.0
        fmt = ["%d", "%.1f", "%.2f", "%.2g", "%.3g", "%g", "%e", "%.3e", "%.0f"][int(rng.integers(0, 9))]
        return (fmt % (int(x) if fmt == "%d" else x)).replace("e", "E" if rng.random() < 0.1 else "e")

    naccept = ntotal = 0
    for it in range(6000):
        f = [name(), name(), "%.2f" % rng.uniform(0, 100) if rng.random() < 0.8 else str(int(rng.integers(0, 101)))]
        f += [str(int(rng.integers(0, 10 ** int(rng.integers(1, 7))))) for _ in range(7)]
        f += [num("evalue"), num("score")]
        ln = "\t".join(f)
        pad = it % 41
        raw = (b"#" * pad) + ln.encode()
        buf = raw + (b"\n" if it % 3 else b"\r\n") + b"\0" * 24
        s, n = pad, len(raw)
        spans = (ctypes.c_int * 4)(); score = ctypes.c_float(); eflag = ctypes.c_int()
        ends = (ctypes.c_int * 2)(); hs = (ctypes.c_ulonglong * 2)()
        ok = probe.probe_fast_line5(buf, len(buf), s, n, spans, ctypes.byref(score), ctypes.byref(eflag), ends, hs)
        gsp = (ctypes.c_int * 4)(); gsc = ctypes.c_float(); gnc = ctypes.c_int(); gef = ctypes.c_int()
        fl = probe.probe_line(ln.encode(), len(ln), gsp, ctypes.byref(gsc), ctypes.byref(gnc), ctypes.byref(gef))
        ntotal += 1
        if not ok:
            continue
        naccept += 1
        assert fl == 0 and gnc.value == 12, ln
        assert [x - pad for x in spans] == list(gsp), ln
        assert f32bits(score.value) == f32bits(gsc.value), ln
        assert eflag.value == gef.value, ln
        for k in (0, 1):
            nm = ln[gsp[2 * k]:gsp[2 * k + 1]]
            b = probe.probe_strip(nm.encode(), len(nm))
            assert ends[k] - pad - gsp[2 * k] == b, (ln, k)
            assert hs[k] == probe.probe_hash(nm[:b].encode(), b), (ln, k)
    assert naccept > 0.6 * ntotal, (naccept, ntotal)


def test_slot_hash_word_route_equals_byte_route(probe):
    """Host table build hashes names byte-wise (slot_hash_bytes); the kernel hashes six zero-padded
    words taken from aligned text words at any alignment (name_words6 + slot_hash6).  They must agree,
    and the words must be the name's bytes followed by zeros."""
    rng = np.random.default_rng(31)
    alphabet = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_.-|", dtype=np.uint8)
    for it in range(4000):
        L = int(rng.integers(1, 24))
        pad = int(rng.integers(0, 9))
        name = bytes(rng.choice(alphabet, L))
        buf = b"\t" * pad + name + b"\t" + bytes(rng.choice(alphabet, 40))      # junk after the name must not leak in
        words = (ctypes.c_uint * 6)()
        h6 = probe.probe_slot_hash6(buf, pad, L, words)
        off = np.array([0, L], dtype=np.uint32)
        out = np.zeros(1, dtype=np.uint32)
        probe.probe_slot_hash_many(name, off.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(1), out.ctypes.data_as(ctypes.c_void_p))
        assert (h6 & 0xFFFFFFFF) == int(out[0]), (name, pad)
        assert np.array(list(words), dtype="<u4").tobytes() == name + b"\0" * (24 - L), (name, pad)


def test_fast_line6_agrees_with_generic_scanner(probe):
    """v6 (jx_fast6.cuh, the kernel's first-level route): whatever it ACCEPTS must equal the sscanf-faithful generic
    scanner field for field -- every alignment, short and long names, the number formats the aligners write, and
    near-regular lines with one defect (those must be refused or still parse identically)."""
    rng = np.random.default_rng(606)
    alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789_.-"

    def name():
        n = int(rng.integers(1, 17)) if rng.random() < 0.9 else int(rng.integers(17, 40))
        return "".join(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n))

    def evalue():
        r = rng.random()
        if r < 0.5:
            x = float(10 ** rng.uniform(-300, 3))
            return ["%.1e", "%.2e", "%.0e", "%.3g", "%g", "%.2g"][int(rng.integers(0, 6))] % x
        if r < 0.6:
            return ["0.0", "0", "0.000", "1e-10", "9.99e-11", "1.0e-10", "1e-11", "0.5e-9", "0.0e-12", "1e-9", "1e-100",
                    "2e-05", "1E-20", "1.5e+02", "0.00000000001", "0.0000000001", "12.5e-12", "9e-011", "1e-0100"][int(rng.integers(0, 19))]
        return ["%.3f", "%.1f", "%d", "%.5f", "%.9f"][int(rng.integers(0, 5))] % (
            rng.uniform(0, 10) if rng.random() < 0.7 else 0)

    def score():
        r = rng.random()
        if r < 0.6:
            return ["%.1f", "%.0f", "%.2f", "%.3f", "%.4f"][int(rng.integers(0, 5))] % float(10 ** rng.uniform(0, 5))
        if r < 0.8:
            return str(int(10 ** rng.uniform(0, 8)))
        return ["16777215", "16777216", "16777217", "1677721.5", "999.9999", "0.5", "0", "100.", "1e3", "3.2e+02", ".5", "1.23456",
                "99999999", "8388608.5", "00012.5"][int(rng.integers(0, 15))]

    naccept = ntotal = 0
    for it in range(12000):
        f = [name(), name(), "%.2f" % rng.uniform(0, 100) if rng.random() < 0.8 else str(int(rng.integers(0, 101)))]
        f += [str(int(rng.integers(0, 10 ** int(rng.integers(1, 7))))) for _ in range(7)]
        f += [evalue(), score()]
        if it % 7 == 0:                                  # one defect
            k = int(rng.integers(0, 8))
            j = int(rng.integers(0, 12))
            if k == 0: f[j] = ""
            elif k == 1: f[j] = f[j] + " "
            elif k == 2: f[j] = f[j][:1] + "x" + f[j][1:]
            elif k == 3: f[j] = "-" + f[j]
            elif k == 4: f = f[:j] + f[j + 1:]
            elif k == 5: f[j] = f[j] + "\t"
            elif k == 6: f[2] = f[2].replace(".", "..") if "." in f[2] else "."
            else: f[j] = f[j].replace("e", "e+") if "e" in f[j] else f[j] + ".7."
        ln = "\t".join(f)
        pad = it % 41
        raw = (b"#" * pad) + ln.encode()
        buf = raw + b"\n" + b"Zq9\tnext\t1.0\t" * 3
        s, n = pad, len(raw)
        spans = (ctypes.c_int * 4)(); sc = ctypes.c_float(); eflag = ctypes.c_int()
        ends = (ctypes.c_int * 2)(); hs = (ctypes.c_uint * 2)()
        ok = probe.probe_fast_line6(buf, len(buf), s, n, spans, ctypes.byref(sc), ctypes.byref(eflag), ends, hs)
        ntotal += 1
        if not ok:
            continue
        naccept += 1
        gsp = (ctypes.c_int * 4)(); gsc = ctypes.c_float(); gnc = ctypes.c_int(); gef = ctypes.c_int()
        fl = probe.probe_line(ln.encode(), len(ln), gsp, ctypes.byref(gsc), ctypes.byref(gnc), ctypes.byref(gef))
        assert fl == 0 and gnc.value == 12, ln
        assert [x - pad for x in spans] == list(gsp), ln
        assert f32bits(sc.value) == f32bits(gsc.value), ln
        assert eflag.value == gef.value, ln
        # and against glibc itself
        q = ctypes.create_string_buffer(256); sb = ctypes.create_string_buffer(256)
        gs = ctypes.c_float(); ge = ctypes.c_double(); ints = (ctypes.c_int * 7)()
        assert probe.probe_line_glibc(ln.encode(), q, sb, ctypes.byref(gs), ctypes.byref(ge), ints) == 12, ln
        assert f32bits(gs.value) == f32bits(sc.value) and (1 if ge.value < 1e-10 else 0) == eflag.value, ln
        for k in (0, 1):
            nm = ln[gsp[2 * k]:gsp[2 * k + 1]]
            if "|" in nm:
                continue
            b = probe.probe_strip(nm.encode(), len(nm))
            assert ends[k] - pad - gsp[2 * k] == b, (ln, k)
            off = np.array([0, b], dtype=np.uint32); out = np.zeros(1, dtype=np.uint32)
            probe.probe_slot_hash_many(nm[:b].encode() or b"\0", off.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(1),
                                       out.ctypes.data_as(ctypes.c_void_p))
            assert hs[k] == int(out[0]), (ln, k)
    assert naccept > 0.35 * ntotal, (naccept, ntotal)
    # the shapes the aligners write must ALL take this route (LAST / BLAST+ tab12: %.2f pctid, %.1e / %.0e / 0.0 evalues,
    # integer or one-decimal scores)
    for it in range(3000):
        x = float(10 ** rng.uniform(-250, -1))
        ev = ["%.1e" % x, "%.0e" % x, "0.0", "%.2e" % x][it % 4]
        scv = float(10 ** rng.uniform(1, 4.5))
        f = ["Ag%06d.%d" % (rng.integers(0, 10 ** 6), rng.integers(1, 4)), "Bg%06d" % rng.integers(0, 10 ** 6), "%.2f" % rng.uniform(20, 100)]
        f += [str(int(rng.integers(1, 5000))) for _ in range(7)] + [ev, ["%.1f" % scv, "%d" % scv][it % 2]]
        ln = "\t".join(f)
        pad = it % 37
        raw = (b"x" * pad) + ln.encode()
        buf = raw + b"\n" + b"Zq9\tnext\t1.0\t" * 3
        spans = (ctypes.c_int * 4)(); sc = ctypes.c_float(); eflag = ctypes.c_int()
        ends = (ctypes.c_int * 2)(); hs = (ctypes.c_uint * 2)()
        assert probe.probe_fast_line6(buf, len(buf), pad, len(raw), spans, ctypes.byref(sc), ctypes.byref(eflag), ends, hs) == 1, ln
        assert f32bits(sc.value) == f32bits(np.float32(f[11])), ln
        assert eflag.value == (1 if float(ev) < 1e-10 else 0), ln


def test_name_table_build_is_perfect_and_compact(probe):
    """jx_bed_load's host build (jx_nametable.h) with the multiply-add name hash: every name is found by the lookup the
    kernels do, the table has the minimum power-of-two size for load <= 0.5, equal 32-bit hashes stay at the birthday rate."""
    import math
    rng = np.random.default_rng(5)
    sets = [[b"Ag%06d" % i for i in range(40000)], [b"AT%dG%05d" % (c, i * 10) for c in range(1, 6) for i in range(7000)],
            [b"evm.model.scaffold_%d.%d" % (c, i) for c in range(1, 200) for i in range(150)],
            list({bytes(rng.integers(97, 123, size=rng.integers(1, 40)).astype(np.uint8)) for _ in range(60000)})]
    probe.probe_build_table.restype = ctypes.c_long
    for names in sets:
        blob = b"".join(names)
        off = np.zeros(len(names) + 1, dtype=np.uint32)
        np.cumsum([len(x) for x in names], out=off[1:])
        out = (ctypes.c_long * 4)()
        missing = probe.probe_build_table(blob, off.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(names)), out)
        assert missing == 0
        assert out[0] == math.ceil(math.log2(2 * len(names) + 2)), list(out)
        assert out[2] <= 3 + len(names) ** 2 // 2 ** 31, list(out)

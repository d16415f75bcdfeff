"""CPU checks of the exact "%.2f" / "%.2g" / "%.3g" re-implementation (jcvi_b200/csrc/
jx_format.cuh, host build) against glibc snprintf -- the writer of `.filtered` lines
(cblast.pyx:17,144-156)."""
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
libc.strtod.restype = ctypes.c_double
libc.strtod.argtypes = [ctypes.c_char_p, ctypes.c_void_p]


@pytest.fixture(scope="module")
def probe():
    srcs = [os.path.join(CSRC, f) for f in ("jx_hostprobe.cpp", "jx_parse.cuh", "jx_format.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                               srcs[0], "-o", SO])
    L = ctypes.CDLL(SO)
    L.probe_fmt_3g_range.restype = ctypes.c_long
    L.probe_fmt_2f_range.restype = ctypes.c_long
    return L


def bits(x):
    return int(np.float32(x).view(np.uint32))


def test_3g_sweep(probe):
    bad = ctypes.c_uint(); unsup = ctypes.c_long()
    assert probe.probe_fmt_3g_range(bits(1e-3), bits(1e12), 61, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    assert unsup.value == 0
    assert probe.probe_fmt_3g_range(bits(20.0), bits(5000.0), 1, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    assert probe.probe_fmt_3g_range(bits(9.9e4), bits(1.1e6), 1, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)


def test_2f_sweep(probe):
    bad = ctypes.c_uint(); unsup = ctypes.c_long()
    assert probe.probe_fmt_2f_range(bits(1e-4), bits(1e9), 53, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    assert unsup.value == 0
    assert probe.probe_fmt_2f_range(bits(0.5), bits(101.0), 1, ctypes.byref(bad), ctypes.byref(unsup)) == 0, hex(bad.value)
    out = ctypes.create_string_buffer(64)
    for x in (0.0, -0.0, 100.0, 99.995, 0.005, 0.015, -3.14159, 1e-30):
        n = probe.probe_fmt_2f(ctypes.c_float(x), out)
        assert n >= 0 and out.value.decode() == "%.2f" % float(np.float32(x)), x


def test_2g_from_text(probe):
    rng = np.random.default_rng(8)
    cases = ["0", "0.0", "1e-50", "6.5e-21", "0.001", "2e-05", "2.0e-5", "1", "10", "12", "99", "100", "1234", "0.5",
             "0.00012", "9.96", "9.94", "0.0996", "9.96e-10", "1.5e+03", "7e22", "1e-400", "-3.5e-7", "1.25e-05",
             "1.35", "12.5", "0.000125", "99.5e3", "1e-4", "1e-5", "0.0001", "0.00001", "123456789012345e-30"]
    for _ in range(40000):
        nd = int(rng.integers(1, 7))
        cases.append("%.*e" % (nd - 1, 10.0 ** rng.uniform(-200, 30)))
    for _ in range(30000):
        cases.append("%.*f" % (int(rng.integers(0, 8)), rng.uniform(0, 2000)))
    out = ctypes.create_string_buffer(64)
    handled = 0
    for s in cases:
        n = probe.probe_fmt_2g_text(s.encode(), out)
        assert n != -2, s
        if n == -1:
            continue
        handled += 1
        want = "%.2g" % libc.strtod(s.encode(), None)
        assert out.value.decode() == want, (s, out.value, want)
    assert handled > 0.97 * len(cases)


def test_whole_line_matches_sscanf_sprintf(probe):
    rng = np.random.default_rng(9)
    out = ctypes.create_string_buffer(512)
    q = ctypes.create_string_buffer(256); sb = ctypes.create_string_buffer(256)
    gs = ctypes.c_float(); ge = ctypes.c_double(); ints = (ctypes.c_int * 7)()
    libc.snprintf.restype = ctypes.c_int
    n_ok = 0
    for i in range(20000):
        sc = "%.1f" % rng.uniform(20, 3000) if i % 3 else "%d" % rng.integers(20, 100000)
        ev = "%.1e" % 10.0 ** rng.uniform(-180, 1) if i % 5 else "0.0"
        f = ["Ag%06d.1" % rng.integers(0, 99999), "Bg%06d.1" % rng.integers(0, 99999), "%.2f" % rng.uniform(20, 100)]
        f += [str(rng.integers(0, 5000)) for _ in range(7)]
        f += [ev, sc]
        if i % 7 == 0:
            f = f[:int(rng.integers(2, 12))]                # short line: remaining fields print as zeros
        if i % 11 == 0 and len(f) == 12:
            f[6], f[7] = str(5000 + i), str(10)           # start > stop: normalised
        ln = "\t".join(f)
        k = probe.probe_fmt_line(ln.encode(), len(ln), out)
        if k < 0:
            continue
        n_ok += 1
        # the reference: sscanf + start/stop swap + sprintf
        pct = ctypes.c_float()
        libc.sscanf((ln + "\n").encode(), b"%127s\t%127s\t%f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%lf\t%f", q, sb, ctypes.byref(pct),
                    *[ctypes.byref(ctypes.c_int.from_buffer(ints, 4 * j)) for j in range(7)], ctypes.byref(ge), ctypes.byref(gs))
        v = list(ints)
        if v[3] > v[4]: v[3], v[4] = v[4], v[3]
        if v[5] > v[6]: v[5], v[6] = v[6], v[5]
        want = "\t%.2f\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2g\t%.3g\n" % (pct.value, *v, ge.value, gs.value)
        assert out.value.decode() == want, (ln, out.value, want)
        for c in (pct, gs): c.value = 0
        ge.value = 0
        for j in range(7): ints[j] = 0
    assert n_ok > 19000

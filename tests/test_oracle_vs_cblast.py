"""The oracle's BlastLine parse/format and the product's host-built tokenizer arithmetic against
the REFERENCE'S OWN compiled code: oracle/_ref/cblast*.so is built from
/root/reference/src/jcvi/formats/cblast.pyx by oracle/build_ref.py (setup.py:16-20 recipe) and
travels to the GPU box.  Skipped when it has not been built."""
import ctypes
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def cblast():
    if not glob.glob(os.path.join(REFDIR, "cblast*.so")):
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import build_ref
            if build_ref.build() is None:
                pytest.skip("reference sources not present and oracle/_ref not built")
        except Exception as e:  # no Cython / compiler
            pytest.skip("cannot build oracle/_ref: %s" % e)
    sys.path.insert(0, REFDIR)
    import cblast as m
    return m


def _lines():
    rng = np.random.default_rng(17)
    out = ["Os09g11510\tOs08g13650\t92.31\t39\t3\t0\t2273\t2311\t3237\t3199\t0.001\t54.0"]   # cblast.pyx:71-77
    for i in range(20000):
        sc = ["%.1f" % rng.uniform(20, 3000), "%d" % rng.integers(20, 200000), "%.3g" % rng.uniform(1, 1e6),
              "%.2e" % rng.uniform(1, 1e5)][i % 4]
        ev = ["%.1e" % 10.0 ** rng.uniform(-250, 1), "0.0", "%.3f" % rng.uniform(0, 1), "1e-10", "%.2e" % 10.0 ** rng.uniform(-12, -8)][i % 5]
        f = ["g%05d.%d" % (rng.integers(0, 99999), rng.integers(1, 4)), "h%05d" % rng.integers(0, 99999), "%.2f" % rng.uniform(20, 100)]
        f += [str(rng.integers(0, 5000)) for _ in range(7)] + [ev, sc]
        if i % 9 == 0:
            f[6], f[7] = str(int(f[7]) + 5000), f[6]
        if i % 13 == 0:
            f[8], f[9] = str(int(f[9]) + 7000), f[8]
        out.append("\t".join(f))
    return out


def test_oracle_blastline_str_equals_reference_cblast(cblast):
    import oracle
    for ln in _lines():
        assert oracle.blastline_str(ln) == str(cblast.BlastLine(ln)), ln


def test_product_tokenizer_arithmetic_equals_reference_cblast(cblast):
    so = os.path.join(ROOT, "jcvi_b200", "libjx_hostprobe.so")
    if not os.path.exists(so):
        pytest.skip("host probe not built")
    L = ctypes.CDLL(so)
    out = ctypes.create_string_buffer(512)
    for ln in _lines():
        b = cblast.BlastLine(ln)
        spans = (ctypes.c_int * 4)(); score = ctypes.c_float(); nconv = ctypes.c_int(); eflag = ctypes.c_int()
        assert L.probe_line(ln.encode(), len(ln), spans, ctypes.byref(score), ctypes.byref(nconv), ctypes.byref(eflag)) == 0
        assert ln[spans[0]:spans[1]] == b.query and ln[spans[2]:spans[3]] == b.subject
        assert np.float32(score.value) == np.float32(b.score), ln              # float32 like cblast.pyx:88
        assert eflag.value == (1 if b.evalue < 1e-10 else 0), ln                # tandem_grouper's test
        k = L.probe_fmt_line(ln.encode(), len(ln), out)                         # the device writer's arithmetic
        if k >= 0:
            assert b.query + "\t" + b.subject + out.value.decode() == str(b) + "\n", ln

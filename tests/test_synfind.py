"""`jcvi.compara.synfind` (src/jcvi/compara/synfind.py:45-240): the oracle's restatement against
golden vectors produced by the unmodified reference (tests/golden/make_golden_synfind.py).  The GPU drop-in for this
module (jcvi_b200/compara/synfind.py, jx_synfind_text) is checked against the same golden files and against the
oracle on larger tables whose windows land in every size class of the kernel."""
import os

import pytest

from conftest import read


def test_oracle_synfind_matches_reference_golden(golden):
    import oracle
    d, cases = golden("synfind")
    assert len(cases) >= 6
    cwd = os.getcwd()
    os.chdir(d)                     # the notes columns carry the BED file names as given
    try:
        for c in cases:
            kw = {}
            for a in c["argv"]:
                k, _, v = a.lstrip("-").partition("=")
                kw[k] = int(v) if k == "window" else float(v)
            oracle.synfind(c["last"], c["qbed"], c["sbed"], c["out"] + ".orc", **kw)
            assert read(c["out"] + ".orc") == read(c["out"]), c
            assert len(read(c["out"])) > 1000
    finally:
        os.chdir(cwd)


def _in_dir(d):
    class _Cd(object):
        def __enter__(self):
            self.cwd = os.getcwd()
            os.chdir(d)

        def __exit__(self, *a):
            os.chdir(self.cwd)
    return _Cd()


@pytest.mark.gpu
def test_gpu_synfind_matches_reference_golden(golden):
    from jcvi_b200.compara import synfind
    d, cases = golden("synfind")
    with _in_dir(d):
        for c in cases:
            synfind.main([c["last"], "--qbed=" + c["qbed"], "--sbed=" + c["sbed"], "-o", c["out"] + ".gpu"] + c["argv"])
            assert read(c["out"] + ".gpu") == read(c["out"]), c


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["sparse", "dense", "self", "huge_window"])
def test_gpu_synfind_matches_oracle(tmp_path, shape):
    """Windows of a few hundred points (shared-memory class 0), of several thousand (classes 1 and 2: a small gene set
    with many hits per gene), beyond 16 k points (the global-scratch class), and a self comparison."""
    import oracle
    from jcvi_b200 import synth
    from jcvi_b200.compara import synfind
    d = str(tmp_path)
    kw = {}
    if shape == "sparse":
        synth.write_dataset(d, "A", "B", 91, 6000, 5000, 120000, Kq=5, Ks=4)
        argv = []
    elif shape == "dense":
        synth.write_dataset(d, "A", "B", 92, 1500, 1200, 400000, Kq=3, Ks=3)
        argv = ["--window=30", "--cutoff=0.2"]
        kw = dict(window=30, cutoff=0.2)
    elif shape == "huge_window":
        synth.write_dataset(d, "A", "B", 94, 700, 600, 250000, Kq=2, Ks=2)
        argv = ["--window=120", "--cutoff=0.05"]
        kw = dict(window=120, cutoff=0.05)
    else:
        synth.write_dataset(d, "A", "A", 93, 4000, 4000, 150000, Kq=4, Ks=4, self_mode=True)
        argv = []
    sb = "A.bed" if shape == "self" else "B.bed"
    last = "A.A.last" if shape == "self" else "A.B.last"
    with _in_dir(d):
        synfind.main([last, "--qbed=A.bed", "--sbed=" + sb, "-o", "gpu.synfind"] + argv)
        oracle.synfind(last, "A.bed", sb, "orc.synfind", **kw)
        got, want = read("gpu.synfind"), read("orc.synfind")
        assert len(want) > 10000
        assert got == want


@pytest.mark.gpu
def test_gpu_synfind_sqlite_batch_query_and_errors(golden, tmp_path):
    import sqlite3
    from jcvi_b200.compara import synfind
    from jcvi_b200.compara.synteny import _load_bed
    d, cases = golden("synfind")
    c = cases[0]
    with _in_dir(d):
        synfind.main([c["last"], "--qbed=" + c["qbed"], "--sbed=" + c["sbed"], "--sqlite=out.db", "--qnote=7", "--snote=9"])
        rows = sqlite3.connect("out.db").execute("select * from synteny").fetchall()
        want = [ln.split("\t") for ln in read(c["out"]).decode().splitlines()]
        assert len(rows) == len(want)
        qn, sn = want[0][6], want[0][7]
        for r, w in zip(rows, want):
            assert [str(x) for x in r[:6]] == w[:6]
            assert (r[6], r[7]) == (("7", "9") if (w[6], w[7]) == (qn, sn) else ("9", "7"))
        # batch_query on the points of the file == the two halves of the file
        qbed, sbed = _load_bed(c["qbed"]), _load_bed(c["sbed"])
        pts = set()
        from jcvi_b200.formats.blast import gene_name
        for ln in read(c["last"]).decode().splitlines():
            if not ln or ln[0] == "#":
                continue
            f = ln.split("\t")
            q, s = gene_name(f[0].encode()).decode(), gene_name(f[1].encode()).decode()
            if q in qbed.order and s in sbed.order:
                pts.add((qbed.order[q][0], sbed.order[s][0]))

        class O(object):
            window, cutoff, scoring, qnote, snote = 40, 0.1, "collinear", "null", "null"
        with open("bq.out", "w") as fw:
            synfind.batch_query(qbed, sbed, sorted(pts), O, fw=fw)
            synfind.batch_query(qbed, sbed, sorted(pts), O, fw=fw, transpose=True)
        assert read("bq.out") == read(c["out"])
        with pytest.raises(UnboundLocalError):
            synfind.main([c["last"], "--qbed=" + c["qbed"], "--sbed=" + c["sbed"], "--scoring=density", "-o", "x"])
        open("empty.last", "w").close()
        with pytest.raises(ValueError):
            synfind.main(["empty.last", "--qbed=" + c["qbed"], "--sbed=" + c["sbed"], "-o", "y"])

"""`jcvi.compara.synteny simple` (second half; src/jcvi/compara/synteny.py:1137-1316): the oracle's
restatement against golden vectors produced by the unmodified reference (tests/golden/make_golden_simple.py).
The GPU drop-in for this action is not built yet; this pins its checker first."""
import os

from conftest import read


def test_oracle_simple_matches_reference_golden(golden):
    import oracle
    d, cases = golden("simple")
    assert len(cases) >= 12
    cwd = os.getcwd()
    os.chdir(d)                     # the block ids carry the anchors path as given ("A-B-block-0")
    try:
        for c in cases:
            kw = dict(rich="--rich" in c["argv"], coords="--coords" in c["argv"], noheader="--noheader" in c["argv"],
                      bed="--bed" in c["argv"])
            out = c["out"] + ".orc"
            assert oracle.simple(c["anchors"], c["qbed"], c["sbed"], out, **kw)
            assert read(out) == read(c["out"]), c
            if c["bed_out"]:
                assert read(out + ".bed") == read(c["bed_out"]), c
        open("empty.anchors", "w").close()
        assert oracle.simple("empty.anchors", "A.bed", "B.bed", "empty.simple") is False
    finally:
        os.chdir(cwd)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_simple_matches_reference_golden(golden):
    """The drop-in (`jcvi_b200.compara.synteny.simple`: per-block reductions on the device, jx_block_stats) against
    the reference's files, every mode, pair and self comparison."""
    from jcvi_b200.compara import synteny as gsyn
    d, cases = golden("simple")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for c in cases:
            gsyn.simple([c["anchors"], "--qbed=" + c["qbed"], "--sbed=" + c["sbed"]] + c["argv"])
            produced = c["anchors"].rsplit(".", 1)[0] + ".simple"
            assert read(produced) == read(c["out"]), c
            if c["bed_out"]:
                assert read(produced + ".bed") == read(c["bed_out"]), c
        with pytest.raises(KeyError):
            open("bad.anchors", "w").write("###\nnobody\tnothing\t10\n")
            gsyn.simple(["bad.anchors", "--qbed=A.bed", "--sbed=B.bed"])
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_gpu_simple_matches_oracle_on_random_blocks(tmp_path):
    """Random small blocks incl. single anchors, repeated genes and exactly flat blocks (orientation column excluded
    there: the reference's own answer is rounding noise, see tests/live_reference_check.py)."""
    import random
    import oracle
    from jcvi_b200.compara import synteny as gsyn
    d = str(tmp_path)
    cwd = os.getcwd()
    os.chdir(d)
    try:
        rng = random.Random(9)
        G = 400
        for nm in ("x", "y"):
            with open(nm + ".bed", "w") as fw:
                for i in range(G):
                    fw.write("%schr%d\t%d\t%d\t%s%03d\n" % (nm, 1 + i // 100, 100 * i, 100 * i + 40 + 7 * (i % 5), nm, i))
        flat = []
        with open("x.y.anchors", "w") as fw:
            for b in range(600):
                fw.write("###\n")
                lo = 100 * rng.randint(0, 3)
                n = rng.choice([1, 2, 3, 5, 9, 40])
                xs = [lo + rng.randint(0, 99) for _ in range(n)]
                ys = [lo + rng.randint(0, 99) for _ in range(n)]
                if rng.random() < 0.15:
                    ys = [ys[0]] * n
                for x, y in zip(xs, ys):
                    fw.write("x%03d\ty%03d\t%d\n" % (x, y, rng.randint(10, 500)))
                flat.append(n >= 2 and n * sum(a * b for a, b in zip(xs, ys)) - sum(xs) * sum(ys) == 0)
        for argv, kw, per, col in ((["--rich"], dict(rich=True), 1, 5), (["--coords"], dict(coords=True), 2, 8)):
            gsyn.simple(["x.y.anchors", "--qbed=x.bed", "--sbed=y.bed"] + argv)
            oracle.simple("x.y.anchors", "x.bed", "y.bed", "orc.simple", **kw)
            got, want = read("x.y.simple").decode().split("\n"), read("orc.simple").decode().split("\n")
            assert len(got) == len(want) and got[0] == want[0]
            for k, (a, b) in enumerate(zip(got[1:], want[1:])):
                if a != b:
                    assert flat[k // per], (a, b)
                    fa, fb = a.split("\t"), b.split("\t")
                    assert fa[:col] + fa[col + 1:] == fb[:col] + fb[col + 1:], (a, b)
    finally:
        os.chdir(cwd)

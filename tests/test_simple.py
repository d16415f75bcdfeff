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

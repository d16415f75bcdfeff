"""`jcvi.compara.synfind` (src/jcvi/compara/synfind.py:45-240): the oracle's restatement against
golden vectors produced by the unmodified reference (tests/golden/make_golden_synfind.py).  The GPU drop-in for this
module is the next row to build; its checker is pinned first."""
import os

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

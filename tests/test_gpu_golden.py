"""GPU parity against the committed golden vectors (outputs of the unmodified reference,
tests/golden/make_golden.py, plus the reference's own golden files).  Everything goes through
the drop-in entry points -> ctypes -> libjcvi_b200.so (C ABI)."""
import os
import shutil

import pytest

from conftest import read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from jcvi_b200.compara import blastfilter, synteny
    return blastfilter, synteny


def test_reference_scan_golden(golden, mods):
    _, syn = mods
    d, _ = golden("ref_fixtures")
    out = os.path.join(d, "o.anchors")
    assert syn.scan([os.path.join(d, "testtigs.testchr.last"), out]) == out
    assert read(out) == read(os.path.join(d, "golden.anchors"))


def test_reference_liftover_golden(golden, mods):
    _, syn = mods
    d, _ = golden("ref_fixtures")
    out = os.path.join(d, "o.lifted.anchors")
    r = syn.liftover([os.path.join(d, "test_empty_blocks.last"), os.path.join(d, "test_empty_blocks.anchors"),
                      "--qbed=" + os.path.join(d, "test_empty_blocks.a.bed"),
                      "--sbed=" + os.path.join(d, "test_empty_blocks.b.bed"), "--outfile=" + out])
    assert r == out
    assert read(out) == read(os.path.join(d, "golden.lifted.anchors"))


@pytest.mark.parametrize("name", ["syn_pair", "syn_self", "edge_lines"])
def test_stages_against_reference_outputs(golden, mods, name):
    bf, syn = mods
    d, cases = golden(name)
    assert cases
    for c in cases:
        j = lambda f: os.path.join(d, f)
        beds = ["--qbed=" + j(c["qbed"]), "--sbed=" + j(c["sbed"])]
        # stage 1 -- writes <last>.filtered next to the input
        bf_args = [a.replace("--exclude=", "--exclude=" + d + os.sep) if a.startswith("--exclude=") else a
                   for a in c["bf_args"]]
        bf.main([j(c["last"])] + beds + bf_args)
        assert read(j(c["last"] + ".filtered")) == read(j(c["filtered"])), (name, c["tag"], "filtered")
        # stage 2 on the reference's filtered file
        out = j("gpu." + c["anchors"])
        err = None
        try:
            syn.scan([j(c["filtered"]), out] + beds + c["scan_args"])
        except ValueError as e:
            err = str(e)
        assert err == c["scan_error"]
        assert read(out) == read(j(c["anchors"])), (name, c["tag"], "anchors")
        # stage 3 on the reference's anchors
        if c.get("lifted"):
            out = j("gpu." + c["lifted"])
            syn.liftover([j(c["last"]), j(c["anchors"])] + beds + ["--outfile=" + out] + c["lift_args"])
            assert read(out) == read(j(c["lifted"])), (name, c["tag"], "lifted")


def test_chained_pipeline_with_liftover_option(golden, mods):
    """scan --liftover=<raw last> (what catalog.ortholog calls, catalog.py:736-750)."""
    bf, syn = mods
    d, cases = golden("syn_pair")
    c = cases[0]
    j = lambda f: os.path.join(d, f)
    bf.main([j(c["last"]), "--cscore=0.7"])
    r = syn.scan([j(c["last"] + ".filtered"), j("A.B.anchors"), "--liftover=" + j(c["last"]), "--liftover_dist=10"])
    assert r == j("A.B.lifted.anchors")
    assert read(j("A.B.anchors")) == read(j(c["anchors"]))
    assert read(r) == read(j(c["lifted"]))

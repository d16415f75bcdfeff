"""Pins the CPU oracle (oracle/jcvi_oracle.cpp) byte-for-byte against
 (a) the reference's own golden files (tests/compara/synteny.py/references/*), and
 (b) outputs of the unmodified reference run in the build container
     (tests/golden/make_golden.py; Cython cblast code path).
"""
import os

import pytest

import oracle
from conftest import read


def test_reference_scan_golden(golden):
    d, _ = golden("ref_fixtures")
    out = os.path.join(d, "o.anchors")
    st = oracle.scan(os.path.join(d, "testtigs.testchr.last"), out, os.path.join(d, "testtigs.bed"),
                     os.path.join(d, "testchr.bed"))
    assert st == {"hits": 1000, "clusters": 70, "anchors": 965}
    assert read(out) == read(os.path.join(d, "golden.anchors"))


def test_reference_liftover_golden(golden):
    d, _ = golden("ref_fixtures")
    out = os.path.join(d, "o.lifted.anchors")
    st = oracle.liftover(os.path.join(d, "test_empty_blocks.last"), os.path.join(d, "test_empty_blocks.anchors"),
                         os.path.join(d, "test_empty_blocks.a.bed"), os.path.join(d, "test_empty_blocks.b.bed"),
                         out_path=out)
    assert st["lifted"] == 0 and st["removed"] == 2
    assert read(out) == read(os.path.join(d, "golden.lifted.anchors"))


def _opt(args, name, default, cast=float):
    for i, a in enumerate(args):
        if a.startswith("--" + name + "="):
            return cast(a.split("=", 1)[1])
        if a in ("--" + name, "-" + name) and i + 1 < len(args):
            return cast(args[i + 1])
    return default


def run_oracle_case(d, c, outdir=None):
    """Run the three stages of one golden case through the oracle; returns output paths."""
    outdir = outdir or d
    j = lambda f: os.path.join(d, f)
    o = lambda f: os.path.join(outdir, "orc." + f)
    bfa, sca, lfa = c["bf_args"], c["scan_args"], c["lift_args"]
    excl = _opt(bfa, "exclude", None, str)
    oracle.blastfilter(j(c["last"]), j(c["qbed"]), j(c["sbed"]), out_path=o(c["filtered"]),
                       strip_names="--no_strip_names" not in bfa, cscore=_opt(bfa, "cscore", 0.7),
                       tandem_Nmax=_opt(bfa, "tandem_Nmax", 10, int), exclude=j(excl) if excl else None)
    res = {"filtered": o(c["filtered"])}
    # like the golden generator, scan reads the REFERENCE's filtered file name (bed inference is by argv)
    try:
        oracle.scan(res["filtered"], o(c["anchors"]), j(c["qbed"]), j(c["sbed"]), dist=_opt(sca, "dist", 20, int),
                    min_size=_opt(sca, "n", 4, int), intrabound=_opt(sca, "intrabound", 300, int))
        res["scan_error"] = None
    except oracle.ZeroAnchors as e:
        res["scan_error"] = str(e)
    res["anchors"] = o(c["anchors"])
    if res["scan_error"] is None and c.get("lifted"):
        oracle.liftover(j(c["last"]), res["anchors"], j(c["qbed"]), j(c["sbed"]), out_path=o(c["lifted"]),
                        strip_names="--no_strip_names" not in lfa, dist=_opt(lfa, "dist", 10, int))
        res["lifted"] = o(c["lifted"])
    return res


@pytest.mark.parametrize("name", ["syn_pair", "syn_self", "edge_lines"])
def test_oracle_matches_reference_outputs(golden, name):
    d, cases = golden(name)
    assert cases
    for c in cases:
        res = run_oracle_case(d, c)
        assert read(res["filtered"]) == read(os.path.join(d, c["filtered"])), (name, c["tag"], "filtered")
        assert res["scan_error"] == c["scan_error"]
        assert read(res["anchors"]) == read(os.path.join(d, c["anchors"])), (name, c["tag"], "anchors")
        if c.get("lifted"):
            assert read(res["lifted"]) == read(os.path.join(d, c["lifted"])), (name, c["tag"], "lifted")


def test_gene_name_and_natsort():
    # utils/cbook.py:308-329
    assert oracle.gene_name("Ag000123.1") == "Ag000123"
    assert oracle.gene_name("Ag000123.12") == "Ag000123.12"
    assert oracle.gene_name("evm.model.x.5") == "evm.model.x.5"
    assert oracle.gene_name("a.b.c|x.1") == "a.b"
    assert oracle.gene_name("abc") == "abc"
    assert oracle.gene_name("abc.") == "abc."
    # natsort documented behaviour
    assert oracle.natcmp("chr2", "chr10") < 0
    assert oracle.natcmp("chr10", "chr2") > 0
    assert oracle.natcmp("chr1", "chr01") == 0
    assert oracle.natcmp("1x", "x1") < 0
    assert oracle.natcmp("a1b2", "a1b10") < 0
    assert oracle.natcmp("a", "a1") < 0


def test_blastline_str_known_answers():
    # cblast.pyx:71-77 docstring example + SURVEY probes (float32 score, %.3g)
    s = oracle.blastline_str("Os09g11510\tOs08g13650\t92.31\t39\t3\t0\t2273\t2311\t3237\t3199\t0.001\t54.0")
    assert s == "Os09g11510\tOs08g13650\t92.31\t39\t3\t0\t2273\t2311\t3199\t3237\t0.001\t54"
    s = oracle.blastline_str("a\tb\t99.5\t1\t2\t3\t4\t5\t6\t7\t2e-5\t1234.7")
    assert s.endswith("\t2e-05\t1.23e+03")
    s = oracle.blastline_str("a\tb\t99.5\t1\t2\t3\t4\t5\t6\t7\t0\t99999")
    assert s.endswith("\t0\t1e+05")

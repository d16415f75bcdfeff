"""`jcvi.compara.synteny mcscan` (src/jcvi/compara/synteny.py:1538-1652).

CPU: the oracle's restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden_mcscan.py).  GPU: the drop-in (`jcvi_b200.compara.synteny.mcscan`) against the same
vectors and against the oracle on a larger pair."""
import os

import pytest

from conftest import read


def _kwargs(argv, d):
    kw = dict(iter=100, ascii=False, Nm=10, trackids=False, mergetandem=None)
    for a in argv:
        k, _, v = a.lstrip("-").partition("=")
        if k in ("ascii", "trackids"):
            kw[k] = True
        elif k == "mergetandem":
            kw[k] = os.path.join(d, v)
        else:
            kw[k] = int(v)
    return kw


def test_oracle_mcscan_matches_reference_golden(golden):
    import oracle
    d, cases = golden("mcscan")
    assert len(cases) >= 9
    for c in cases:
        out = os.path.join(d, c["out"] + ".orc")
        oracle.mcscan(os.path.join(d, c["bed"]), os.path.join(d, c["anchors"]), out, **_kwargs(c["argv"], d))
        assert read(out) == read(os.path.join(d, c["out"])), c
        if c["tracks"]:
            assert read(out + ".tracks") == read(os.path.join(d, c["tracks"])), c
    with pytest.raises(KeyError):          # a block whose genes are in neither BED column order
        bad = os.path.join(d, "bad.anchors")
        open(bad, "w").write("###\nnobody\tnothing\t10\n")
        oracle.mcscan(os.path.join(d, "A.bed"), bad, bad + ".blocks")


@pytest.mark.gpu
def test_gpu_mcscan_matches_reference_golden(golden):
    from jcvi_b200.compara import synteny as gsyn
    d, cases = golden("mcscan")
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for c in cases:
            out = c["out"] + ".gpu"
            gsyn.mcscan([c["bed"], c["anchors"], "-o", out] + c["argv"])
            assert read(out) == read(c["out"]), c
            if c["tracks"]:
                assert read(out + ".tracks") == read(c["tracks"]), c
    finally:
        os.chdir(cwd)


@pytest.mark.gpu
def test_gpu_mcscan_matches_oracle_on_a_larger_pair(tmp_path):
    """The whole path on the GPU (blastfilter -> scan --liftover) and then mcscan from both sides, against the oracle's
    mcscan on the same lifted anchors."""
    import oracle
    from jcvi_b200 import synth
    from jcvi_b200.compara import blastfilter as gbf, synteny as gsyn
    d = str(tmp_path)
    qbed, sbed, last = synth.write_dataset(d, "A", "B", 71, 12000, 4000, 400000, Kq=8, Ks=4)
    gbf.main([last, "--cscore=0.7"])
    lifted = gsyn.scan([last + ".filtered", os.path.join(d, "A.B.anchors"), "--liftover=" + last, "--min_size=3"])
    assert lifted.endswith(".lifted.anchors")
    for bed, argv, kw in ((qbed, [], {}), (sbed, ["--trackids", "--Nm=6"], dict(trackids=True, Nm=6)),
                          (sbed, ["--iter=3", "--ascii"], dict(iter=3, ascii=True))):
        tag = os.path.basename(bed) + "".join(argv)
        out = os.path.join(d, tag + ".blocks")
        gsyn.mcscan([bed, lifted, "-o", out] + argv)
        oracle.mcscan(bed, lifted, out + ".orc", **kw)
        assert read(out) == read(out + ".orc"), tag
        assert len(read(out)) > 10000
        if kw.get("trackids"):
            assert read(out + ".tracks") == read(out + ".orc.tracks")
    with pytest.raises(KeyError):
        bad = os.path.join(d, "bad.anchors")
        open(bad, "w").write("###\nnobody\tnothing\t10\n")
        gsyn.mcscan([qbed, bad, "-o", bad + ".blocks"])

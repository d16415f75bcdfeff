"""`jcvi.formats.blast cscore` (src/jcvi/formats/blast.py:1098-1189).

CPU: the oracle's restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden_cscore.py).  GPU: the drop-in (`jcvi_b200.formats.blast.cscore`) against the same
vectors and against the oracle on a larger synthetic table."""
import os

import numpy as np
import pytest

from conftest import read


def _argv_to_kwargs(argv):
    kw = dict(cutoff=0.9999, pct=False, writeblast=False, strip_names=True)
    for a in argv:
        if a.startswith("--cutoff="):
            kw["cutoff"] = float(a.split("=", 1)[1])
        elif a == "--pct":
            kw["pct"] = True
        elif a == "--writeblast":
            kw["writeblast"] = True
        elif a == "--no_strip_names":
            kw["strip_names"] = False
    return kw


def test_oracle_cscore_matches_reference_golden(golden):
    import oracle
    d, cases = golden("cscore")
    assert len(cases) >= 7
    for c in cases:
        inp = os.path.join(d, c["input"])
        if c.get("error"):
            with pytest.raises(ZeroDivisionError):
                oracle.cscore(inp, inp + ".orc")
            continue
        out = os.path.join(d, c["out"] + ".orc")
        oracle.cscore(inp, out, **_argv_to_kwargs(c["argv"]))
        assert read(out) == read(os.path.join(d, c["out"])), c
        if c["blast_out"]:
            assert read(out + ".filtered.blast") == read(os.path.join(d, c["blast_out"])), c


@pytest.mark.gpu
def test_gpu_cscore_matches_reference_golden(golden):
    from jcvi_b200.formats import blast as gblast
    d, cases = golden("cscore")
    for c in cases:
        inp = os.path.join(d, c["input"])
        if c.get("error"):
            with pytest.raises((ZeroDivisionError, NotImplementedError)):     # the product refuses such tables loudly
                gblast.cscore([inp, "-o", inp + ".gpu"])
            continue
        out = os.path.join(d, c["out"] + ".gpu")
        gblast.cscore([inp, "-o", out] + c["argv"])
        assert read(out) == read(os.path.join(d, c["out"])), c
        if c["blast_out"]:
            assert read(out + ".filtered.blast") == read(os.path.join(d, c["blast_out"])), c


@pytest.mark.gpu
def test_gpu_cscore_matches_oracle_on_a_larger_table(tmp_path):
    import oracle
    from jcvi_b200 import synth
    from jcvi_b200.formats import blast as gblast
    d = str(tmp_path)
    gq, gs, text = synth.make_pair(33, 6000, 6000, 400000, Kq=5, Ks=5)
    last = os.path.join(d, "A.B.last")
    open(last, "wb").write(bytes(text))
    for tag, argv in (("rbh", []), ("c05", ["--cutoff=0.5", "--pct", "--writeblast"]), ("ns", ["--no_strip_names", "--cutoff=0.8"])):
        kw = _argv_to_kwargs(argv)
        oracle.cscore(last, last + "." + tag + ".orc", **kw)
        gblast.cscore([last, "-o", last + "." + tag + ".gpu"] + argv)
        assert read(last + "." + tag + ".gpu") == read(last + "." + tag + ".orc"), tag
        assert len(read(last + "." + tag + ".gpu")) > 1000
        if kw["writeblast"]:
            assert read(last + "." + tag + ".gpu.filtered.blast") == read(last + "." + tag + ".orc.filtered.blast")


@pytest.mark.gpu
def test_gpu_cscore_pairs_python_route_equals_one_call(golden):
    """`cscore_pairs` (the same steps from Python, for library use) and `jx_blast_cscore` agree."""
    from jcvi_b200.formats import blast as gblast
    from jcvi_b200.apps_base import read_text_file
    d, cases = golden("cscore")
    c = [x for x in cases if x["argv"] and "--writeblast" in x["argv"]][0]
    inp = os.path.join(d, c["input"])
    pairs = gblast.cscore_pairs(read_text_file(inp), 0.7, True)
    lines = ["%s\t%s\t%.2f\t%.1f" % (q.decode(), s.decode(), v[0], v[1]) for (q, s), v in sorted(pairs.items())]
    assert ("\n".join(lines) + "\n").encode() == read(os.path.join(d, c["out"]))
    assert ("\n".join(v[2].decode() for _, v in sorted(pairs.items())) + "\n").encode() == read(os.path.join(d, c["blast_out"]))

#!/usr/bin/env python
"""Golden vectors for `jcvi.compara.synteny mcscan` (src/jcvi/compara/synteny.py:1538-1652),
produced by running the UNMODIFIED reference (refenv.py) in this container.

    python tests/golden/make_golden_mcscan.py      ->  tests/golden/mcscan.tar.gz

Inputs: lifted anchors of a synthetic two-genome pair (3:1 polyploid shape, so that several tracks fill) and of a
self comparison (blocks are stacked from both sides, base.py:44-49), made by the reference's own scan + liftover;
both BEDs as the reference order.  Option sets: defaults, --iter=2, --Nm=0, --Nm=40, --ascii, --trackids,
--mergetandem."""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refenv  # noqa: E402
from make_golden import pack  # noqa: E402


def main():
    refenv.activate()
    from jcvi.compara import blastfilter as ref_bf, synteny as ref_syn
    from jcvi_b200 import synth

    d = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(d)
    cases = []
    try:
        synth.write_dataset(d, "A", "B", 61, 3000, 1000, 60000, Kq=6, Ks=3)          # three A copies per B gene
        ref_bf.main(["A.B.last", "--cscore=0.7"])
        ref_syn.scan(["A.B.last.filtered", "A.B.anchors", "--liftover=A.B.last", "--min_size=3"])
        synth.write_dataset(d, "S", "S", 62, 2000, 2000, 50000, Kq=5, Ks=5, self_mode=True)
        ref_bf.main(["S.S.last", "--cscore=0.5"])
        ref_syn.scan(["S.S.last.filtered", "S.S.anchors", "--liftover=S.S.last", "--min_size=3"])
        # tandem file for --mergetandem: a few groups of B genes
        names = [ln.split("\t")[3].strip() for ln in open("B.bed") if ln.count("\t") >= 3]
        with open("B.tandems", "w") as fw:
            for i in range(0, 300, 3):
                fw.write("\t".join(names[i:i + 2 + (i % 2)]) + "\n")
        for f in list(os.listdir(d)):
            if f.endswith(".last") or f.endswith(".filtered") or f in ("A.B.anchors", "S.S.anchors"):
                os.remove(f)
        runs = (
            ("A.bed", "A.B.lifted.anchors", "dflt", []),
            ("B.bed", "A.B.lifted.anchors", "bref", []),
            ("A.bed", "A.B.lifted.anchors", "iter2", ["--iter=2"]),
            ("B.bed", "A.B.lifted.anchors", "nm0", ["--Nm=0", "--trackids"]),
            ("B.bed", "A.B.lifted.anchors", "nm40", ["--Nm=40"]),
            ("B.bed", "A.B.lifted.anchors", "ascii", ["--ascii", "--iter=5"]),
            ("A.bed", "A.B.lifted.anchors", "tandem", ["--mergetandem=B.tandems"]),
            ("S.bed", "S.S.lifted.anchors", "self", ["--trackids"]),
            ("S.bed", "S.S.lifted.anchors", "selfnm", ["--Nm=4", "--iter=3"]),
        )
        for bed, anchors, tag, argv in runs:
            out = "%s.%s.blocks" % (anchors, tag)
            ref_syn.mcscan([bed, anchors, "-o", out] + argv)
            cases.append({"bed": bed, "anchors": anchors, "argv": argv, "out": out,
                          "tracks": out + ".tracks" if "--trackids" in argv else None})
    finally:
        os.chdir(cwd)
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    for f in sorted(os.listdir(d)):
        print(f, os.path.getsize(os.path.join(d, f)))
    pack(d, os.path.join(HERE, "mcscan.tar.gz"))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for `jcvi.compara.synteny simple` (src/jcvi/compara/synteny.py:1137-1316),
produced by running the UNMODIFIED reference (refenv.py) in this container.

    python tests/golden/make_golden_simple.py      ->  tests/golden/simple.tar.gz

Inputs: the lifted anchors of tests/golden/mcscan.tar.gz (a 3:1 pair and a self comparison) with their BEDs.
Modes: default, --rich, --coords, --bed, --noheader, --rich --noheader."""
import json
import os
import shutil
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import refenv  # noqa: E402
from make_golden import pack  # noqa: E402


def main():
    refenv.activate()
    from jcvi.compara import synteny as ref_syn
    src = tempfile.mkdtemp()
    tarfile.open(os.path.join(HERE, "mcscan.tar.gz")).extractall(src)
    sub = src if os.path.exists(os.path.join(src, "cases.json")) else os.path.join(src, os.listdir(src)[0])
    d = tempfile.mkdtemp()
    for f in ("A.bed", "B.bed", "S.bed", "A.B.lifted.anchors", "S.S.lifted.anchors"):
        shutil.copy(os.path.join(sub, f), os.path.join(d, f))
    cwd = os.getcwd()
    os.chdir(d)
    cases = []
    try:
        for anchors, qbed, sbed in (("A.B.lifted.anchors", "A.bed", "B.bed"), ("S.S.lifted.anchors", "S.bed", "S.bed")):
            for tag, argv in (("dflt", []), ("rich", ["--rich"]), ("coords", ["--coords"]), ("bed", ["--bed"]),
                              ("nohdr", ["--noheader"]), ("richnohdr", ["--rich", "--noheader"])):
                ref_syn.simple([anchors, "--qbed=" + qbed, "--sbed=" + sbed] + argv)
                produced = anchors.rsplit(".", 1)[0] + ".simple"
                out = "%s.%s.simple" % (anchors, tag)
                os.rename(produced, out)
                bed_out = None
                if "--bed" in argv:
                    bed_out = out + ".bed"
                    os.rename(produced + ".bed", bed_out)
                cases.append({"anchors": anchors, "qbed": qbed, "sbed": sbed, "argv": argv, "out": out, "bed_out": bed_out})
    finally:
        os.chdir(cwd)
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    for f in sorted(os.listdir(d)):
        print(f, os.path.getsize(os.path.join(d, f)))
    pack(d, os.path.join(HERE, "simple.tar.gz"))


if __name__ == "__main__":
    main()

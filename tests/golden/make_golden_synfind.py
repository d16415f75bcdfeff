#!/usr/bin/env python
"""Golden vectors for `jcvi.compara.synfind` (src/jcvi/compara/synfind.py), produced by running the
UNMODIFIED reference (refenv.py) in this container.

    python tests/golden/make_golden_synfind.py      ->  tests/golden/synfind.tar.gz

Inputs: a small synthetic pair and a self comparison (raw LAST tables + BEDs).  Option sets: defaults
(window 40, cutoff 0.1, collinear scoring), --window=20 --cutoff=0.2, --window=60 --cutoff=0.05."""
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
    from jcvi.compara import synfind as ref_sf
    from jcvi_b200 import synth
    d = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(d)
    cases = []
    try:
        synth.write_dataset(d, "A", "B", 81, 1200, 900, 9000, Kq=3, Ks=3)
        synth.write_dataset(d, "S", "S", 82, 1000, 1000, 8000, Kq=3, Ks=3, self_mode=True)
        for last, qbed, sbed in (("A.B.last", "A.bed", "B.bed"), ("S.S.last", "S.bed", "S.bed")):
            for tag, argv in (("dflt", []), ("w20", ["--window=20", "--cutoff=0.2"]), ("w60", ["--window=60", "--cutoff=0.05"])):
                out = "%s.%s.synfind" % (last, tag)
                ref_sf.main([last, "--qbed=" + qbed, "--sbed=" + sbed, "-o", out] + argv)
                cases.append({"last": last, "qbed": qbed, "sbed": sbed, "argv": argv, "out": out})
    finally:
        os.chdir(cwd)
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    for f in sorted(os.listdir(d)):
        print(f, os.path.getsize(os.path.join(d, f)))
    pack(d, os.path.join(HERE, "synfind.tar.gz"))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Golden vectors for `jcvi.formats.blast cscore`, produced by running the
UNMODIFIED reference (Cython cblast build, refenv.py) in this container.

    python tests/golden/make_golden_cscore.py      ->  tests/golden/cscore.tar.gz

Inputs: a synthetic two-genome LAST table (isoform suffixes, multi-HSP duplicates, tandem copies) and an
irregular variant of it ('#' lines mid-file, CRLF, extra columns, spaces as separators, '|' names,
exponent scores, lines that stop after eight fields -> score 0).  Option sets: defaults (RBH, cutoff
0.9999), --cutoff=0.7 --pct --writeblast, --no_strip_names --cutoff=0.5.  A blank line makes the reference
raise ZeroDivisionError (best_score[""] is 0.0); that is recorded too."""
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
    from jcvi.formats import blast as ref_blast
    from jcvi_b200 import synth
    import numpy as np

    d = tempfile.mkdtemp()
    gq, gs, text = synth.make_pair(21, 900, 900, 12000, Kq=3, Ks=3)
    lines = [ln for ln in bytes(text).decode().split("\n") if ln]
    open(os.path.join(d, "A.B.last"), "w").write("\n".join(lines) + "\n")
    out = []
    for i, ln in enumerate(lines):
        r = i % 89
        if r == 5:
            out.append("# comment in the middle")
        if r == 7:
            ln = ln + "\textra\tcolumns"
        if r == 11:                                          # stops after eight fields: score / evalue stay 0; the
            f = ln.split("\t"); f[:2] = lines[i - 1].split("\t")[:2]   # names are scored elsewhere (else the
            ln = "\t".join(f[:8])                            # reference divides by zero, see blank.last)
        if r == 13:
            ln = ln.replace("\t", " ")
        if r == 19:
            f = ln.split("\t"); f[11] = "%.2e" % (float(f[11]) * 37); ln = "\t".join(f)
        if r == 23:
            ln = ln + "\r"
        if r == 29:
            f = ln.split("\t"); f[0] = f[0] + "|alt"; ln = "\t".join(f)
        if r == 31:
            f = ln.split("\t"); f[1] = "ev" + f[1]; ln = "\t".join(f)       # gene_name's exclude list
        if r == 37:
            ln = "  " + ln
        if r == 41:
            f = ln.split("\t"); f[11] = f[11] + "abc"; ln = "\t".join(f)
        if r == 43:
            f = ln.split("\t"); f[0] = "a.very.long.name.with.dots.and.more.than.24.bytes.%d.2" % i; ln = "\t".join(f)
        out.append(ln)
    open(os.path.join(d, "A.B.edge.last"), "w", newline="").write("\n".join(out))   # no trailing newline
    open(os.path.join(d, "blank.last"), "w").write("\n".join(lines[:50]) + "\n\n" + "\n".join(lines[50:100]) + "\n")

    cases = []
    cwd = os.getcwd()
    os.chdir(d)
    try:
        for inp in ("A.B.last", "A.B.edge.last"):
            for tag, argv in (("rbh", []), ("c07", ["--cutoff=0.7", "--pct", "--writeblast"]),
                              ("nostrip", ["--no_strip_names", "--cutoff=0.5", "--pct"])):
                outp = "%s.%s.cscore" % (inp, tag)
                ref_blast.cscore([inp, "-o", outp] + argv)
                cases.append({"input": inp, "argv": argv, "out": outp,
                              "blast_out": outp + ".filtered.blast" if "--writeblast" in argv else None})
        try:
            ref_blast.cscore(["blank.last", "-o", "blank.cscore"])
            err = None
        except ZeroDivisionError as e:
            err = "ZeroDivisionError"
        cases.append({"input": "blank.last", "argv": [], "out": None, "error": err})
    finally:
        os.chdir(cwd)
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    for f in os.listdir(d):
        print(f, os.path.getsize(os.path.join(d, f)))
    pack(d, os.path.join(HERE, "cscore.tar.gz"))


if __name__ == "__main__":
    main()

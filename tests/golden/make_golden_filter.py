#!/usr/bin/env python
"""Golden vectors for `jcvi.formats.blast filter` (src/jcvi/formats/blast.py:620-709),
produced by running the UNMODIFIED reference (Cython cblast build, refenv.py) in this container.

    python tests/golden/make_golden_filter.py      ->  tests/golden/blast_filter.tar.gz

Inputs: a synthetic self comparison (A.A.last: self hits, isoform suffixes), a two-genome table, and an
irregular variant ('#' lines mid-file, blank lines, CRLF and lone CR, trailing white space, spaces as
separators, lines that stop early, exponent scores, pctid / evalue / score written at the cutoffs, leading
'+', many-digit decimals).  Option sets: defaults; catalog.ortholog's `--pctid=98 --inverse --noself`
(catalog.py:712-718); score / hitlen / evalue cutoffs; --ids (commas, '#' rows); --inverse of those."""
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


def edge_lines(lines):
    out = []
    for i, ln in enumerate(lines):
        r = i % 97
        f = ln.split("\t")
        if r == 3:
            out.append("# comment in the middle")
        if r == 5:
            out.append("")                                   # blank row: BlastLine of "\n", every field 0
        if r == 7:
            ln = ln + "\textra\tcolumns  "
        if r == 11:
            ln = "\t".join(f[:8])                            # stops early: evalue / score stay 0
        if r == 13:
            ln = ln.replace("\t", " ")
        if r == 17:
            f[11] = "%.2e" % (float(f[11]) * 37); ln = "\t".join(f)
        if r == 19:
            ln = ln + "\r"                                   # CR LF
        if r == 23:
            ln = ln + " \t "                                 # rstrip
        if r == 29:
            f[2] = "95"; ln = "\t".join(f)                   # pctid at the default cutoff
        if r == 31:
            f[2] = "94.999999"; ln = "\t".join(f)            # float32(94.999999) == 95.0
        if r == 37:
            f[10] = "0.01"; ln = "\t".join(f)                # evalue at the default cutoff
        if r == 41:
            f[10] = "0.010000000000000001"; ln = "\t".join(f)   # rounds to the same double
        if r == 43:
            f[10] = "1.0000000000000002e-2"; ln = "\t".join(f)  # one ulp above
        if r == 47:
            f[3] = "100"; f[2] = "+99.5"; ln = "\t".join(f)
        if r == 53:
            f[3] = "-5"; ln = "\t".join(f)
        if r == 59:
            f[10] = "0"; f[11] = "5e2"; ln = "\t".join(f)
        if r == 61:
            f[1] = f[0]; ln = "\t".join(f)                   # self hit in a two-genome table
        if r == 67:
            ln = "  " + ln
        if r == 71:
            f[11] = f[11] + "abc"; ln = "\t".join(f)
        if r == 73:
            f[10] = "1e-400"; ln = "\t".join(f)              # underflows to 0
        if r == 79:
            f[3] = "0100"; ln = "\t".join(f)
        if r == 83:
            f[10] = "abc"; ln = "\t".join(f)                 # evalue not converted
        out.append(ln)
    # one lone CR inside the file (universal newlines: a row separator)
    out[40] = out[40] + "\r" + out[41]
    del out[41]
    return "\n".join(out)                                    # no trailing newline


def main():
    refenv.activate()
    from jcvi.formats import blast as ref_blast
    from jcvi_b200 import synth

    d = tempfile.mkdtemp()
    _, _, text = synth.make_pair(23, 400, 400, 2600, Kq=3, Ks=3)
    lines = [ln for ln in bytes(text).decode().split("\n") if ln and ln[0] != "#"]
    # raise pctid on part of the table so that the default cutoff keeps something
    for i in range(0, len(lines), 3):
        f = lines[i].split("\t"); f[2] = "%.2f" % (95.0 + (i % 500) / 100.0); lines[i] = "\t".join(f)
    open(os.path.join(d, "A.B.last"), "w").write("# header\n" + "\n".join(lines) + "\n")
    _, _, stext = synth.make_pair(24, 400, 400, 2000, Kq=3, Ks=3, self_mode=True)
    slines = [ln for ln in bytes(stext).decode().split("\n") if ln and ln[0] != "#"]
    for i in range(0, len(slines), 2):
        f = slines[i].split("\t"); f[2] = "%.1f" % (97.0 + (i % 30) / 10.0); slines[i] = "\t".join(f)
    for i in range(5, len(slines), 40):
        f = slines[i].split("\t"); f[1] = f[0]; f[2] = "100.00"; slines[i] = "\t".join(f)     # self hits
    open(os.path.join(d, "A.A.last"), "w").write("\n".join(slines) + "\n")
    open(os.path.join(d, "A.B.edge.last"), "w", newline="").write(edge_lines(lines))
    names = sorted({ln.split("\t")[0] for ln in lines[::2]} | {ln.split("\t")[1] for ln in lines[::3]})
    with open(os.path.join(d, "keep.ids"), "w") as fw:
        fw.write("# ids to retain\n")
        for i in range(0, len(names), 3):
            fw.write(",".join(names[i:i + 3]) + "\n")
        fw.write("one two\tthree\n")

    cases = []
    cwd = os.getcwd()
    os.chdir(d)
    try:
        option_sets = (
            ("dflt", []),
            ("orth", ["--pctid=98", "--inverse", "--noself"]),
            ("cut", ["--score=400", "--pctid=60", "--hitlen=150", "--evalue=1e-30"]),
            ("cutinv", ["--score=400", "--pctid=60", "--hitlen=150", "--evalue=1e-30", "--inverse", "--noself"]),
            ("ids", ["--pctid=0", "--hitlen=0", "--evalue=10", "--ids=keep.ids"]),
            ("idsinv", ["--pctid=50", "--hitlen=0", "--ids=keep.ids", "--inverse"]),
            ("loose", ["--pctid=0", "--hitlen=-10", "--evalue=1e300", "--score=-5"]),
            ("ev0", ["--pctid=0", "--hitlen=0", "--evalue=0"]),
        )
        for inp in ("A.B.last", "A.A.last", "A.B.edge.last"):
            for tag, argv in option_sets:
                outp = "%s.%s" % (inp, tag)
                got = ref_blast.filter([inp, "-o", outp] + argv)
                cases.append({"input": inp, "argv": argv, "outfile": outp, "out": os.path.basename(got)})
    finally:
        os.chdir(cwd)
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    for f in sorted(os.listdir(d)):
        print(f, os.path.getsize(os.path.join(d, f)))
    pack(d, os.path.join(HERE, "blast_filter.tar.gz"))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate tests/golden/*.tar.gz by running the UNMODIFIED reference (tanghaibao/jcvi,
Cython cblast build, see refenv.py) in this container.

    python tests/golden/make_golden.py

Each case directory holds the inputs (BEDs, .last, optional anchors/exclude file) and the
reference's outputs (`*.filtered`, `*.anchors`, `*.lifted.anchors`), plus `case.json` with
the argv used.  Cases:

  ref_scan / ref_liftover  the reference's own test fixtures and golden files
                           (tests/compara/synteny.py/{inputs,references})
  syn_pair                 synthetic 2-genome pair, default options, full pipeline
  syn_pair_opts            same inputs: --cscore=0.5 --tandem_Nmax=3; --cscore=0 --tandem_Nmax=0;
                           --no_strip_names; --exclude; scan -n 2 --dist=8
  syn_self                 self comparison (A.A.last): swap, intrabound, si-qi<40 rule
  syn_ties                 liftover with many equidistant anchors in different blocks
                           (anchors produced with scan --dist=2 -n 2, liftover --dist=10)
  edge_lines               irregular lines: blank, CRLF, extra columns, short lines, '#' mid-file,
                           exponent scores, spaces as separators
"""
import io
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


def run_case(bf, syn, d, name, q, s, bf_args=(), scan_args=(), lift_args=(), lift_dist=None, tag=None):
    """Full pipeline in directory d on <q>.<s>.last; outputs get `<name>.` inserted."""
    tag = name + "."
    cwd = os.getcwd()
    os.chdir(d)
    try:
        last = "%s.%s.last" % (q, s)
        rec = {"last": last, "qbed": q + ".bed", "sbed": s + ".bed", "bf_args": list(bf_args),
               "scan_args": list(scan_args), "lift_args": list(lift_args), "tag": tag}
        bf.main([last] + list(bf_args))
        filt = "%s.%sfiltered" % (last, tag)
        shutil.move(last + ".filtered", filt)
            # scan infers beds from the file name X.Y.*; keep that shape
        rec["filtered"] = filt
        anchors = "%s.%s.%sanchors" % (q, s, tag)
        err = None
        try:
            syn.scan([filt, anchors, "--qbed=%s.bed" % q, "--sbed=%s.bed" % s] + list(scan_args))
        except ValueError as e:
            err = str(e)
        rec["anchors"] = anchors
        rec["scan_error"] = err
        if err is None:
            lifted = "%s.%s.%slifted.anchors" % (q, s, tag)
            try:
                syn.liftover([last, anchors, "--qbed=%s.bed" % q, "--sbed=%s.bed" % s, "--outfile=" + lifted] + list(lift_args))
            except ValueError as e:
                rec["lift_error"] = str(e)
            rec["lifted"] = lifted
        return rec
    finally:
        os.chdir(cwd)


def pack(d, out):
    with tarfile.open(out, "w:gz", compresslevel=9) as tf:
        for f in sorted(os.listdir(d)):
            ti = tf.gettarinfo(os.path.join(d, f), arcname=f)
            ti.mtime = 0; ti.uid = ti.gid = 0; ti.uname = ti.gname = ""
            with open(os.path.join(d, f), "rb") as fh:
                tf.addfile(ti, fh)
    print("wrote", out, os.path.getsize(out), "bytes")


def main():
    bf, syn = refenv.activate()
    from jcvi_b200 import synth
    T = os.path.join(refenv.REFERENCE, "tests", "compara", "synteny.py")

    # ---- the reference's own fixtures -------------------------------------------------
    d = tempfile.mkdtemp()
    for f in ("testtigs.bed", "testchr.bed", "testtigs.testchr.last"):
        shutil.copy(os.path.join(T, "inputs", f), d)
    shutil.copy(os.path.join(T, "references", "testtigs.testchr.anchors"), os.path.join(d, "golden.anchors"))
    for f in ("test_empty_blocks.a.bed", "test_empty_blocks.b.bed", "test_empty_blocks.last", "test_empty_blocks.anchors"):
        shutil.copy(os.path.join(T, "inputs", f), d)
    shutil.copy(os.path.join(T, "references", "test_empty_blocks.lifted.anchors"), os.path.join(d, "golden.lifted.anchors"))
    shutil.copy(os.path.join(refenv.REFERENCE, "tests", "compara", "test.anchors"), os.path.join(d, "anchorfile_test.anchors"))
    for f in os.listdir(d):
        os.chmod(os.path.join(d, f), 0o644)
    pack(d, os.path.join(HERE, "ref_fixtures.tar.gz"))

    # ---- synthetic pair ------------------------------------------------------------------
    d = tempfile.mkdtemp()
    synth.write_dataset(d, "A", "B", seed=11, Gq=1500, Gs=1400, n_lines=20000, Kq=11, Ks=12)
    cases = [run_case(bf, syn, d, "default", "A", "B", bf_args=["--cscore=0.7"])]
    cases.append(run_case(bf, syn, d, "c05n3", "A", "B", bf_args=["--cscore=0.5", "--tandem_Nmax=3"], tag="c05n3."))
    cases.append(run_case(bf, syn, d, "nofilter", "A", "B", bf_args=["--cscore=0", "--tandem_Nmax=0"],
                          scan_args=["-n", "3"], tag="nofilter."))
    cases.append(run_case(bf, syn, d, "nostrip", "A", "B", bf_args=["--no_strip_names"], scan_args=["--dist=30"],
                          lift_args=["--no_strip_names"], tag="nostrip."))
    # --exclude with the default anchors
    cases.append(run_case(bf, syn, d, "exclude", "A", "B", bf_args=["--exclude=A.B.default.anchors"],
                          scan_args=["-n", "2", "--dist=8"], lift_args=["--dist=5"], tag="exclude."))
    # tie-heavy liftover: many tiny adjacent blocks, wide rescue radius
    cases.append(run_case(bf, syn, d, "ties", "A", "B", bf_args=["--cscore=0.3"], scan_args=["-n", "2", "--dist=2"],
                          lift_args=["--dist=12"], tag="ties."))
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    pack(d, os.path.join(HERE, "syn_pair.tar.gz"))

    # ---- self comparison -----------------------------------------------------------------
    d = tempfile.mkdtemp()
    synth.write_dataset(d, "A", "A", seed=12, Gq=2400, Gs=2400, n_lines=16000, Kq=3, Ks=3, self_mode=True)
    cases = [run_case(bf, syn, d, "self", "A", "A", bf_args=["--cscore=0.5"], scan_args=["--intrabound=100"])]
    cases.append(run_case(bf, syn, d, "self2", "A", "A", bf_args=["--cscore=0.2", "--tandem_Nmax=5"],
                          scan_args=["-n", "2", "--dist=4"], lift_args=["--dist=9"], tag="t2."))
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    pack(d, os.path.join(HERE, "syn_self.tar.gz"))

    # ---- irregular lines -----------------------------------------------------------------
    d = tempfile.mkdtemp()
    gq, gs, text = synth.make_pair(13, 600, 600, 6000, Kq=2, Ks=2, score_style=1)
    import numpy as np
    rng = np.random.default_rng(13)
    gq.write_bed(os.path.join(d, "A.bed"), rng); gs.write_bed(os.path.join(d, "B.bed"), rng)
    lines = bytes(text).decode().split("\n")
    out = []
    for i, ln in enumerate(lines):
        if not ln:
            continue
        r = i % 97
        if r == 3:
            out.append("")                                   # blank line
        if r == 5:
            out.append("# comment in the middle")
        if r == 7:
            ln = ln + "\textra\tcolumns"                     # >12 columns
        if r == 11:
            ln = "\t".join(ln.split("\t")[:8])               # short line: score/evalue stay 0
        if r == 13:
            ln = ln.replace("\t", " ")                       # spaces as separators
        if r == 17:
            f = ln.split("\t"); f[11] = "%.3g" % float(f[11]); f[10] = "%.2g" % float(f[10]); ln = "\t".join(f)
        if r == 19:
            f = ln.split("\t"); f[11] = "%.2e" % (float(f[11]) * 37); ln = "\t".join(f)   # exponent score
        if r == 23:
            ln = ln + "\r"                                   # CRLF
        if r == 29:
            f = ln.split("\t"); f[0] = f[0] + "|alt"; ln = "\t".join(f)   # pipe-suffixed name
        if r == 31:
            f = ln.split("\t"); f[5] = "x" + f[5]; ln = "\t".join(f)      # garbage int: rest of fields unparsed
        if r == 37:
            ln = "  " + ln                                   # leading whitespace
        if r == 41:
            f = ln.split("\t"); f[11] = f[11] + "abc"; ln = "\t".join(f)  # trailing junk after score
        out.append(ln)
    open(os.path.join(d, "A.B.last"), "w", newline="").write("\n".join(out))   # no trailing newline
    cases = [run_case(bf, syn, d, "edge", "A", "B", bf_args=["--cscore=0.4"], scan_args=["-n", "3"])]
    json.dump(cases, open(os.path.join(d, "cases.json"), "w"), indent=1)
    pack(d, os.path.join(HERE, "edge_lines.tar.gz"))


if __name__ == "__main__":
    main()

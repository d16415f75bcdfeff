#!/usr/bin/env python
"""Build the UNMODIFIED reference (tanghaibao/jcvi) into baseline/_ref/ so that bench.py can time its own CPU
implementation of the path -- `jcvi.compara.blastfilter.main`, `jcvi.compara.synteny.scan --liftover` with the
compiled Cython `cblast` -- next to the GPU path.

baseline/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored: it travels to the GPU box
like oracle/_ref/.  Recipe: copy /root/reference/src/jcvi, add jcvi/_version.py (src/jcvi/__init__.py:15-22 needs it),
compile jcvi.formats.cblast exactly like the reference's setup.py:16-20 (Extension + -O3), add shims for the three
absent pure-Python dependencies (natsort, more_itertools, Bio -- SURVEY.md 8c).  Gate: `scan` and `liftover` must
reproduce the reference's own golden files byte for byte.  The work is done by tests/golden/refenv.py (the same
recipe the golden vectors were generated with), pointed at baseline/_ref/.
"""
import filecmp
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "baseline", "_ref")


def build(force=False):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import refenv
    if not refenv.available():
        return None
    if force and os.path.isdir(OUT):
        shutil.rmtree(OUT)
    refenv.SCRATCH = OUT
    paths = refenv.setup()
    gate()
    return paths


def paths():
    """sys.path entries of a built baseline/_ref (shims first), or None"""
    if os.path.exists(os.path.join(OUT, ".built")):
        return [os.path.join(OUT, "shims"), os.path.join(OUT, "src")]
    return None


def gate():
    """the reference's own golden vectors (tests/compara/synteny.py/tests.yml) through the built copy"""
    import subprocess
    ref = os.environ.get("JCVI_REFERENCE", "/root/reference")
    t = os.path.join(ref, "tests", "compara", "synteny.py")
    if not os.path.isdir(t):
        return
    d = tempfile.mkdtemp(prefix="jcvi_ref_gate_")
    for f in os.listdir(os.path.join(t, "inputs")):
        shutil.copy(os.path.join(t, "inputs", f), d)
    code = (
        "import sys, logging; sys.path[:0] = %r\n"
        "import jcvi.compara.synteny as syn\n"
        "logging.getLogger('jcvi').setLevel(logging.ERROR)\n"
        "syn.scan(['testtigs.testchr.last', 'testtigs.testchr.anchors', '--qbed=testtigs.bed', '--sbed=testchr.bed'])\n"
        "syn.liftover(['test_empty_blocks.last', 'test_empty_blocks.anchors', '--qbed=test_empty_blocks.a.bed', "
        "'--sbed=test_empty_blocks.b.bed'])\n" % (paths(),))
    subprocess.run([sys.executable, "-c", code], cwd=d, check=True, capture_output=True)
    for got, want in (("testtigs.testchr.anchors", "testtigs.testchr.anchors"),
                      ("test_empty_blocks.lifted.anchors", "test_empty_blocks.lifted.anchors")):
        if not filecmp.cmp(os.path.join(d, got), os.path.join(t, "references", want), shallow=False):
            raise RuntimeError("baseline/_ref does not reproduce the reference's golden file %s" % want)
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

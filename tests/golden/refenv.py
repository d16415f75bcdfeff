"""Scratch build of the *reference* (tanghaibao/jcvi) used ONLY to generate and
cross-check golden vectors in this container.

Nothing here ships: the reference sources are copied from /root/reference into a
temporary directory outside the repository, its Cython `cblast` module is compiled
there with the reference's own recipe (setup.py:16-20, `-O3`), and three tiny shims
stand in for pure-Python dependencies that are not installed in this image
(natsort, more_itertools, Bio -- see SURVEY.md section 8c).

`/root/reference` does not exist on the GPU box, so only `make_golden.py` and the
optional CPU-side cross-check tests (skipped when the reference is absent) use this.
"""
import os
import shutil
import subprocess
import sys
import textwrap

REFERENCE = os.environ.get("JCVI_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("JCVI_REF_SCRATCH", "/tmp/jcvi_ref_scratch")

_SHIMS = {
    "natsort/__init__.py": '''
        """Minimal stand-in for natsort's default algorithm (ns.DEFAULT):
        split on runs of decimal digits, ints for digit runs, and a leading
        empty string when the first chunk is numeric."""
        import re
        _num = re.compile(r"([0-9]+)")

        def natsort_key(s):
            if not isinstance(s, str):
                return ("", s) if isinstance(s, (int, float)) else tuple(natsort_key(x) for x in s)
            parts = [p for p in _num.split(s) if p != ""]
            out = []
            for p in parts:
                out.append(int(p) if p.isdigit() and p.isascii() else p)
            if out and isinstance(out[0], int):
                out.insert(0, "")
            return tuple(out)

        def natsorted(seq, key=None, reverse=False):
            if key is None:
                return sorted(seq, key=natsort_key, reverse=reverse)
            return sorted(seq, key=lambda x: natsort_key(key(x)), reverse=reverse)

        natsort_keygen = lambda **kw: natsort_key
    ''',
    "more_itertools/__init__.py": '''
        from itertools import pairwise, zip_longest

        def grouper(iterable, n, incomplete="fill", fillvalue=None):
            args = [iter(iterable)] * n
            return zip_longest(*args, fillvalue=fillvalue)

        def flatten(listOfLists):
            from itertools import chain
            return chain.from_iterable(listOfLists)
    ''',
    "Bio/__init__.py": "",
    "Bio/SeqIO/__init__.py": '''
        def parse(*a, **k):
            raise NotImplementedError("Bio stub")
        def write(*a, **k):
            raise NotImplementedError("Bio stub")
        def index(*a, **k):
            raise NotImplementedError("Bio stub")
        def to_dict(*a, **k):
            raise NotImplementedError("Bio stub")
    ''',
    "Bio/Seq.py": "class Seq(str):\n    pass\n",
    "Bio/SeqRecord.py": "class SeqRecord(object):\n    pass\n",
    "Bio/SeqUtils/__init__.py": "",
    "Bio/SeqUtils/CheckSum.py": "def seguid(x):\n    return ''\n",
    "Bio/SeqFeature.py": "class SeqFeature(object):\n    pass\nclass FeatureLocation(object):\n    pass\n",
}


def available():
    return os.path.isdir(os.path.join(REFERENCE, "src", "jcvi"))


def setup(verbose=False):
    """Create (once) the scratch copy and return the list of sys.path entries."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE)
    src = os.path.join(SCRATCH, "src")
    shims = os.path.join(SCRATCH, "shims")
    stamp = os.path.join(SCRATCH, ".built")
    if not os.path.exists(stamp):
        if os.path.isdir(SCRATCH):
            shutil.rmtree(SCRATCH)
        os.makedirs(src)
        shutil.copytree(os.path.join(REFERENCE, "src", "jcvi"), os.path.join(src, "jcvi"))
        for root, dirs, files in os.walk(src):
            for d in dirs:
                os.chmod(os.path.join(root, d), 0o755)
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)
        with open(os.path.join(src, "jcvi", "_version.py"), "w") as fw:
            fw.write('__version__ = "0.0.0"\n')
        for rel, body in _SHIMS.items():
            p = os.path.join(shims, rel)
            os.makedirs(os.path.dirname(p), exist_ok=True)
            with open(p, "w") as fw:
                fw.write(textwrap.dedent(body))
        # build cblast exactly like the reference's setup.py (Extension + -O3)
        setup_py = os.path.join(SCRATCH, "build_cblast.py")
        with open(setup_py, "w") as fw:
            fw.write(textwrap.dedent('''
                from setuptools import setup, Extension
                from Cython.Build import cythonize
                ext = Extension("jcvi.formats.cblast", ["src/jcvi/formats/cblast.pyx"],
                                extra_compile_args=["-O3"])
                setup(name="jcvi-ref-cblast", ext_modules=cythonize([ext], quiet=True),
                      package_dir={"": "src"}, script_args=["build_ext", "--inplace", "-q"])
            '''))
        out = subprocess.run([sys.executable, setup_py], cwd=SCRATCH,
                             capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("cblast build failed:\n" + out.stdout + out.stderr)
        open(stamp, "w").close()
    if verbose:
        print("reference scratch at", SCRATCH)
    return [shims, src]


def activate():
    """Put the scratch reference on sys.path and return the imported modules."""
    for p in reversed(setup()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import logging
    import jcvi.compara.blastfilter as bf
    import jcvi.compara.synteny as syn
    from jcvi.formats.blast import BlastLine
    assert BlastLine.__module__ == "jcvi.formats.cblast", BlastLine.__module__
    logging.getLogger("jcvi").setLevel(logging.ERROR)
    return bf, syn

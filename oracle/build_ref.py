"""Compile the reference's own native unit on the hot path -- src/jcvi/formats/cblast.pyx, the
Cython `BlastLine` (sscanf parse / sprintf output) -- from the sources where they lie under
/root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).  No reference source is
copied into the repository: Cython's generated C goes to a temporary directory.

Recipe = the reference's setup.py:16-20 (Extension("jcvi.formats.cblast", ["…/cblast.pyx"],
extra_compile_args=["-O3"])); the module is importable on its own as `cblast`
(tests/test_oracle_vs_cblast.py uses it to pin the oracle's parse/format against the reference's
compiled code, here and on the GPU box).
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JCVI_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(force=False):
    pyx = os.path.join(REF, "src", "jcvi", "formats", "cblast.pyx")
    if not os.path.exists(pyx):
        return None
    os.makedirs(OUT, exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "cblast" + suffix)
    if os.path.exists(so) and not force and os.path.getmtime(so) >= os.path.getmtime(pyx):
        return so
    tmp = tempfile.mkdtemp(prefix="jcvi_cblast_")
    try:
        c_file = os.path.join(tmp, "cblast.c")
        subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", "cblast", "-o", c_file, pyx])
        inc = sysconfig.get_paths()["include"]
        cc = os.environ.get("CC", "gcc")
        subprocess.check_call([cc, "-O3", "-shared", "-fPIC", "-I" + inc, c_file, "-o", so])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return so


if __name__ == "__main__":
    print(build(force=True))

"""The sliver of jcvi.apps.base the three actions need (src/jcvi/apps/base.py:45-58,133-157,
272-276,333-349,371-373): an argparse wrapper whose parse_args() returns parse_known_args()
(unknown options are tolerated), the shared option bundles, and the "jcvi" logger."""
import logging
import os
import sys
from argparse import ArgumentParser

logger = logging.getLogger("jcvi")
if not logger.handlers:
    _h = logging.StreamHandler(sys.stderr)
    _h.setFormatter(logging.Formatter("%(message)s"))
    logger.addHandler(_h)
    logger.setLevel(logging.DEBUG if os.environ.get("JCVI_B200_DEBUG") else logging.INFO)
    logger.propagate = False


class OptionParser(ArgumentParser):
    def __init__(self, doc):
        usage = doc.replace("%prog", "%(prog)s") if doc else None
        super().__init__(usage=usage)

    def parse_args(self, args=None):
        return self.parse_known_args(args)

    def set_beds(self):
        self.add_argument("--qbed", help="Path to qbed")
        self.add_argument("--sbed", help="Path to sbed")

    def set_stripnames(self, default=True):
        if default:
            self.add_argument("--no_strip_names", dest="strip_names", action="store_false", default=True,
                              help="do not strip alternative splicing (e.g. At5g06540.1 -> At5g06540)")
        else:
            self.add_argument("--strip_names", action="store_true", default=False,
                              help="strip alternative splicing (e.g. At5g06540.1 -> At5g06540)")

    def set_outfile(self, outfile="stdout"):
        self.add_argument("-o", "--outfile", default=outfile, help="Outfile name")


class no_gc(object):
    """The writers build hundreds of thousands of small row objects next to BEDs of 10^5 objects; every young-generation
    overflow would make the cyclic collector walk all of them (measured: 0.3 s of a 0.5 s C2 run).  Nothing here creates
    reference cycles, so collection is paused for the duration of an action."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.disable()

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()


def file_key(path):
    """Identity of a file's current contents for in-process caches: (absolute path, mtime in ns, size)."""
    import os
    st = os.stat(path)
    return (os.path.abspath(path), st.st_mtime_ns, st.st_size)


def read_text_file(path):
    """Whole file as a uint8 array (gz/bz2 transparently, like must_open formats/base.py:473-550)."""
    import numpy as np
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as fh:
            return np.frombuffer(fh.read(), dtype=np.uint8)
    if path.endswith(".bz2"):
        import bz2
        with bz2.open(path, "rb") as fh:
            return np.frombuffer(fh.read(), dtype=np.uint8)
    import os
    if os.path.getsize(path) >= (64 << 20):
        # a large table is mapped, not copied: the upload reads the page cache directly (read-only view)
        try:
            return np.memmap(path, dtype=np.uint8, mode="r")
        except (OSError, ValueError):
            pass
    return np.fromfile(path, dtype=np.uint8)

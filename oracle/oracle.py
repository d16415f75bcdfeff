"""ctypes binding of oracle/libjcvi_oracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleError(RuntimeError):
    pass


class ZeroAnchors(ValueError):
    """Mirrors `ValueError("A total of 0 anchor was found. Aborted.")` (synteny.py:1463-1465)."""


def build(force=False):
    so = os.path.join(_HERE, "libjcvi_oracle.so")
    src = os.path.join(_HERE, "jcvi_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libjcvi_oracle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libjcvi_oracle.so")
        if not os.path.exists(so) or (
            os.path.exists(os.path.join(_HERE, "jcvi_oracle.cpp"))
            and os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "jcvi_oracle.cpp"))
        ):
            build()
        L = ctypes.CDLL(so)
        L.orc_last_error.restype = ctypes.c_char_p
        _LIB = L
    return _LIB


def _b(s):
    return None if s is None else os.fsencode(s)


def _check(rc):
    if rc < 0:
        raise OracleError(_lib().orc_last_error().decode())
    return rc


def blastfilter(blast_file, qbed, sbed, out_path=None, strip_names=True, cscore=0.7,
                tandem_Nmax=10, exclude=None, qdups=None, sdups=None, qnewbed=None, snewbed=None):
    """blastfilter.py:35-161.  Returns stats dict; writes `<blast>.filtered` (or out_path)."""
    out_path = out_path or (blast_file + ".filtered")
    stats = (ctypes.c_long * 6)()
    rc = _lib().orc_blastfilter(_b(blast_file), _b(qbed), _b(sbed), int(bool(strip_names)),
                                ctypes.c_double(cscore or 0.0), int(tandem_Nmax or 0), _b(exclude),
                                _b(out_path), _b(qdups), _b(sdups), _b(qnewbed), _b(snewbed), stats)
    _check(rc)
    return dict(zip(("parsed", "joined", "after_exclude", "after_cscore", "after_tandem", "warnings"), stats))


def scan(blast_file, anchor_file, qbed, sbed, dist=20, min_size=4, intrabound=300):
    """synteny.py:1849-1904.  Raises ZeroAnchors like the reference's summary()."""
    stats = (ctypes.c_long * 3)()
    rc = _check(_lib().orc_scan(_b(blast_file), _b(qbed), _b(sbed), int(dist), int(min_size),
                                int(intrabound), _b(anchor_file), stats))
    out = dict(zip(("hits", "clusters", "anchors"), stats))
    if rc == 1:
        raise ZeroAnchors("A total of 0 anchor was found. Aborted.")
    return out


def liftover(blast_file, anchor_file, qbed, sbed, out_path=None, strip_names=True, dist=10):
    """synteny.py:1919-1971."""
    out_path = out_path or (anchor_file.rsplit(".", 1)[0] + ".lifted.anchors")
    stats = (ctypes.c_long * 5)()
    rc = _check(_lib().orc_liftover(_b(blast_file), _b(anchor_file), _b(qbed), _b(sbed),
                                    int(bool(strip_names)), int(dist), _b(out_path), stats))
    out = dict(zip(("hits", "anchors", "lifted", "removed", "corrected"), stats))
    if rc == 1:
        raise ZeroAnchors("A total of 0 anchor was found. Aborted.")
    return out


def cscore(blast_file, out_path, cutoff=0.9999, pct=False, writeblast=False, strip_names=True):
    """formats/blast.py:1098-1189 (`jcvi.formats.blast cscore`).  Writes out_path and, with writeblast,
    out_path + ".filtered.blast".  ZeroDivisionError like the reference when a line's names never scored."""
    rc = _lib().orc_cscore(_b(blast_file), _b(out_path), _b(out_path + ".filtered.blast") if writeblast else _b(None),
                           ctypes.c_double(cutoff), int(bool(pct)), int(bool(strip_names)))
    if rc == -3:
        raise ZeroDivisionError("float division by zero")
    _check(rc)


def blast_filter(blast_file, out_path, score=0, pctid=95.0, hitlen=100, evalue=0.01, noself=False, ids=None, inverse=False):
    """formats/blast.py:620-709 (`jcvi.formats.blast filter`); writes out_path."""
    _check(_lib().orc_blast_filter(_b(blast_file), _b(out_path), _b(ids), ctypes.c_double(score), ctypes.c_double(pctid),
                                   ctypes.c_long(int(hitlen)), ctypes.c_double(evalue), int(bool(noself)), int(bool(inverse))))


def mcscan(bedfile, anchorfile, out_path, iter=100, ascii=False, Nm=10, trackids=False, mergetandem=None):
    """compara/synteny.py:1538-1652 (`jcvi.compara.synteny mcscan`); writes out_path (+ out_path + ".tracks")."""
    rc = _lib().orc_mcscan(_b(bedfile), _b(anchorfile), _b(out_path), int(iter), int(bool(ascii)), int(Nm),
                           _b(out_path + ".tracks") if trackids else _b(None), _b(mergetandem))
    if rc == -4:
        raise KeyError(_lib().orc_last_error().decode())
    if rc == -5:
        raise ValueError(_lib().orc_last_error().decode())
    _check(rc)


def simple(anchorfile, qbed, sbed, out_path, rich=False, coords=False, noheader=False, bed=False):
    """compara/synteny.py:1137-1316 (`jcvi.compara.synteny simple`); writes out_path (+ out_path + ".bed" with bed=True).
    Returns False when the anchors file holds no block (the reference logs an error and returns)."""
    mode = (1 if rich else 0) | (2 if coords else 0) | (4 if noheader else 0) | (8 if bed else 0)
    rc = _lib().orc_simple(_b(anchorfile), _b(qbed), _b(sbed), _b(out_path), mode, _b(out_path + ".bed") if bed else _b(None))
    if rc == -4:
        raise KeyError(_lib().orc_last_error().decode())
    if rc == -5:
        raise ValueError(_lib().orc_last_error().decode())
    if rc == -7:
        raise AssertionError("block end genes on different seqids")
    _check(rc)
    return rc == 0


def synfind(blast_file, qbed, sbed, out_path, strip_names=True, window=40, cutoff=0.1):
    """compara/synfind.py:205-240 (`jcvi.compara.synfind`, --scoring=collinear, text output); writes out_path."""
    rc = _lib().orc_synfind(_b(blast_file), _b(qbed), _b(sbed), _b(out_path), int(bool(strip_names)), int(window),
                            ctypes.c_double(cutoff))
    if rc == -5:
        raise ValueError(_lib().orc_last_error().decode())
    _check(rc)


def kdtree(anchors, queries, bound):
    """Restated cKDTree(anchors, leafsize=16).query(queries, p=1, distance_upper_bound=bound)."""
    a = np.ascontiguousarray(anchors, dtype=np.int64).reshape(-1, 2)
    q = np.ascontiguousarray(queries, dtype=np.int64).reshape(-1, 2)
    n, m = len(a), len(q)
    idx = np.zeros(max(n, 1), dtype=np.int32)
    near = np.zeros(max(m, 1), dtype=np.int32)
    dist = np.zeros(max(m, 1), dtype=np.float64)
    _lib().orc_kdtree(a.ctypes.data_as(ctypes.c_void_p), n, q.ctypes.data_as(ctypes.c_void_p), m,
                      ctypes.c_double(bound), idx.ctypes.data_as(ctypes.c_void_p),
                      near.ctypes.data_as(ctypes.c_void_p), dist.ctypes.data_as(ctypes.c_void_p))
    return idx[:n], near[:m], dist[:m]


def gene_name(s):
    buf = ctypes.create_string_buffer(1024)
    _check(_lib().orc_gene_name(s.encode(), buf, 1024))
    return buf.value.decode()


def natcmp(a, b):
    return _lib().orc_natcmp(a.encode(), b.encode())


def bed_order(bed_path, out_path):
    _check(_lib().orc_bed_order(_b(bed_path), _b(out_path)))


def blastline_str(line):
    buf = ctypes.create_string_buffer(2048)
    _check(_lib().orc_blastline_str(line.encode(), buf, 2048))
    return buf.value.decode()

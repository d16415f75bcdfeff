"""
%prog rice.sorghum.last --qbed=rice.bed --sbed=sorghum.bed

B200-native drop-in for `jcvi.compara.synfind` (src/jcvi/compara/synfind.py): for every gene the syntenic
regions of the window around it, from both sides.  Same argv, same output lines
(query, anchor, gray S/G/F, score, flank distance, orientation, qnote, snote), same sqlite table.
The per-gene work (windows, single linkage along the subject, flankers, LIS / LDS) runs in libjcvi_b200.so
(jx_synfind_text / jx_synfind_points, jcvi_b200/csrc/jx_synfind.cu); the host only prints the rows.
"""
import os.path as op
import sqlite3
import sys

import numpy as np

from ..apps_base import OptionParser, logger
from ..engine import get_engine
from .synteny import check_beds

_GRAY = ("S", "G", "F")


def _rows(res, lo, hi, qbed, sbed, qnote, snote):
    """synfind.py:164-198 for rows [lo, hi) of a device result: the printed columns."""
    if hi <= lo:
        return []
    qa = [b.accn for b in qbed]
    sa = [b.accn for b in sbed]
    seq = {}
    schr = np.fromiter((seq.setdefault(b.seqid, len(seq)) for b in sbed), dtype=np.int64, count=len(sbed))
    spos = np.fromiter((b.start for b in sbed), dtype=np.int64, count=len(sbed))
    q, a, le, ri = (x[lo:hi].astype(np.int64) for x in (res.query, res.syntelog, res.left, res.right))
    ld = np.where(schr[a] == schr[le], np.abs(spos[a] - spos[le]), 0)
    rd = np.where(schr[a] == schr[ri], np.abs(spos[a] - spos[ri]), 0)
    flank = (np.maximum(ld, rd) // 10000 + 1) * 10000
    sc, gr, mi = res.score[lo:hi].tolist(), res.gray[lo:hi].tolist(), res.minus[lo:hi].tolist()
    return [(qa[qq], sa[aa], _GRAY[g], s, f, "-" if m else "+", qnote, snote)
            for qq, aa, g, s, f, m in zip(q.tolist(), a.tolist(), gr, sc, flank.tolist(), mi)]


def _notes(qbed, sbed, opts):
    qnote, snote = opts.qnote, opts.snote
    if qnote == "null" or snote == "null":
        qnote = op.basename(qbed.filename).split(".")[0]
        snote = op.basename(sbed.filename).split(".")[0]
    return qnote, snote


def _emit(rows, fw, c):
    if fw:
        fw.write("".join("%s\t%s\t%s\t%d\t%d\t%s\t%s\t%s\n" % r for r in rows))
    else:
        c.executemany("insert into synteny values (?,?,?,?,?,?,?,?)", rows)


def _check_scoring(opts):
    if opts.scoring != "collinear":
        # the reference leaves `orientation` unassigned for --scoring=density (synfind.py:94-124)
        raise UnboundLocalError("cannot access local variable 'orientation' where it is not associated with a value")


def batch_query(qbed, sbed, all_data, opts, fw=None, c=None, transpose=False):
    """synfind.py:130-202 on distinct (qi, si) pairs."""
    _check_scoring(opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, qbed is sbed or qbed.filename == sbed.filename)
    cutoff = int(opts.cutoff * opts.window)
    window = opts.window // 2
    qnote, snote = _notes(qbed, sbed, opts)
    if transpose:
        if not all_data:
            raise ValueError("not enough values to unpack (expected 2, got 0)")     # transposed([]), synfind.py:40-42
    qi = np.fromiter((p[0] for p in all_data), dtype=np.int32, count=len(all_data))
    si = np.fromiter((p[1] for p in all_data), dtype=np.int32, count=len(all_data))
    res = eng.synfind_points(qi, si, window, cutoff, passes=2 if transpose else 1)
    n = res.n_forward + res.n_mirror
    if transpose:
        rows = _rows(res, 0, n, sbed, qbed, snote, qnote)
    else:
        rows = _rows(res, 0, n, qbed, sbed, qnote, snote)
    _emit(rows, fw, c)


def synfind(blastfile, p, opts):
    """synfind.py:205-240"""
    _check_scoring(opts)
    sqlite = opts.sqlite
    qbed, sbed, qorder, sorder, is_self = check_beds(blastfile, p, opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, is_self)
    res = eng.synfind_text(eng.upload_file(blastfile), strip_names=opts.strip_names, window=opts.window, cutoff=opts.cutoff)
    logger.debug("A total of %d BLAST imported from `%s`.", res.stats["hits"], blastfile)

    c = conn = None
    if sqlite:
        conn = sqlite3.connect(sqlite)
        c = conn.cursor()
        c.execute("drop table if exists synteny")
        c.execute("create table synteny (query text, anchor text, gray varchar(1), score integer, dr integer, "
                  "orientation varchar(1), qnote text, snote text)")
        fw = None
    else:
        fw = sys.stdout if opts.outfile in ("stdout", "-") else open(opts.outfile, "w")

    qnote, snote = _notes(qbed, sbed, opts)
    _emit(_rows(res, 0, res.n_forward, qbed, sbed, qnote, snote), fw, c)
    if qbed.filename == sbed.filename:
        logger.debug("Self comparisons, mirror ignored")
    else:
        if res.n_points == 0:
            raise ValueError("not enough values to unpack (expected 2, got 0)")     # transposed([]), synfind.py:40-42
        _emit(_rows(res, res.n_forward, res.n_forward + res.n_mirror, sbed, qbed, snote, qnote), fw, c)

    if sqlite:
        c.execute("create index q on synteny (query)")
        conn.commit()
        c.close()
    elif fw is not sys.stdout:
        fw.close()


def main(args):
    p = OptionParser(__doc__)
    p.set_beds()
    p.set_stripnames()
    p.set_outfile()

    coge_group = p.add_argument_group("CoGe-specific options")
    coge_group.add_argument("--sqlite", help="Write sqlite database")
    coge_group.add_argument("--qnote", default="null", help="Query dataset group id")
    coge_group.add_argument("--snote", default="null", help="Subject dataset group id")

    params_group = p.add_argument_group("Synteny parameters")
    params_group.add_argument("--window", type=int, default=40, help="Synteny window size")
    params_group.add_argument("--cutoff", type=float, default=0.1, help="Minimum number of anchors to call synteny")
    params_group.add_argument("--scoring", choices=("collinear", "density"), default="collinear", help="Scoring scheme")

    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    (blastfile,) = args
    synfind(blastfile, p, opts)


if __name__ == "__main__":
    main(sys.argv[1:])

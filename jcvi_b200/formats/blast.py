"""Drop-in for `jcvi.formats.blast cscore` (src/jcvi/formats/blast.py:1098-1189), the reciprocal-best-hit
C-score filter that runs on a raw BLAST/LAST table without BED files.

`cscore()` is one C-ABI call (`jx_blast_cscore`): the two passes over the table run on the GPU
(`jx_distinct_names` finds one representative line per distinct name, the names become the name table of
both sides, `jx_cscore_begin` / `jx_cscore_finish` are the path's own tokenizer + per-name best score +
float64 ratio predicate); the library's host side does what is O(distinct names) + O(surviving lines):
the `pairs` dictionary, `sorted()` and the `{:.2f}` writer, on values parsed with the same glibc calls
cblast makes.  `cscore_pairs()` is the same thing step by step in Python, for library use."""
import sys

import numpy as np

from ..apps_base import OptionParser, logger, read_text_file
from ..engine import get_engine


def gene_name(st, exclude=(b"ev",), sep=b"."):
    """cbook.py:308-329, on bytes."""
    if any(st.startswith(x) for x in exclude):
        sep = None
    st = st.split(b"|")[0]
    if sep and sep in st:
        name, suffix = st.rsplit(sep, 1)
    else:
        name, suffix = st, b""
    if len(suffix) != 1:
        name = st
    return name


def _open_out(path):
    if path == "stdout":
        return sys.stdout, False
    if path == "stderr":
        return sys.stderr, False
    return open(path, "w"), True


def cscore(args):
    """
    %prog cscore blastfile > cscoreOut

    See supplementary info for sea anemone genome paper, C-score formula:

        cscore(A,B) = score(A,B) /
             max(best score for A, best score for B)

    A C-score of one is the same as reciprocal best hit (RBH).

    Output file will be 3-column (query, subject, cscore). Use --cutoff to
    select a different cutoff.
    """
    p = OptionParser(cscore.__doc__)
    p.add_argument("--cutoff", default=0.9999, type=float, help="Minimum C-score to report")
    p.add_argument("--pct", default=False, action="store_true", help="Also include pct as last column")
    p.add_argument("--writeblast", default=False, action="store_true", help="Also write filtered blast file")
    p.set_stripnames()
    p.set_outfile()
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    (blastfile,) = args
    logger.debug("Register best scores ..")
    out, blast_out, _ = get_engine().blast_cscore(read_text_file(blastfile), cutoff=opts.cutoff, strip_names=opts.strip_names,
                                                  pct=opts.pct, writeblast=opts.writeblast)
    fw, close_fw = _open_out(opts.outfile)
    fw.write(out.decode())
    if close_fw:
        fw.close()
    else:
        fw.flush()
    if opts.writeblast:
        with open(opts.outfile + ".filtered.blast", "w") as fwb:
            fwb.write(blast_out.decode())


def filtered_blastfile_name(blastfile, pctid, hitlen, inverse=False):
    """blast.py:604-617."""
    pctid_str = "{:.1f}".format(pctid).replace(".", "_").replace("_0", "")
    newblastfile = blastfile + ".P{0}L{1}".format(pctid_str, hitlen)
    if inverse:
        newblastfile += ".inverse"
    return newblastfile


def filter(args):
    """
    %prog filter test.blast

    Produce a new blast file and filter based on:
    - score: >= cutoff
    - pctid: >= cutoff
    - hitlen: >= cutoff
    - evalue: <= cutoff
    - ids: valid ids

    Use --inverse to obtain the complementary records for the criteria above.

    - noself: remove self-self hits
    """
    p = OptionParser(filter.__doc__)
    p.add_argument("--score", dest="score", default=0, type=int, help="Score cutoff")
    p.add_argument("--pctid", default=95, type=float, help="Sequence percent identity")
    p.add_argument("--hitlen", default=100, type=int, help="Minimum overlap length")
    p.add_argument("--evalue", default=0.01, type=float, help="E-value cutoff")
    p.add_argument("--noself", default=False, action="store_true", help="Remove self-self hits")
    p.add_argument("--ids", help="Path to file with ids to retain")
    p.add_argument("--inverse", default=False, action="store_true", help="Similar to grep -v, inverse")
    p.set_outfile(outfile=None)
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    ids = None
    if opts.ids:
        ids = set()
        for row in read_text_file(opts.ids).tobytes().splitlines():
            if row[:1] == b"#":
                continue
            ids.update(row.replace(b",", b"\t").split())
    (blastfile,) = args
    newblastfile = filtered_blastfile_name(opts.outfile or blastfile, opts.pctid, opts.hitlen, opts.inverse)
    text = read_text_file(blastfile)
    with open(newblastfile, "wb") as fw:
        get_engine().blast_filter(text, score=opts.score, pctid=opts.pctid, hitlen=opts.hitlen, evalue=opts.evalue,
                                  noself=opts.noself, inverse=opts.inverse, ids=ids, sink=fw)
    return newblastfile


def cscore_pairs(text, cutoff, strip_names=True, engine=None):
    """-> {(query, subject): (cscore, pctid, str(BlastLine))} exactly as the reference's `pairs` dict
    (blast.py:1158-1172); names are bytes."""
    eng = engine or get_engine()
    strip = bool(strip_names)
    dev = eng.upload(np.ascontiguousarray(text, dtype=np.uint8)) if not hasattr(text, "ptr") else text
    try:
        for seed in range(4):
            logger.debug("Register best scores ..")
            lines, sides, st = eng.distinct_names(dev, strip_names=strip, seed=seed)
            if st["short"]:
                # a BlastLine with an empty query or subject: best_score[""] stays 0.0 in the reference
                # and its second pass divides by it (or silently keys the dict with "")
                raise NotImplementedError("%d line(s) with fewer than two fields in the BLAST table" % st["short"])
            names = set()
            for ln, side in zip(lines, sides.tolist()):
                f = ln.split(None, 2)
                nm = f[side]
                names.add(gene_name(nm) if strip else nm)
            names = sorted(names)
            index = {nm: i for i, nm in enumerate(names)}
            eng.load_names(names)
            eng.cscore_begin(dev, strip_names=strip)
            if eng.session_counts()["unmapped"] == 0:
                break
            logger.debug("two names share a 64-bit hash (seed %d); retrying", seed)
        else:
            raise RuntimeError("could not separate the names of the table")
        best = eng.best_table(len(names))
        kept, n_kept = eng.cscore_finish(float(cutoff))
    finally:
        if not hasattr(text, "ptr"):
            eng.release_text()
    pairs = {}
    if n_kept:
        for row in bytes(memoryview(kept)).split(b"\n")[:n_kept]:
            nconv, query, subject, pctid, score, as_str = eng.host_blastline(row)
            if strip:
                query, subject = gene_name(query), gene_name(subject)
            s = float(score) / max(float(best[index[query]]), float(best[index[subject]]))
            if s > cutoff:
                pair = (query, subject)
                if pair not in pairs or s > pairs[pair][0]:
                    pairs[pair] = (s, float(pctid), as_str)
    return pairs


def main():
    actions = {"cscore": cscore, "filter": filter}
    if len(sys.argv) < 2 or sys.argv[1] not in actions:
        print("usage: python -m jcvi_b200.formats.blast cscore|filter blastfile [options]", file=sys.stderr)
        sys.exit(1)
    actions[sys.argv[1]](sys.argv[2:])


if __name__ == "__main__":
    main()

"""
%prog blast_file --qbed query.bed --sbed subject.bed

B200-native drop-in for `jcvi.compara.blastfilter` (src/jcvi/compara/blastfilter.py): same
argv, same `<blast>.filtered` bytes (and `.localdups` / `.nolocaldups.bed` with
--tandems_only).  The per-hit work runs in libjcvi_b200.so (jx_blastfilter); this module only
parses options, loads the BEDs and writes files.
"""
import os.path as op
import sys

import numpy as np

from ..apps_base import OptionParser, logger, file_key, no_gc
from ..engine import get_engine


def _exclude_keys(exclude, qorder, sorder):
    """filter_exclude (blastfilter.py:205-222): name pairs, either orientation -> rank keys."""
    from .base import AnchorFile
    keys = []
    ac = AnchorFile(exclude)
    for a, b, _ in ac.iter_pairs():
        for x, y in ((a, b), (b, a)):
            if x in qorder and y in sorder:
                keys.append((qorder[x][0] << 32) | sorder[y][0])
    return np.array(keys, dtype=np.uint64)


def tandem_families(mother, bed):
    """Families of local duplicates from the device's mother map, each ordered like the reference
    orders them (longest gene first, then accn: write_localdups, blastfilter.py:164-175); the
    first member is the family's mother."""
    members = {}
    for gene, m in enumerate(mother):
        if m != gene:
            members.setdefault(int(m), {int(m)}).add(gene)
    fams = []
    for ranks in members.values():
        feats = sorted((bed[i] for i in ranks), key=lambda f: (-abs(f.end - f.start), f.accn))
        fams.append([f.accn for f in feats])
    fams.sort()
    return fams


def write_localdups(mother, bed, dups_fh=None):
    """-> {duplicate accn: mother accn}; one tab-separated family per line into dups_fh."""
    fams = tandem_families(mother, bed)
    if dups_fh and fams:
        logger.debug("write local dups to file {}".format(dups_fh.name))
        dups_fh.write("".join("\t".join(f) + "\n" for f in fams))
    return {dup: f[0] for f in fams for dup in f[1:]}


def write_new_bed(bed, children):
    """<bed>.nolocaldups.bed: the annotation without the collapsed duplicates (blastfilter.py:188-197)."""
    stem, ext = op.splitext(bed.filename)
    out_name = stem + ".nolocaldups" + ext
    logger.debug("write tandem-filtered bed file %s" % out_name)
    with open(out_name, "w") as fh:
        fh.write("".join(str(row) + "\n" for row in bed if row.accn not in children))


def blastfilter_main(blast_file, p, opts):
    from .synteny import check_beds

    qbed, sbed, qorder, sorder, is_self = check_beds(blast_file, p, opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, is_self)
    excl = _exclude_keys(opts.exclude, qorder, sorder) if opts.exclude else None
    if excl is not None:
        logger.debug("running excluded pairs (--exclude `{}`) ..".format(opts.exclude))
    # one upload: the table stays resident, and `synteny liftover` of the same (unchanged) file in this process --
    # catalog.ortholog's next step -- neither reads nor copies it again
    key = file_key(blast_file)
    dev = eng.upload_file(blast_file)
    res = eng.blastfilter(dev, strip_names=opts.strip_names, cscore=opts.cscore, tandem_Nmax=opts.tandem_Nmax,
                          exclude_keys=excl, want_text=True, want_mothers=bool(opts.tandems_only))
    eng.remember_text(key, dev)
    st = res.stats
    logger.debug("Load BLAST file `{}` (total {} lines)".format(blast_file, st["lines"]))
    if st["unmapped"]:
        logger.warning("{} hits with names not in the BED files were skipped".format(st["unmapped"]))
    logger.debug("after cscore/dedupe/tandem filters: {} -> {} -> {} -> {}".format(
        st["mapped"], st["after_cscore"], st["after_dedupe"], st["after_tandem"]))
    if opts.tandem_Nmax and opts.tandems_only:
        with open(op.splitext(opts.qbed)[0] + ".localdups", "w") as fh:
            qd = write_localdups(res.q_mother, qbed, fh)
        write_new_bed(qbed, qd)
        if not is_self:
            with open(op.splitext(opts.sbed)[0] + ".localdups", "w") as fh:
                sd = write_localdups(res.s_mother, sbed, fh)
            write_new_bed(sbed, sd)
    with open(blast_file + ".filtered", "wb") as fw:
        fw.write(memoryview(res.text))
    return res


_OPTIONS = (
    (("--tandems_only",), dict(dest="tandems_only", action="store_true", default=False,
                               help="also write <bed>.localdups and <bed>.nolocaldups.bed")),
    (("--tandem_Nmax",), dict(type=int, default=10, help="collapse tandem copies at most this many genes apart")),
    (("--cscore",), dict(type=float, default=0.7,
                         help="keep a hit when score / max(best score of either gene) exceeds this; 0 disables")),
    (("--exclude",), dict(help="anchors file whose pairs are dropped (either orientation)")),
)


def main(args):
    p = OptionParser(__doc__)
    p.set_beds()
    p.set_stripnames()
    for flags, kw in _OPTIONS:
        p.add_argument(*flags, **kw)
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    with no_gc():
        blastfilter_main(args[0], p, opts)


if __name__ == "__main__":
    main(sys.argv[1:])

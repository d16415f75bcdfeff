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

from ..apps_base import OptionParser, logger, read_text_file
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


def write_localdups(mother, bed, dups_fh=None):
    """write_localdups (blastfilter.py:164-185) from the device's mother map."""
    groups = {}
    for g, m in enumerate(mother):
        if m != g:
            groups.setdefault(int(m), [int(m)]).append(g)
    tandem_groups = []
    for m, members in groups.items():
        rows = [bed[i] for i in members]
        rows.sort(key=lambda a: (-abs(a.end - a.start), a.accn))
        tandem_groups.append([x.accn for x in rows])
    dups_to_mother = {}
    n = 1
    for accns in sorted(tandem_groups):
        if dups_fh:
            print("\t".join(accns), file=dups_fh)
            if n:
                n -= 1
                logger.debug("write local dups to file {}".format(dups_fh.name))
        for dup in accns[1:]:
            dups_to_mother[dup] = accns[0]
    return dups_to_mother


def write_new_bed(bed, children):
    out_name = "%s.nolocaldups%s" % op.splitext(bed.filename)
    logger.debug("write tandem-filtered bed file %s" % out_name)
    with open(out_name, "w") as fh:
        for row in bed:
            if row["accn"] in children:
                continue
            print(row, file=fh)


def blastfilter_main(blast_file, p, opts):
    from .synteny import check_beds

    qbed, sbed, qorder, sorder, is_self = check_beds(blast_file, p, opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, is_self)
    text = read_text_file(blast_file)
    excl = _exclude_keys(opts.exclude, qorder, sorder) if opts.exclude else None
    if excl is not None:
        logger.debug("running excluded pairs (--exclude `{}`) ..".format(opts.exclude))
    res = eng.blastfilter(text, strip_names=opts.strip_names, cscore=opts.cscore, tandem_Nmax=opts.tandem_Nmax,
                          exclude_keys=excl, want_text=True, want_mothers=bool(opts.tandems_only))
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


def main(args):
    p = OptionParser(__doc__)
    p.set_beds()
    p.set_stripnames()
    p.add_argument("--tandems_only", dest="tandems_only", action="store_true", default=False,
                   help="only calculate tandems, write .localdup file and exit.")
    p.add_argument("--tandem_Nmax", type=int, default=10, help="merge tandem genes within distance")
    p.add_argument("--cscore", type=float, default=0.7,
                   help="retain hits that have good bitscore. a value of 0.5 means keep all values that are 50%% "
                        "or greater of the best hit. higher is more stringent")
    p.add_argument("--exclude", help="Remove anchors from a previous run")
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    (blastfile,) = args
    blastfilter_main(blastfile, p, opts)


if __name__ == "__main__":
    main(sys.argv[1:])

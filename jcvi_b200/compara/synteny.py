"""B200-native drop-in for the hot subset of `jcvi.compara.synteny`
(src/jcvi/compara/synteny.py): `scan`, `liftover`, `batch_scan`, `synteny_scan`,
`synteny_liftover`, `check_beds`, `get_bed_filenames`, `summary` -- same argv, same return
values, same `.anchors` / `.lifted.anchors` bytes, same exceptions.  Per-hit work runs in
libjcvi_b200.so (jx_scan_text, jx_liftover_text, jx_scan_points, jx_nearest_anchor).
"""
import os.path as op
import sys
from collections.abc import Iterable

import numpy as np

from ..apps_base import OptionParser, file_key, logger, no_gc, read_text_file
from ..engine import get_engine
from ..formats.bed import Bed
from .base import AnchorFile


def _score(cluster):
    """synteny.py:214-219"""
    x, y = list(zip(*cluster))[:2]
    return min(len(set(x)), len(set(y)))


def get_bed_filenames(hintfile, p, opts):
    """--qbed/--sbed, or <dir>/<X>.bed and <dir>/<Y>.bed inferred from a file called X.Y.*
    (synteny.py:476-488)."""
    if opts.qbed and opts.sbed:
        return opts.qbed, opts.sbed
    folder, base = op.split(hintfile)
    stems = base.split(".")
    if len(stems) < 2:
        print("Options --qbed and --sbed are required", file=sys.stderr)
        sys.exit(not p.print_help())
    opts.qbed, opts.sbed = (op.join(folder, x + ".bed") for x in stems[:2])
    logger.debug("Assuming --qbed={0} --sbed={1}".format(opts.qbed, opts.sbed))
    return opts.qbed, opts.sbed


_BED_CACHE = {}


def _load_bed(path):
    import os
    st = os.stat(path)
    key = (op.abspath(path), st.st_mtime_ns, st.st_size)
    b = _BED_CACHE.get(key)
    if b is None:
        if len(_BED_CACHE) > 8:
            _BED_CACHE.clear()
        b = _BED_CACHE[key] = Bed(path)
    return b


def check_beds(hintfile, p, opts, sorted=True):
    """synteny.py:491-503"""
    qbed_file, sbed_file = get_bed_filenames(hintfile, p, opts)
    is_self = qbed_file == sbed_file
    if is_self:
        logger.debug("Looks like self-self comparison.")
    qbed = _load_bed(opts.qbed)
    sbed = qbed if is_self else _load_bed(opts.sbed)
    return qbed, sbed, qbed.order, sbed.order, is_self


def add_arguments(p, args, dist=10):
    """synteny.py:506-523"""
    p.set_beds()
    p.add_argument("--dist", default=dist, type=int, help="Extent of flanking regions to search")
    opts, args = p.parse_args(args)
    if len(args) != 2:
        sys.exit(not p.print_help())
    blast_file, anchor_file = args
    return blast_file, anchor_file, opts.dist, opts


# ---- library API on points ------------------------------------------------------------------
def _points_to_arrays(points):
    qi = np.fromiter((p[0] for p in points), dtype=np.int64, count=len(points))
    si = np.fromiter((p[1] for p in points), dtype=np.int64, count=len(points))
    return qi, si


def synteny_scan(points, xdist, ydist, N, is_self=False, intrabound=300):
    """synteny.py:402-434: single linkage on one chromosome pair; returns clusters of the input
    tuples, each sorted, in the reference's (Grouper insertion) order."""
    if not points:
        return []
    qi, si = _points_to_arrays(points)
    sc = np.zeros(len(points), dtype=np.float32)
    res = get_engine().scan_points(qi, si, sc, None, xdist=xdist, ydist=ydist, N=N, is_self=is_self,
                                   intrabound=intrabound)
    bs = res.block_start
    return [sorted(points[j] for j in res.index[bs[b]:bs[b + 1]]) for b in range(res.n_blocks)]


def batch_scan(points, xdist=20, ydist=20, N=5, is_self=False, intrabound=300):
    """synteny.py:437-452: accepts hit objects (.qseqid/.sseqid/.qi/.si/.score) or 3-tuples."""
    if not points:
        return []
    if isinstance(points[0], Iterable) and len(points[0]) == 3:
        return synteny_scan(list(points), xdist, ydist, N, is_self=is_self, intrabound=intrabound)
    pairs = sorted(set((b.qseqid, b.sseqid) for b in points))
    gid = {cp: i for i, cp in enumerate(pairs)}
    tup = [(b.qi, b.si, b.score) for b in points]
    qi, si = _points_to_arrays(tup)
    sc = np.fromiter((t[2] for t in tup), dtype=np.float32, count=len(tup))
    grp = np.fromiter((gid[(b.qseqid, b.sseqid)] for b in points), dtype=np.int64, count=len(points))
    res = get_engine().scan_points(qi, si, sc, grp, xdist=xdist, ydist=ydist, N=N, is_self=is_self,
                                   intrabound=intrabound)
    bs = res.block_start
    return [sorted(tup[j] for j in res.index[bs[b]:bs[b + 1]]) for b in range(res.n_blocks)]


def synteny_liftover(points, anchors, dist):
    """synteny.py:455-473: yields (point, nearest anchor) for points with 0 < L1 distance < dist,
    the nearest anchor being cKDTree's choice among equidistant ones."""
    points = np.array(points, dtype=int)
    ppoints = points[:, :2] if points.shape[1] > 2 else points
    anchors = np.asarray(anchors)
    near, dd = get_engine().nearest_anchor(ppoints, anchors, dist)
    n = len(anchors)
    for point, d, idx in zip(points, dd, near):
        if idx == n:
            continue
        if d == 0:
            continue
        yield point, tuple(anchors[idx])


# ---- actions -------------------------------------------------------------------------------------
def _describe(values):
    v = np.array(values)
    return "Min={} Max={} N={} Mean={:.2f} SD={:.2f} Median={} Sum={}".format(
        v.min(), v.max(), v.size, np.mean(v), np.std(v), np.median(v), v.sum())


def _summary_blocks(blocks, prefix=None):
    if not blocks or blocks == [[]]:
        logger.debug("A total of 0 anchor was found. Aborted.")
        raise ValueError("A total of 0 anchor was found. Aborted.")
    sizes = [len(b) for b in blocks]
    nr_sizes = [_score(b) for b in blocks]          # non-redundant: min(#distinct a, #distinct b)
    err = sys.stderr
    print("A total of {0} (NR:{1}) anchors found in {2} clusters.".format(sum(sizes), sum(nr_sizes), len(blocks)), file=err)
    print("Stats:", _describe(sizes), file=err)
    print("NR stats:", _describe(nr_sizes), file=err)
    if prefix:
        width = len(str(len(blocks)))
        for i, n in enumerate(sizes, 1):
            print("{0}{1:0{2}d}\t{3}".format(prefix, i, width, n))


def summary(args):
    """Block statistics on stderr; ValueError when the file holds no anchor (synteny.py:1445-1484)."""
    p = OptionParser(summary.__doc__)
    p.add_argument("--prefix", help="Generate per block stats")
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    _summary_blocks(AnchorFile(args[0]).blocks, opts.prefix)


def _int_str(x):
    return str(int(x))


def write_anchors(res, qbed, sbed, anchor_file):
    """scan's writer (synteny.py:1897-1903).  Returns the blocks as AnchorFile would parse them back."""
    qa = qbed.tables()["accns"]
    sa = sbed.tables()["accns"]
    out = []
    blocks = []
    bs = res.block_start
    q, s, sc = res.q_rank.tolist(), res.s_rank.tolist(), res.score.astype(np.float64).tolist()
    for b in range(res.n_blocks):
        out.append("###\n")
        rows = [[qa[q[t]], sa[s[t]], _int_str(sc[t])] for t in range(bs[b], bs[b + 1])]
        out.extend("%s\t%s\t%s\n" % (r[0], r[1], r[2]) for r in rows)
        blocks.append(rows)
    with open(anchor_file, "w") as fw:
        fw.write("".join(out))
    return blocks


def scan(args):
    with no_gc():
        return _scan(args)


def _scan(args):
    """
    %prog scan blastfile anchor_file [options]

    pull out syntenic anchors from blastfile based on single-linkage algorithm
    """
    p = OptionParser(_scan.__doc__)
    p.add_argument("-n", "--min_size", dest="n", type=int, default=4, help="minimum number of anchors in a cluster")
    p.add_argument("--intrabound", default=300, type=int,
                   help="Lower bound of intra-chromosomal blocks (only for self comparison)")
    p.add_argument("--liftover", help="Scan BLAST file to find extra anchors")
    p.add_argument("--liftover_dist", type=int, help="Distance to extend from liftover. Defaults to half of --dist")
    p.set_stripnames()

    blast_file, anchor_file, dist, opts = add_arguments(p, args, dist=20)
    qbed, sbed, qorder, sorder, is_self = check_beds(blast_file, p, opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, is_self)
    logger.debug("Chaining distance = {0}".format(dist))
    res = eng.scan_text(read_text_file(blast_file), strip_names=False, xdist=dist, ydist=dist, N=opts.n,
                        intrabound=opts.intrabound)
    logger.debug("A total of %d BLAST imported from `%s`.", res.stats["hits"], blast_file)
    blocks = write_anchors(res, qbed, sbed, anchor_file)
    _summary_blocks(blocks)                        # of the file just written (the reference re-reads it)

    lo = opts.liftover
    if not lo:
        return anchor_file
    dargs = ["--qbed=" + opts.qbed, "--sbed=" + opts.sbed]
    if not opts.strip_names:
        dargs += ["--no_strip_names"]
    liftover_dist = opts.liftover_dist or dist // 2
    dargs += ["--dist={}".format(liftover_dist)]
    block_of = np.repeat(np.arange(res.n_blocks, dtype=np.int32), np.diff(res.block_start))
    return liftover([lo, anchor_file] + dargs, _blocks=blocks, _ranks=(res.q_rank, res.s_rank, block_of))


def _anchor_ranks(lines, qorder, sorder):
    """ranks of every anchor line (-1 when a name is not in its BED) + its block index"""
    a_qi = np.fromiter((qorder[r[0]][0] if (len(r) > 1 and r[0] in qorder and r[1] in sorder) else -1 for r, _ in lines),
                       dtype=np.int32, count=len(lines))
    a_si = np.fromiter((sorder[r[1]][0] if (len(r) > 1 and r[0] in qorder and r[1] in sorder) else -1 for r, _ in lines),
                       dtype=np.int32, count=len(lines))
    a_block = np.fromiter((bi for _, bi in lines), dtype=np.int32, count=len(lines))
    return a_qi, a_si, a_block


def liftover(args, _blocks=None, _ranks=None):
    with no_gc():
        return _liftover(args, _blocks, _ranks)


def _liftover(args, _blocks=None, _ranks=None):
    """
    %prog liftover blastfile anchorfile [options]

    Typical use for this program is given a list of anchors (syntennic
    genes), choose from the blastfile the pairs that are close to the anchors.
    """
    p = OptionParser(_liftover.__doc__)
    p.set_stripnames()
    p.set_outfile(outfile=None)

    blast_file, anchor_file, dist, opts = add_arguments(p, args)
    qbed, sbed, qorder, sorder, is_self = check_beds(blast_file, p, opts)
    eng = get_engine()
    eng.load_pair(qbed, sbed, is_self)

    # _blocks: scan() hands over the blocks of the anchors file it has just written
    ac = AnchorFile.from_blocks(anchor_file, _blocks) if _blocks is not None else AnchorFile(anchor_file)
    lines = [(row, bi) for bi, block in enumerate(ac.blocks) for row in block]
    if _ranks is not None and len(_ranks[0]) == len(lines):
        # scan() also hands over the ranks of the rows it wrote (the names came from these ranks)
        a_qi, a_si, a_block = (np.ascontiguousarray(x, dtype=np.int32) for x in _ranks)
    else:
        a_qi, a_si, a_block = _anchor_ranks(lines, qorder, sorder)
    logger.debug("A total of {0} anchors imported.".format(int((a_qi >= 0).sum())))

    # the raw table blastfilter uploaded in this process, if it is still resident and the file is unchanged
    key = file_key(blast_file)
    text = eng.recall_text(key)
    if text is None:
        text = eng.upload_file(blast_file)
        eng.remember_text(key, text)
    res = eng.liftover_text(text, a_qi, a_si, a_block, strip_names=opts.strip_names, dist=dist)
    logger.debug("A total of %d BLAST imported from `%s`.", res.stats["hits"], blast_file)

    qa = qbed.tables()["accns"]
    sa = sbed.tables()["accns"]
    logger.debug("%d new pairs found (dist=%d).", res.n_lifted, dist)
    newanchorfile = opts.outfile or anchor_file.rsplit(".", 1)[0] + ".lifted.anchors"
    # The lifted pairs join their blocks, then AnchorFile.filter_blocks (compara/base.py:52-83).  `accepted` maps a pair to
    # the raw score of the pair, so the dictionary is not needed: the device already looked every anchor line up
    # (res.accepted / res.raw_score, index-aligned with `lines`), and a lifted pair carries its own raw score + "L".
    lifted_sc = res.score.astype(np.int64).astype(str).tolist()          # str(int(score)), synteny.py:1961
    if res.stats["hits"]:
        raw = res.raw_score.astype(np.int64).astype(str).tolist()
        blocks = [[] for _ in ac.blocks]
        dropped = fixed = 0
        for (row, bi), ok, rs in zip(lines, res.accepted.tolist(), raw):
            if not ok:
                dropped += 1
                continue
            sc = row[2]
            if sc != rs and sc != rs + "L":
                sc, fixed = rs, fixed + 1
            blocks[bi].append((row[0], row[1], sc))
    else:
        blocks = [list(b) for b in ac.blocks]
        dropped = fixed = None
    for q, s, sc, b in zip(res.q_rank.tolist(), res.s_rank.tolist(), lifted_sc, res.block.tolist()):
        blocks[b].append((qa[q], sa[s], sc + "L"))
    if dropped is not None:
        kept = [b for b in blocks if b]
        logger.debug("Removed %d existing anchors", dropped)
        if len(kept) != len(blocks):
            logger.debug("Removed %d empty blocks", len(blocks) - len(kept))
        logger.debug("Corrected scores for %d anchors", fixed)
        blocks = kept
    ac.blocks = blocks
    ac.print_to_file(filename=newanchorfile)
    _summary_blocks(ac.blocks)                     # of the file just written
    return newanchorfile


scan.__doc__ = _scan.__doc__
liftover.__doc__ = _liftover.__doc__


def _best_pairs(qs, ss, ts):
    """compara/base.py:139-150 get_best_pair: per gene the partner with the largest score, the first one on ties."""
    pairs = {}
    for q, s, t in zip(qs, ss, ts):
        t = int(t[:-1]) if t[-1] == "L" else int(t)
        if q not in pairs or pairs[q][1] < t:
            pairs[q] = (s, t)
    return dict((q, s) for q, (s, t) in pairs.items())


def make_ranges(ac, order, clip=10):
    """AnchorFile.make_ranges / make_range (compara/base.py:29-50,152-164) -> (ranges as (start, end, score, block id)
    in the reference's order, block_pairs: one dict per block)."""
    ranges = []
    block_pairs = [dict() for _ in ac.blocks]

    def one(q, s, t, i):
        pairs = _best_pairs(q, s, t)
        block_pairs[i].update(pairs)
        ranks = sorted(order[x][0] for x in q)
        qmin, qmax = ranks[0], ranks[-1]
        if qmax - qmin >= 2 * clip:
            qmin += clip / 2
            qmax -= clip / 2
        ranges.append((qmin, qmax, len(pairs), i))

    for i, ib in enumerate(ac.blocks):
        q, s, t = zip(*ib)
        if q[0] not in order:
            q, s = s, q
        one(q, s, t, i)
        assert q[0] in order
        if s[0] not in order:
            continue
        one(s, q, t, i)                 # self comparison: the block seen from the other side
    return ranges, block_pairs


def mcscan(args):
    """
    %prog mcscan bedfile anchorfile [options]

    Stack synteny blocks on a reference bed, MCSCAN style. The first column in
    the output is the reference order, given in the bedfile. Then each column
    next to it are separate 'tracks'.

    If --mergetandem=tandem_file is specified, tandem_file should have each
    tandem cluster as one line, tab separated.
    """
    p = OptionParser(mcscan.__doc__)
    p.add_argument("--iter", default=100, type=int, help="Max number of chains to output")
    p.add_argument("--ascii", default=False, action="store_true", help="Output symbols rather than gene names")
    p.add_argument("--Nm", default=10, type=int, help="Clip block ends to allow slight overlaps")
    p.add_argument("--trackids", action="store_true", help="Track block IDs in separate file")
    p.add_argument("--mergetandem", default=None,
                   help="merge tandems genes in output acoording to PATH-TO-TANDEM_FILE, cannot be used with --ascii")
    p.set_outfile()
    opts, args = p.parse_args(args)
    if len(args) != 2:
        sys.exit(not p.print_help())
    bedfile, anchorfile = args
    ascii_ = opts.ascii
    ofile = opts.outfile
    bed = Bed(bedfile)
    order = bed.order
    tandems = {}
    if opts.mergetandem:
        assert not ascii_
        for row in open(opts.mergetandem):
            row = row.split()
            s = ";".join(row)
            for atom in row:
                tandems[atom] = s
    ac = AnchorFile(anchorfile)
    ranges, block_pairs = make_ranges(ac, order, clip=opts.Nm)
    # dense ids: reference-BED accns (rows of the table) and partner names
    gid = {}
    for b in bed:
        gid.setdefault(b.accn, len(gid))
    names, nid = [], {}
    pg, pp, pb = [], [], []
    for i, pairs in enumerate(block_pairs):
        for q, s in pairs.items():
            if q not in gid:            # a key of a block dictionary that no BED row carries: never looked up
                continue
            if s not in nid:
                nid[s] = len(names)
                names.append(s)
            pg.append(gid[q]); pp.append(nid[s]); pb.append(i)
    print("Chain started: {0} blocks".format(len(ranges)), file=sys.stderr)
    r = np.array(ranges, dtype=np.float64).reshape(-1, 4)
    tracks, scores, table = get_engine().mcscan(r[:, 0], r[:, 1], r[:, 2].astype(np.int64), r[:, 3].astype(np.int32), opts.iter,
                                                pg, pp, pb, len(gid), len(ac.blocks))
    left = len(ranges)
    rid = r[:, 3].astype(np.int64)
    for it, (track, score) in enumerate(zip(tracks, scores)):
        left -= int(np.isin(rid, list(set(track))).sum())
        rid = rid[~np.isin(rid, list(set(track)))]
        msg = "Chain {0}: score={1}".format(it, score)
        msg += " {0} blocks remained..".format(left) if left else " done!"
        print(msg, file=sys.stderr)
    if opts.trackids:
        with open(ofile + ".tracks", "w") as fwlog:
            for track in tracks:
                print(",".join(str(x) for x in sorted(set(track))), file=fwlog)
    # partner id -> text (-1 = "."), --ascii / --mergetandem applied per name once
    text = np.empty(len(names) + 1, dtype=object)
    for k, nm in enumerate(names):
        text[k] = "x" if ascii_ else (tandems.get(nm, nm) if tandems else nm)
    text[len(names)] = tandems.get(".", ".") if tandems else "."
    cells = text[table]                                   # -1 indexes the last entry
    sep = "" if ascii_ else "\t"
    close = ofile not in ("stdout", "-")
    fw = open(ofile, "w") if close else sys.stdout
    try:
        for b in bed:
            fw.write(b.accn + "\t" + sep.join(cells[gid[b.accn]]) + "\n")
    finally:
        if close:
            fw.close()
    logger.debug("MCscan blocks written to `{0}`.".format(ofile))
    if opts.trackids:
        logger.debug("Block IDs written to `{0}`.".format(ofile + ".tracks"))


def get_orientation(ia, ib):
    """synteny.py:222-236 (the reference's own numpy call; used here only for blocks whose covariance is exactly 0,
    where its sign is rounding noise that no restatement can reproduce)."""
    if len(ia) != len(ib) or len(ia) < 2:
        return "+"
    slope, _ = np.polyfit(ia, ib, 1)
    return "+" if slope >= 0 else "-"


def simple(args):
    """
    %prog simple anchorfile --qbed=qbedfile --sbed=sbedfile [options]

    Write the block ends for each block in the anchorfile.
    GeneA1    GeneA2    GeneB1    GeneB2   +/-      score

    Optional additional columns:
    orderA1   orderA2   orderB1   orderB2  sizeA    sizeB   size    block_id

    With base coordinates (--coords):
    block_id  seqidA    startA    endA     bpSpanA  GeneA1   GeneA2  geneSpanA
    block_id  seqidB    startB    endB     bpSpanB  GeneB1   GeneB2  geneSpanB
    """
    p = OptionParser(simple.__doc__)
    p.add_argument("--rich", default=False, action="store_true", help="Output additional columns")
    p.add_argument("--coords", default=False, action="store_true", help="Output columns with base coordinates")
    p.add_argument("--bed", default=False, action="store_true", help="Generate BED file for the blocks")
    p.add_argument("--noheader", default=False, action="store_true", help="Don't output header")
    p.set_beds()
    opts, args = p.parse_args(args)
    if len(args) != 1:
        sys.exit(not p.print_help())
    (anchorfile,) = args
    additional, coords, header, bed = opts.rich, opts.coords or opts.bed, not opts.noheader, opts.bed
    ac = AnchorFile(anchorfile)
    simplefile = anchorfile.rsplit(".", 1)[0] + ".simple"
    qbed, sbed, qorder, sorder, is_self = check_beds(anchorfile, p, opts)
    pf = "-".join(anchorfile.split(".", 2)[:2])
    if ac.is_empty:
        logger.error("No blocks found in `%s`. Aborting ..", anchorfile)
        return
    if coords:
        h = "Block|Chr|Start|End|Span|StartGene|EndGene|GeneSpan|Orientation"
    else:
        h = "StartGeneA|EndGeneA|StartGeneB|EndGeneB|Orientation|Score"
        if additional:
            h += "|StartOrderA|EndOrderA|StartOrderB|EndOrderB|SizeA|SizeB|Size|Block"
    blocks = ac.blocks
    blk, ia, ib = [], [], []
    for i, block in enumerate(blocks):
        a, b, scores = zip(*block)
        ia.extend(qorder[x][0] for x in a)          # KeyError like the reference for genes outside the BEDs
        ib.extend(sorder[x][0] for x in b)
        blk.extend([i] * len(block))
    st = get_engine().block_stats(blk, ia, ib, len(blocks))      # the per-block reductions, on the device
    n = st["size"].astype(object)
    cov = n * st["sum_qs"].astype(object) - st["sum_q"].astype(object) * st["sum_s"].astype(object)   # exact Python ints
    starts = np.concatenate(([0], np.cumsum(st["size"].astype(np.int64))))
    out = ["\t".join(h.split("|"))] if header else []
    bbed = []
    atotalbase = btotalbase = 0
    for i in range(len(blocks)):
        astarti, aendi, bstarti, bendi = int(st["qmin"][i]), int(st["qmax"][i]), int(st["smin"][i]), int(st["smax"][i])
        astart, aend, bstart, bend = qbed[astarti].accn, qbed[aendi].accn, sbed[bstarti].accn, sbed[bendi].accn
        sizeA, sizeB, size = int(st["distinct_q"][i]), int(st["distinct_s"][i]), int(st["size"][i])
        if size < 2:
            orientation = "+"
        elif size * (aendi - astarti + 1) * (bendi - bstarti + 1) >= (1 << 62):     # the device's 64-bit sums could wrap: exact on the host
            lo, hi = int(starts[i]), int(starts[i + 1])
            c = size * sum(x * y for x, y in zip(ia[lo:hi], ib[lo:hi])) - sum(ia[lo:hi]) * sum(ib[lo:hi])
            orientation = ("+" if c > 0 else "-") if c else get_orientation(ia[lo:hi], ib[lo:hi])
        elif cov[i] != 0:
            orientation = "+" if cov[i] > 0 else "-"
        else:                                       # exactly flat: the reference's answer is polyfit's rounding noise
            lo, hi = int(starts[i]), int(starts[i + 1])
            orientation = get_orientation(ia[lo:hi], ib[lo:hi])
        aspan = aendi - astarti + 1
        bspan = bendi - bstarti + 1
        score = str(int((aspan * bspan) ** 0.5))
        block_id = pf + "-block-{0}".format(i)
        if coords:
            (s, e) = qbed[astarti], qbed[aendi]
            assert s.seqid == e.seqid
            aseqid, astartbase, aendbase = s.seqid, min((s.start, s.end), (e.start, e.end))[0], max(s.end, e.end)
            (s, e) = sbed[bstarti], sbed[bendi]
            assert s.seqid == e.seqid
            bseqid, bstartbase, bendbase = s.seqid, min((s.start, s.end), (e.start, e.end))[0], max(s.end, e.end)
            abase = aendbase - astartbase + 1
            bbase = bendbase - bstartbase + 1
            atotalbase += abase
            btotalbase += bbase
            out.append("\t".join(str(x) for x in (block_id, aseqid, astartbase, aendbase, abase, astart, aend, aspan, "+")))
            out.append("\t".join(str(x) for x in (block_id, bseqid, bstartbase, bendbase, bbase, bstart, bend, bspan, orientation)))
            if bed:
                bbed.append((bseqid, bstartbase, bendbase, "{}:{}-{}".format(aseqid, astartbase, aendbase), size, orientation))
            continue
        row = [astart, aend, bstart, bend, score, orientation]
        if additional:
            row += [astarti, aendi, bstarti, bendi, sizeA, sizeB, size, block_id]
        out.append("\t".join(str(x) for x in row))
    with open(simplefile, "w") as fws:
        fws.write("\n".join(out) + "\n")
    logger.debug("A total of {0} blocks written to `{1}`.".format(len(blocks), simplefile))
    if coords:
        print("Total block span in {0}: {1}".format(qbed.filename, atotalbase), file=sys.stderr)
        print("Total block span in {0}: {1}".format(sbed.filename, btotalbase), file=sys.stderr)
        print("Ratio: {0:.1f}x".format(max(atotalbase, btotalbase) * 1.0 / min(atotalbase, btotalbase)), file=sys.stderr)
    if bed:
        from ..formats.bed import natsort_key
        bedfile = simplefile + ".bed"
        bbed.sort(key=lambda r: (natsort_key(r[0]), r[1], r[3]))       # Bed.print_to_file(sorted=True)
        with open(bedfile, "w") as fw:
            for seqid, start, end, accn, size, orientation in bbed:
                fw.write("\t".join(str(x) for x in (seqid, max(start, 1) - 1, end, accn, size, orientation)) + "\n")
        logger.debug("Bed file written to `{}`".format(bedfile))


def main():
    actions = {"scan": scan, "liftover": liftover, "summary": summary, "mcscan": mcscan, "simple": simple}
    if len(sys.argv) < 2 or sys.argv[1] not in actions:
        print("usage: python -m jcvi_b200.compara.synteny {scan,liftover,summary,mcscan,simple} ...", file=sys.stderr)
        sys.exit(1)
    actions[sys.argv[1]](sys.argv[2:])


if __name__ == "__main__":
    main()

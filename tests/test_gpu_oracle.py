"""GPU parity against the CPU oracle on seeded synthetic tables (sizes the oracle finishes in
seconds), plus API-level checks of the point interfaces."""
import os

import numpy as np
import pytest

from conftest import read

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from jcvi_b200.compara import blastfilter, synteny
    return blastfilter, synteny


def _pipeline(tmp, mods, seed, Gq, Gs, n, self_mode=False, bf_args=(), scan_args=(), lift_args=(), group_by_query=False, **kw):
    import oracle
    from jcvi_b200 import synth
    bf, syn = mods
    d = str(tmp)
    q, s = ("A", "A") if self_mode else ("A", "B")
    qbed, sbed, last = synth.write_dataset(d, q, s, seed, Gq, Gs, n, self_mode=self_mode, **kw)
    if self_mode:
        sbed = qbed
    if group_by_query:
        # the order lastal / BLAST write: all alignments of a query together (the tokenizer then merges a warp's
        # equal-query best-score reductions -- its MERGE_Q instantiation)
        rows = read(last).split(b"\n")
        body = [r for r in rows if r and not r.startswith(b"#")]
        body.sort(key=lambda r: r.split(b"\t", 1)[0])                  # stable: file order inside a query
        with open(last, "wb") as fw:
            fw.write(b"\n".join([r for r in rows if r.startswith(b"#")] + body) + b"\n")
    beds = ["--qbed=" + qbed, "--sbed=" + sbed]

    def opt(args, name, default, cast):
        for a in args:
            if a.startswith("--" + name + "="):
                return cast(a.split("=", 1)[1])
        return default

    ost = oracle.blastfilter(last, qbed, sbed, out_path=last + ".orc.filtered",
                             strip_names="--no_strip_names" not in bf_args, cscore=opt(bf_args, "cscore", 0.7, float),
                             tandem_Nmax=opt(bf_args, "tandem_Nmax", 10, int))
    bf.main([last] + beds + list(bf_args))
    assert read(last + ".filtered") == read(last + ".orc.filtered")
    oracle.scan(last + ".filtered", last + ".orc.anchors", qbed, sbed, dist=opt(scan_args, "dist", 20, int),
                min_size=opt(scan_args, "min_size", 4, int), intrabound=opt(scan_args, "intrabound", 300, int))
    syn.scan([last + ".filtered", last + ".gpu.anchors"] + beds + list(scan_args))
    assert read(last + ".gpu.anchors") == read(last + ".orc.anchors")
    oracle.liftover(last, last + ".orc.anchors", qbed, sbed, out_path=last + ".orc.lifted",
                    strip_names="--no_strip_names" not in lift_args, dist=opt(lift_args, "dist", 10, int))
    syn.liftover([last, last + ".gpu.anchors"] + beds + ["--outfile=" + last + ".gpu.lifted"] + list(lift_args))
    assert read(last + ".gpu.lifted") == read(last + ".orc.lifted")
    return ost


def test_pair_300k(tmp_path, mods):
    st = _pipeline(tmp_path, mods, seed=21, Gq=8000, Gs=7500, n=300000, Kq=13, Ks=12)
    assert st["after_tandem"] > 1000


def test_pair_query_grouped_lines(tmp_path, mods):
    st = _pipeline(tmp_path, mods, seed=27, Gq=6000, Gs=6000, n=400000, Kq=9, Ks=9, group_by_query=True)
    assert st["after_tandem"] > 1000


def test_pair_2m_lastlike_scores(tmp_path, mods):
    _pipeline(tmp_path, mods, seed=22, Gq=30000, Gs=30000, n=2000000, Kq=12, Ks=12, score_style=1)


def test_pair_polyploid_shape(tmp_path, mods):
    # 3 sub-genomes vs a diploid (C3 shape), looser C-score, small clusters, wide liftover
    _pipeline(tmp_path, mods, seed=23, Gq=9000, Gs=3000, n=250000, Kq=21, Ks=7, bf_args=["--cscore=0.5"],
              scan_args=["--min_size=3", "--dist=10"], lift_args=["--dist=15"])


def test_self_comparison(tmp_path, mods):
    _pipeline(tmp_path, mods, seed=24, Gq=6000, Gs=6000, n=120000, Kq=5, Ks=5, self_mode=True,
              bf_args=["--cscore=0.4"], scan_args=["--intrabound=150"])


def test_tie_heavy_liftover(tmp_path, mods):
    # many tiny adjacent blocks + wide rescue radius: cross-block ties decided by cKDTree order
    _pipeline(tmp_path, mods, seed=25, Gq=3000, Gs=3000, n=150000, Kq=3, Ks=3, bf_args=["--cscore=0.2"],
              scan_args=["--min_size=2", "--dist=2"], lift_args=["--dist=14"])


@pytest.mark.parametrize("dist", [0, 1])
def test_liftover_dist_zero_and_one(tmp_path, mods, dist):
    """`scan --dist=1 --liftover` gives liftover_dist 0: nothing is lifted, but every anchor whose pair occurs in the raw
    table stays accepted (AnchorFile.filter_blocks); the oracle was checked against the live reference for dist 0 / 1 / 2."""
    _pipeline(tmp_path, mods, seed=41, Gq=3000, Gs=3000, n=60000, Kq=4, Ks=4, lift_args=["--dist=%d" % dist])


def test_writer_declined_lines_are_spliced(tmp_path, mods):
    """Lines the device writer declines (decimal rounding ties in the evalue, integers beyond
    int32, 17-digit percent identities) are formatted by the host's glibc and spliced in place."""
    import oracle
    from jcvi_b200 import synth
    bf, _ = mods
    d = str(tmp_path)
    qbed, sbed, last = synth.write_dataset(d, "A", "B", 31, 3000, 3000, 60000, Kq=4, Ks=4)
    lines = open(last).read().split("\n")
    out = []
    for i, ln in enumerate(lines):
        f = ln.split("\t")
        if len(f) == 12:
            if i % 5 == 0:
                f[10] = "1.25e-%02d" % (12 + i % 40)          # tie at two significant digits
            elif i % 5 == 1:
                f[3] = "123456789012"                         # %d of a value beyond int32
            elif i % 5 == 2:
                f[10] = "3.45e-%02d" % (11 + i % 30)          # tie again, other parity
        out.append("\t".join(f))
    open(last, "w").write("\n".join(out))
    oracle.blastfilter(last, qbed, sbed, out_path=last + ".orc", cscore=0.5)
    bf.main([last, "--cscore=0.5"])
    got, want = read(last + ".filtered"), read(last + ".orc")
    assert got.count(b"1.2e-") + got.count(b"1.3e-") > 50       # the declined lines really survive the filters
    assert got == want


def test_empty_and_comment_only(tmp_path, mods):
    bf, syn = mods
    from jcvi_b200 import synth
    d = str(tmp_path)
    qbed, sbed, last = synth.write_dataset(d, "A", "B", 1, 200, 200, 500, Kq=2, Ks=2)
    for body in (b"", b"# only a comment\n", b"\n\n\n"):
        with open(last, "wb") as fw:
            fw.write(body)
        bf.main([last])
        assert read(last + ".filtered") == b""
        with pytest.raises(ValueError, match="A total of 0 anchor was found"):
            syn.scan([last + ".filtered", last + ".anchors"])
        assert read(last + ".anchors") == b""


def test_zero_division_and_unsupported(tmp_path, mods):
    bf, _ = mods
    from jcvi_b200 import synth
    d = str(tmp_path)
    qbed, sbed, last = synth.write_dataset(d, "A", "B", 2, 200, 200, 500, Kq=2, Ks=2)
    with open(last, "w") as fw:
        fw.write("Ag000001.1\tBg000002.1\t90.0\t10\t0\t0\t1\t10\t1\t10\t1e-5\t0\n")
    with pytest.raises(ZeroDivisionError):                        # blastfilter.py:235
        bf.main([last])
    with open(last, "w") as fw:
        fw.write("Ag000001.1\tBg000002.1\t90.0\t10\t0\t0\t1\t10\t1\t10\t1e-5\tinf\n")
    with pytest.raises(NotImplementedError):
        bf.main([last])


def test_synteny_scan_points_api(mods):
    """synteny_scan / batch_scan on plain tuples (synteny.py:402-452), checked against a direct
    restatement of the reference loop with its Grouper."""
    _, syn = mods
    rng = np.random.default_rng(7)
    for trial in range(6):
        n = int(rng.integers(50, 3000))
        span = int(rng.integers(60, 1500))
        pts = set()
        while len(pts) < n:
            x = int(rng.integers(0, span)); y = x + int(rng.integers(-40, 41)) if rng.random() < 0.7 else int(rng.integers(0, span))
            pts.add((x, max(y, 0)))
        pts = [(x, y, float(rng.integers(1, 999))) for x, y in pts]
        rng.shuffle(pts)
        xd, yd, N = int(rng.integers(2, 25)), int(rng.integers(2, 25)), int(rng.integers(1, 6))
        got = syn.synteny_scan(list(pts), xd, yd, N)
        # reference algorithm
        mapping = {}
        P = sorted(pts)
        for i in range(len(P)):
            for j in range(i - 1, -1, -1):
                if P[i][0] - P[j][0] > xd:
                    break
                if abs(P[i][1] - P[j][1]) > yd:
                    continue
                a, b = P[i], P[j]
                sa = mapping.setdefault(a, [a])
                sb = mapping.get(b)
                if sb is None:
                    sa.append(b); mapping[b] = sa
                elif sb is not sa:
                    if len(sb) > len(sa):
                        sa, sb = sb, sa
                    sa.extend(sb)
                    for e in sb:
                        mapping[e] = sa
        seen, want = set(), []
        for e, g in mapping.items():
            if e not in seen:
                seen.update(g)
                if min(len(set(p[0] for p in g)), len(set(p[1] for p in g))) >= N:
                    want.append(sorted(g))
        assert got == want, trial


def test_nearest_anchor_matches_scipy(mods):
    _, syn = mods
    spatial = pytest.importorskip("scipy.spatial")
    rng = np.random.default_rng(9)
    for n in (1, 5, 17, 200, 5000):
        x = rng.integers(0, max(4, n // 2), size=n)
        a = np.unique(np.stack([x, x + rng.integers(-3, 4, size=n)], 1), axis=0)
        a = a[rng.permutation(len(a))]
        q = rng.integers(a.min() - 12, a.max() + 13, size=(4000, 2))
        for dist in (5, 10, 20):
            d, i = spatial.cKDTree(a, leafsize=16).query(q, p=1, distance_upper_bound=dist)
            want = [(tuple(p), tuple(a[k])) for p, dd, k in zip(q, d, i) if k != len(a) and dd != 0]
            got = [(tuple(p), nn) for p, nn in syn.synteny_liftover(q, a, dist)]
            assert got == want, (n, dist)


def test_tandems_only_side_outputs(tmp_path, mods):
    """--tandems_only: `.localdups` and `.nolocaldups.bed` (write_localdups / write_new_bed,
    blastfilter.py:164-197) from the device's mother maps, against the oracle."""
    import oracle
    from jcvi_b200 import synth
    bf, _ = mods
    for self_mode in (False, True):
        d = str(tmp_path / ("self" if self_mode else "pair"))
        q, s = ("A", "A") if self_mode else ("A", "B")
        qbed, sbed, last = synth.write_dataset(d, q, s, 61, 4000, 4000, 120000, Kq=5, Ks=5, self_mode=self_mode, p_tandem=0.25)
        if self_mode:
            sbed = qbed
        o = lambda f: os.path.join(d, "orc." + f)
        oracle.blastfilter(last, qbed, sbed, out_path=o("filtered"), cscore=0.5, qdups=o("q.localdups"),
                           sdups=None if self_mode else o("s.localdups"), qnewbed=o("q.nolocaldups.bed"),
                           snewbed=None if self_mode else o("s.nolocaldups.bed"))
        bf.main([last, "--qbed=" + qbed, "--sbed=" + sbed, "--cscore=0.5", "--tandems_only"])
        assert read(last + ".filtered") == read(o("filtered"))
        assert read(os.path.splitext(qbed)[0] + ".localdups") == read(o("q.localdups"))
        assert read(os.path.splitext(qbed)[0] + ".nolocaldups.bed") == read(o("q.nolocaldups.bed"))
        assert len(read(o("q.localdups"))) > 100
        if not self_mode:
            assert read(os.path.splitext(sbed)[0] + ".localdups") == read(o("s.localdups"))
            assert read(os.path.splitext(sbed)[0] + ".nolocaldups.bed") == read(o("s.nolocaldups.bed"))


def test_batch_scan_on_hit_objects(mods):
    """batch_scan with objects carrying qseqid/sseqid/qi/si/score (synteny.py:239-252,437-452):
    groups are scanned in sorted chromosome-pair order."""
    _, syn = mods

    class Hit(object):
        def __init__(self, qseqid, sseqid, qi, si, score):
            self.qseqid, self.sseqid, self.qi, self.si, self.score = qseqid, sseqid, qi, si, score

    rng = np.random.default_rng(11)
    hits, seen = [], set()
    for cp in [("chr2", "c1"), ("chr10", "c1"), ("chr1", "c2"), ("chr1", "c10")]:
        base = int(rng.integers(0, 1000))
        for k in range(400):
            x = base + int(rng.integers(0, 300)); y = x + int(rng.integers(-15, 16)) if rng.random() < 0.8 else int(rng.integers(0, 2000))
            if (cp, x, y) in seen or y < 0:
                continue
            seen.add((cp, x, y))
            hits.append(Hit(cp[0], cp[1], x, y, float(rng.integers(10, 500))))
    rng.shuffle(hits)
    got = syn.batch_scan(hits, xdist=12, ydist=12, N=3)
    want = []
    for cp in sorted(set((h.qseqid, h.sseqid) for h in hits)):
        pts = [(h.qi, h.si, h.score) for h in hits if (h.qseqid, h.sseqid) == cp]
        want.extend(syn.synteny_scan(pts, 12, 12, 3))
    assert got == want and len(got) >= 4


def test_record_reuse_switch_gives_identical_results():
    """jx_set_record_reuse(0): liftover tokenizes the resident table again instead of re-using
    blastfilter's records -- same lifted pairs either way."""
    import torch
    import bench
    from jcvi_b200.engine import Engine
    w = bench.make_workload(7, 300_000, genes=3000)
    e = Engine(0)
    try:
        e.load_pair(w["qbed"], w["sbed"], False)
        dev = e.upload(torch.from_numpy(w["text"]).pin_memory())
        outs = []
        for on in (True, False, True):
            e.set_record_reuse(on)
            r = e.blastfilter(dev, want_text=False, raw=True, skip_arrays=True)
            s = e.scan_filtered(r)
            e.free_filter(r)
            blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
            before = e.tokenizer_timing(reset=True)[1]
            l = e.liftover_text(dev, s.q_rank, s.s_rank, blk, dist=10)
            launches = e.tokenizer_timing(reset=True)[1]
            assert (launches == 0) == on            # reuse on: no tokenizer launch inside liftover
            outs.append((s.q_rank.copy(), l.q_rank.copy(), l.s_rank.copy(), l.block.copy(), l.accepted.copy(), l.raw_score.copy()))
        assert outs[0][1].size > 0
        for o in outs[1:]:
            for a, b in zip(outs[0], o):
                assert np.array_equal(a, b)
    finally:
        e.close()

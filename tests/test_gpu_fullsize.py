"""BASELINE.json configs[1] at full size (40 k + 40 k genes, 20 M hits): size-independent
properties of the GPU path, the 2-slice sharded path against the unsharded one, and the oracle on
a 1 M-line prefix."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import bench
    from jcvi_b200.engine import Engine
    w = bench.make_workload(100, 20_000_000)
    e = Engine(0)
    e.load_pair(w["qbed"], w["sbed"], False)
    yield w, e
    e.close()


def test_filter_scan_liftover_properties(big):
    w, e = big
    text = w["text"]
    f = e.blastfilter(text, cscore=0.7, want_text=True)
    n = f.n
    assert 100_000 < n < 2_000_000
    # (1) output order: score descending, ties in file order
    sc, off = f.score.astype(np.float64), f.line_off.astype(np.int64)
    assert (np.diff(sc) <= 0).all()
    tie = np.diff(sc) == 0
    assert (np.diff(off)[tie] > 0).all()
    # (2) pairs are unique, never q == s by name (distinct prefixes here), ranks in range
    key = (f.q_rank.astype(np.int64) << 32) | f.s_rank.astype(np.int64)
    assert len(np.unique(key)) == n
    assert f.q_rank.min() >= 0 and f.q_rank.max() < 40000 and f.s_rank.max() < 40000
    # (3) every survivor's line starts right after a newline of the input and its text agrees
    assert (text[off[off > 0] - 1] == 10).all()
    lines = bytes(memoryview(f.text)).split(b"\n")
    assert len(lines) == n + 1 and lines[-1] == b""
    qa, sa = w["qbed"].tables()["accns"], w["sbed"].tables()["accns"]
    for t in (0, 1, n // 2, n - 1):
        c = lines[t].split(b"\t")
        assert len(c) == 12 and c[0].decode() == qa[f.q_rank[t]] and c[1].decode() == sa[f.s_rank[t]]
        assert float(c[11]) == float("%.3g" % f.score[t])
    # (4) fixed point: the filtered table passes through the filter unchanged when C-score and
    #     tandem collapsing are off (names are already gene-level, pairs unique, order kept)
    g = e.blastfilter(np.frombuffer(bytes(memoryview(f.text)), dtype=np.uint8), cscore=0, tandem_Nmax=0, strip_names=False)
    assert g.n == n
    assert np.array_equal(g.q_rank, f.q_rank) and np.array_equal(g.s_rank, f.s_rank)
    # (5) scan: anchors are filtered pairs, clusters ascending, sizes >= min_size in distinct genes
    s = e.scan_text(f.text)
    akey = (s.q_rank.astype(np.int64) << 32) | s.s_rank.astype(np.int64)
    assert np.isin(akey, key).all() and len(np.unique(akey)) == s.n_anchors
    bs = s.block_start
    for b in (0, s.n_blocks // 2, s.n_blocks - 1):
        q, y = s.q_rank[bs[b]:bs[b + 1]], s.s_rank[bs[b]:bs[b + 1]]
        assert (np.diff(q) >= 0).all() and min(len(np.unique(q)), len(np.unique(y))) >= 4
    # in-memory hand-over == text round trip
    raw = e.blastfilter(text, cscore=0.7, want_text=False, raw=True)
    s2 = e.scan_filtered(raw)
    e.free_filter(raw)
    assert np.array_equal(s.q_rank, s2.q_rank) and np.array_equal(s.block_start, s2.block_start) and np.array_equal(s.score, s2.score)
    # (6) liftover: lifted pairs are new, within dist of an anchor of their block, every anchor accepted
    blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(bs))
    l = e.liftover_text(text, s.q_rank, s.s_rank, blk, dist=10)
    # anchors whose pair only exists after tandem-mother substitution are NOT in the raw table
    # (the reference removes them in filter_blocks); all others are accepted with their raw score
    assert l.n_lifted > 0 and 0.5 < l.accepted.mean() <= 1.0
    assert (l.raw_score[l.accepted > 0] > 0).all() and (l.raw_score[l.accepted == 0] == 0).all()
    lkey = (l.q_rank.astype(np.int64) << 32) | l.s_rank.astype(np.int64)
    assert not np.isin(lkey, akey).any() and len(np.unique(lkey)) == l.n_lifted
    assert (np.diff(l.block) >= 0).all()
    rng = np.random.default_rng(0)
    for t in rng.integers(0, l.n_lifted, size=200):
        m = blk == l.block[t]
        d = np.abs(s.q_rank[m] - l.q_rank[t]) + np.abs(s.s_rank[m] - l.s_rank[t])
        assert 0 < d.min() < 10


def test_sharded_equals_unsharded_at_full_size(big):
    import torch
    from jcvi_b200.engine import Engine
    from jcvi_b200.sharded import slice_bounds
    w, e0 = big
    text = w["text"]
    want = bytes(memoryview(e0.blastfilter(text, cscore=0.7).text))
    e1 = Engine(0)
    e1.load_pair(w["qbed"], w["sbed"], False)
    parts = []
    handles = []
    for r, e in ((0, e0), (1, e1)):
        lo, hi = slice_bounds(text, r, 2)
        handles.append((e, e.cscore_begin(text[lo:hi])))
    t0 = handles[0][0].best_table_tensor(*handles[0][1]); t1 = handles[1][0].best_table_tensor(*handles[1][1])
    m = torch.maximum(t0, t1); t0.copy_(m); t1.copy_(m); torch.cuda.synchronize()
    for e, _ in handles:
        parts.append(e.cscore_finish(0.7)[0])
    got = bytes(memoryview(e0.blastfilter(np.concatenate(parts), cscore=0).text))
    e1.close()
    assert got == want


def test_prefix_against_oracle(big, tmp_path):
    import bench
    import oracle
    from jcvi_b200.compara import blastfilter as bf
    w, _ = big
    sample = bench.head_lines(w["text"], 1_000_003)
    last = str(tmp_path / "A.B.last")
    sample.tofile(last)
    oracle.blastfilter(last, w["qbed_path"], w["sbed_path"], out_path=last + ".orc")
    bf.main([last, "--qbed=" + w["qbed_path"], "--sbed=" + w["sbed_path"]])
    assert open(last + ".filtered", "rb").read() == open(last + ".orc", "rb").read()


def _whole_table_vs_oracle(tmp_path, text, qbed_path, sbed_path, bf_args, scan_args, lift_dist):
    """`.filtered`, `.anchors`, `.lifted.anchors` of the drop-in CLI functions against the CPU oracle, byte for byte.  The
    oracle's two long stages (blastfilter and liftover both parse the whole raw table) run side by side: liftover is given
    the GPU's `.anchors`, which is legitimate only because those are asserted identical to the oracle's own afterwards."""
    import threading
    import oracle
    from jcvi_b200.compara import blastfilter as bf, synteny as syn
    last = str(tmp_path / "A.B.last")
    text.tofile(last)
    beds = ["--qbed=" + qbed_path, "--sbed=" + sbed_path]
    bf.main([last] + beds + bf_args)
    ganch = str(tmp_path / "g.anchors")
    glift = syn.scan([last + ".filtered", ganch] + beds + scan_args + ["--liftover=" + last, "--liftover_dist=%d" % lift_dist])
    kw = {a.split("=")[0].lstrip("-"): a.split("=")[1] for a in bf_args + scan_args}
    err = []

    def run_bf():
        try:
            oracle.blastfilter(last, qbed_path, sbed_path, out_path=last + ".orc", cscore=float(kw.get("cscore", 0.7)),
                               tandem_Nmax=int(kw.get("tandem_Nmax", 10)))
        except Exception as e:      # pragma: no cover
            err.append(e)

    def run_lift():
        try:
            oracle.liftover(last, ganch, qbed_path, sbed_path, out_path=str(tmp_path / "o.lifted"), dist=lift_dist)
        except Exception as e:      # pragma: no cover
            err.append(e)
    th = [threading.Thread(target=run_bf), threading.Thread(target=run_lift)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not err, err
    rd = lambda p: open(p, "rb").read()
    assert rd(last + ".filtered") == rd(last + ".orc")
    oracle.scan(last + ".orc", str(tmp_path / "o.anchors"), qbed_path, sbed_path, dist=int(kw.get("dist", 20)),
                min_size=int(kw.get("min_size", 4)))
    assert rd(ganch) == rd(str(tmp_path / "o.anchors"))
    assert rd(glift) == rd(str(tmp_path / "o.lifted"))
    return rd(ganch).count(b"\n"), rd(glift).count(b"L\n")


@pytest.mark.skipif(os.environ.get("JCVI_B200_SKIP_WHOLE_C2") == "1", reason="whole-table oracle run switched off")
def test_whole_c2_table_against_oracle(big, tmp_path):
    """BASELINE configs[1] -- ALL 20 M lines, every stage, byte for byte against the oracle (a few minutes of CPU time)."""
    w, _ = big
    na, nl = _whole_table_vs_oracle(tmp_path, w["text"], w["qbed_path"], w["sbed_path"], [], [], 10)
    assert na > 50_000 and nl > 50_000


def test_c3_shape_8m_against_oracle(tmp_path):
    """BASELINE configs[2] shape (hexaploid vs diploid, 110 k vs 35 k genes, C-score 0.7) at 8 M hits, every stage."""
    import bench
    text, n, qbed, sbed = bench.ss_workload("c3", 8_000_000, 1, 0, str(tmp_path))
    na, nl = _whole_table_vs_oracle(tmp_path, text, str(tmp_path / "W.bed"), str(tmp_path / "H.bed"), ["--cscore=0.7"], [], 10)
    assert na > 1000 and nl > 100


def test_c5_shape_6m_against_oracle(tmp_path):
    """BASELINE configs[4] shape: sparse anchors (min_size=4), liftover-heavy with dist=20, at 6 M raw hits."""
    import bench
    text, n, qbed, sbed = bench.ss_workload("c5", 6_000_000, 1, 0, str(tmp_path))
    na, nl = _whole_table_vs_oracle(tmp_path, text, str(tmp_path / "W.bed"), str(tmp_path / "H.bed"), [],
                                    ["--min_size=4", "--dist=20"], 20)
    assert na > 1000 and nl > 1000

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

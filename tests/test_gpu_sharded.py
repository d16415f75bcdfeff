"""Sharding one genome pair: two "ranks" emulated on one GPU as two engines (two jx contexts)
whose best-score tables are max-reduced on the device -- the same data path as the NCCL run
(tools/check_sharded.py covers the real 2-GPU all-reduce)."""
import os

import numpy as np
import pytest

from conftest import read

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("self_mode,cscore", [(False, 0.7), (False, 0.4), (True, 0.5)])
def test_two_slices_equal_unsharded(tmp_path, self_mode, cscore):
    import torch
    from jcvi_b200 import synth
    from jcvi_b200.engine import Engine
    from jcvi_b200.formats.bed import Bed
    from jcvi_b200.sharded import slice_bounds

    d = str(tmp_path)
    q, s = ("A", "A") if self_mode else ("A", "B")
    qbed_p, sbed_p, last = synth.write_dataset(d, q, s, 41, 5000, 5000, 200000, Kq=6, Ks=6, self_mode=self_mode)
    qbed = Bed(qbed_p)
    sbed = qbed if self_mode else Bed(sbed_p)
    text = np.fromfile(last, dtype=np.uint8)

    e0, e1 = Engine(0), Engine(0)
    for e in (e0, e1):
        e.load_pair(qbed, sbed, self_mode)
    whole = e0.blastfilter(text, cscore=cscore)
    want = bytes(memoryview(whole.text))

    # "ranks" 0 and 1: uneven cut to exercise the boundary logic
    lo0, hi0 = slice_bounds(text, 0, 2)
    lo1, hi1 = slice_bounds(text, 1, 2)
    assert lo0 == 0 and hi0 == lo1 and hi1 == len(text) and text[hi0 - 1] == 10
    p0, n0 = e0.cscore_begin(text[lo0:hi0])
    p1, n1 = e1.cscore_begin(text[lo1:hi1])
    t0, t1 = e0.best_table_tensor(p0, n0), e1.best_table_tensor(p1, n1)
    m = torch.maximum(t0, t1)                     # what ncclAllReduce(max) leaves on every rank
    assert (m >= 0).all()
    t0.copy_(m); t1.copy_(m)
    torch.cuda.synchronize()
    l0, c0 = e0.cscore_finish(cscore)
    l1, c1 = e1.cscore_finish(cscore)
    assert 0 < c0 + c1 < 0.5 * text.tobytes().count(b"\n")
    merged = np.concatenate([l0, l1])
    res = e0.blastfilter(merged, cscore=0)
    assert bytes(memoryview(res.text)) == want

    # liftover sharded the same way
    sc = e0.scan_text(np.frombuffer(want, dtype=np.uint8), N=3)
    blk = np.repeat(np.arange(sc.n_blocks, dtype=np.int32), np.diff(sc.block_start))
    full = e0.liftover_text(text, sc.q_rank, sc.s_rank, blk, dist=12)
    n0_, _ = e0.near_lines(text[lo0:hi0], sc.q_rank, sc.s_rank, dist=12)
    n1_, _ = e1.near_lines(text[lo1:hi1], sc.q_rank, sc.s_rank, dist=12)
    part = e0.liftover_text(np.concatenate([n0_, n1_]), sc.q_rank, sc.s_rank, blk, dist=12)
    for k in ("q_rank", "s_rank", "block", "accepted"):
        assert np.array_equal(getattr(full, k), getattr(part, k)), k
    assert np.array_equal(full.score, part.score) and np.array_equal(full.raw_score[full.accepted > 0], part.raw_score[part.accepted > 0])
    assert full.n_lifted > 0
    e0.close(); e1.close()

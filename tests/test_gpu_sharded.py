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


def test_device_results_and_record_reuse_equal_the_host_route(tmp_path):
    """jx_cscore_finish_device / jx_near_lines_device (packed lines left on the device for the NCCL exchange) return
    the bytes of the host variants; near_lines on the resident slice that cscore_begin tokenized re-uses its
    records (one tokenizer pass per slice and step) with the same result; the merged device buffer is accepted
    as device text by the later stages."""
    import torch
    from jcvi_b200 import synth
    from jcvi_b200.engine import Engine
    from jcvi_b200.formats.bed import Bed
    from jcvi_b200.sharded import slice_bounds

    d = str(tmp_path)
    qbed_p, sbed_p, last = synth.write_dataset(d, "A", "B", 43, 5000, 5000, 300000, Kq=6, Ks=6)
    qbed, sbed = Bed(qbed_p), Bed(sbed_p)
    text = np.fromfile(last, dtype=np.uint8)
    e0, e1 = Engine(0), Engine(0)
    for e in (e0, e1):
        e.load_pair(qbed, sbed, False)
    want = e0.blastfilter(text, cscore=0.7)
    sc = e0.scan_text(want.text, N=3)
    blk = np.repeat(np.arange(sc.n_blocks, dtype=np.int32), np.diff(sc.block_start))
    full = e0.liftover_text(text, sc.q_rank, sc.s_rank, blk, dist=12)

    host_lines, host_near, dev_lines, dev_near = [], [], [], []
    slices = [slice_bounds(text, r, 2) for r in range(2)]
    resident = [e.upload(text[lo:hi]) for e, (lo, hi) in zip((e0, e1), slices)]
    for device in (False, True):
        tabs = [e.best_table_tensor(*e.cscore_begin(dv)) for e, dv in zip((e0, e1), resident)]
        m = torch.maximum(tabs[0], tabs[1])
        tabs[0].copy_(m); tabs[1].copy_(m)
        torch.cuda.synchronize()
        for e, dv in zip((e0, e1), resident):
            launches = e.tokenizer_timing(reset=True)[1]
            a, _ = e.cscore_finish(0.7, device=device)
            a = a.clone() if device else a
            b, _ = e.near_lines(dv, sc.q_rank, sc.s_rank, dist=12, device=device)      # re-uses the slice's records
            assert e.tokenizer_timing(reset=True)[1] == 0, "near_lines tokenized the resident slice again"
            (dev_lines if device else host_lines).append(a)
            (dev_near if device else host_near).append(b.clone() if device else b)
    for h, g in zip(host_lines + host_near, dev_lines + dev_near):
        assert g.is_cuda and np.array_equal(h, g.cpu().numpy())
    merged = torch.cat(dev_lines)
    res = e0.blastfilter(merged, cscore=0)                   # device text that never was in host memory
    assert bytes(memoryview(res.text)) == bytes(memoryview(want.text))
    part = e0.liftover_text(torch.cat(dev_near), sc.q_rank, sc.s_rank, blk, dist=12)
    for k in ("q_rank", "s_rank", "block", "accepted", "score"):
        assert np.array_equal(getattr(full, k), getattr(part, k)), k
    e0.close(); e1.close()

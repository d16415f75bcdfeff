"""jx_pipeline (the three stages as one call, everything between them on the device) == the three separate calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blocks(s):
    return np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))


@pytest.mark.parametrize("case", ["pair", "pair_no_tandem", "polyploid", "self"])
def test_pipeline_equals_separate_calls(tmp_path, case):
    from jcvi_b200 import synth
    from jcvi_b200.engine import get_engine
    from jcvi_b200.formats.bed import Bed
    kw = dict(pair=dict(Gq=8000, Gs=7500, n=300000, Kq=13, Ks=12), pair_no_tandem=dict(Gq=5000, Gs=5000, n=150000, Kq=7, Ks=9),
              polyploid=dict(Gq=9000, Gs=3000, n=250000, Kq=21, Ks=7), self=dict(Gq=6000, Gs=6000, n=120000, Kq=5, Ks=5))[case]
    self_mode = case == "self"
    q, s = ("A", "A") if self_mode else ("A", "B")
    qbed, sbed, last = synth.write_dataset(str(tmp_path), q, s, 77, kw["Gq"], kw["Gs"], kw["n"], self_mode=self_mode, Kq=kw["Kq"], Ks=kw["Ks"])
    qb = Bed(qbed)
    sb = qb if self_mode else Bed(sbed)
    eng = get_engine()
    eng.load_pair(qb, sb, self_mode)
    text = np.fromfile(last, dtype=np.uint8)
    opts = dict(cscore=0.5 if case != "pair" else 0.7, tandem_Nmax=0 if case == "pair_no_tandem" else 10)
    scan_kw = dict(N=3 if case == "polyploid" else 4, xdist=10 if case == "polyploid" else 20, ydist=10 if case == "polyploid" else 20,
                   intrabound=150)
    dist = 15 if case == "polyploid" else 10
    # separate calls (host text each time)
    f0 = eng.blastfilter(text, want_text=True, **opts)
    s0 = eng.scan_text(f0.text, **scan_kw)
    l0 = eng.liftover_text(text, s0.q_rank, s0.s_rank, _blocks(s0), dist=dist)
    assert s0.n_anchors > 100
    # one call, host text
    f1, s1, l1 = eng.pipeline(text, want_text=True, skip_arrays=False, liftover_dist=dist, **opts, **scan_kw)
    # one call, resident text
    dev = eng.upload(text)
    f2, s2, l2 = eng.pipeline(dev, want_text=True, skip_arrays=False, liftover_dist=dist, **opts, **scan_kw)
    for f, s_, l in ((f1, s1, l1), (f2, s2, l2)):
        assert bytes(memoryview(f.text)) == bytes(memoryview(f0.text))
        assert np.array_equal(f.q_rank, f0.q_rank) and np.array_equal(f.s_rank, f0.s_rank) and np.array_equal(f.score, f0.score)
        assert np.array_equal(s_.q_rank, s0.q_rank) and np.array_equal(s_.s_rank, s0.s_rank) and np.array_equal(s_.score, s0.score)
        assert np.array_equal(s_.block_start, s0.block_start)
        for name in ("q_rank", "s_rank", "score", "block", "accepted", "raw_score"):
            assert np.array_equal(getattr(l, name), getattr(l0, name)), name

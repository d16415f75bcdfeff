"""Chromosome-pair sharding of one genome pair (jcvi_b200.sharded.run_sharded): all ranks emulated in one process on one
GPU (LocalComm: the collectives are tensor copies), every exchange and owner-side stage exercised; results must equal
the single-GPU jx_pipeline arrays."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["owners", "gather"])
@pytest.mark.parametrize("case,world", [("pair", 3), ("polyploid", 4), ("pair", 1), ("no_tandem", 2)])
def test_sharded_equals_single_gpu(tmp_path, case, world, mode):
    from jcvi_b200 import sharded, synth
    from jcvi_b200.engine import Engine, get_engine
    from jcvi_b200.formats.bed import Bed
    kw = dict(pair=dict(Gq=8000, Gs=7500, n=300000, Kq=13, Ks=12), polyploid=dict(Gq=9000, Gs=3000, n=250000, Kq=21, Ks=7),
              no_tandem=dict(Gq=5000, Gs=5000, n=150000, Kq=7, Ks=9))[case]
    qbed, sbed, last = synth.write_dataset(str(tmp_path), "A", "B", 91, kw["Gq"], kw["Gs"], kw["n"], Kq=kw["Kq"], Ks=kw["Ks"])
    qb, sb = Bed(qbed), Bed(sbed)
    text = np.fromfile(last, dtype=np.uint8)
    opts = dict(cscore=0.5 if case == "polyploid" else 0.7, tandem_Nmax=0 if case == "no_tandem" else 10)
    scan_kw = dict(N=3 if case == "polyploid" else 4, xdist=10 if case == "polyploid" else 20, ydist=10 if case == "polyploid" else 20)
    dist = 15 if case == "polyploid" else 10
    e0 = get_engine()
    e0.load_pair(qb, sb, False)
    f0, s0, l0 = e0.pipeline(text, want_text=False, skip_arrays=False, liftover_dist=dist, **opts, **scan_kw)
    assert s0.n_anchors > 100 and l0.n_lifted > 10
    engines = [Engine(0) for _ in range(world)]
    slices = []
    for r in range(world):
        lo, hi = sharded.slice_bounds(text, r, world)
        slices.append(text[lo:hi])
    res = sharded.run_sharded(slices, qb, sb, comm=sharded.LocalComm(world), engines=engines, liftover_dist=dist, mode=mode, **opts,
                              **scan_kw)
    assert res.mode == mode
    f = res.filtered
    assert f.n == f0.n
    assert np.array_equal(f.q_rank, f0.q_rank) and np.array_equal(f.s_rank, f0.s_rank) and np.array_equal(f.score, f0.score)
    s = res.scan
    assert np.array_equal(s.q_rank, s0.q_rank) and np.array_equal(s.s_rank, s0.s_rank) and np.array_equal(s.score, s0.score)
    assert np.array_equal(s.block_start, s0.block_start)
    l = res.lifted
    for name in ("q_rank", "s_rank", "score", "block", "accepted", "raw_score"):
        assert np.array_equal(getattr(l, name), getattr(l0, name)), name
    for e in engines:
        e.shard_reset()
        e.close()


@pytest.mark.parametrize("what", ["unsupported", "zerodiv"])
def test_rank_local_failure_stops_every_rank(tmp_path, what):
    """An unsupported line / a score <= 0 in ONE slice: the failing rank keeps taking part with empty data, its status
    rides on the histogram all-gather, and the driver raises the reference's exception instead of hanging."""
    from jcvi_b200 import sharded, synth
    from jcvi_b200.engine import Engine
    from jcvi_b200.formats.bed import Bed
    qbed, sbed, last = synth.write_dataset(str(tmp_path), "A", "B", 33, 3000, 3000, 60000, Kq=5, Ks=5)
    qb, sb = Bed(qbed), Bed(sbed)
    text = np.fromfile(last, dtype=np.uint8)
    world = 3
    slices = []
    for r in range(world):
        lo, hi = sharded.slice_bounds(text, r, world)
        slices.append(text[lo:hi].copy())
    line = bytes(slices[1][: int(np.flatnonzero(slices[1] == 10)[0])]).split(b"\t")
    if what == "unsupported":
        line[11] = b"inf"
    else:
        # blastfilter.py:235 divides by max(best score of the two genes): 0 only when neither gene scores anywhere else
        qn, sn = line[0].split(b".")[0], line[1].split(b".")[0]
        for r in range(world):
            rows = [x for x in bytes(slices[r]).split(b"\n") if x and qn not in x and sn not in x]
            slices[r] = np.frombuffer(b"\n".join(rows) + b"\n", dtype=np.uint8)
        line[11] = b"0"
    bad = np.frombuffer(b"\t".join(line) + b"\n", dtype=np.uint8)
    slices[1] = np.concatenate([bad, slices[1]])
    engines = [Engine(0) for _ in range(world)]
    try:
        with pytest.raises(NotImplementedError if what == "unsupported" else ZeroDivisionError):
            sharded.run_sharded(slices, qb, sb, comm=sharded.LocalComm(world), engines=engines)
    finally:
        for e in engines:
            e.shard_reset()
            e.close()

#!/usr/bin/env python
"""Library phase trace (JX_TRACE) of one device-resident C3 step (200 M hits) -- where the non-tokenizer time goes at
full size.  Reporting aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from jcvi_b200 import synth
from jcvi_b200.formats.bed import Bed
from jcvi_b200.engine import get_engine
import tempfile
hits = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
d = tempfile.mkdtemp()
rng = np.random.default_rng(3)
gq = synth.Genome("W", 110_000, 21, rng); gs = synth.Genome("H", 35_000, 7, rng)
parts, n = [], 0
while n < hits:
    a = synth.make_hits(gq, gs, min(20_000_000, hits - n), rng)
    parts.append(synth.format_hits(gq, gs, a, header=not parts)); n += len(a[0]); del a
text = np.concatenate(parts); del parts
qp, sp = os.path.join(d, "W.bed"), os.path.join(d, "H.bed")
gq.write_bed(qp, np.random.default_rng(10)); gs.write_bed(sp, np.random.default_rng(11))
eng = get_engine(0); eng.load_pair(Bed(qp), Bed(sp), False)
dev = eng.upload(text); torch.cuda.synchronize()
def step():
    t0 = time.perf_counter()
    r = eng.blastfilter(dev, cscore=0.7, want_text=False, raw=True, skip_arrays=True)
    t1 = time.perf_counter()
    s = eng.scan_filtered(r); eng.free_filter(r)
    t2 = time.perf_counter()
    blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
    l = eng.liftover_text(dev, s.q_rank, s.s_rank, blk, dist=10)
    t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
step(); step()
print("untraced step (ms): blastfilter %.2f scan %.2f liftover %.2f" % step(), flush=True)
print("---- traced step ----", file=sys.stderr, flush=True)
pass  # run with JX_TRACE=1: the library prints its (synchronising) phase times for every call
step()

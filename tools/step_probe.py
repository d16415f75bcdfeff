#!/usr/bin/env python
"""Wall-clock split of one device-resident step (blastfilter / scan / liftover) on the C2 workload.
With JX_TRACE=1 in the environment the library also prints its phase timings (synchronising)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = bench.make_workload(100, hits)
eng = get_engine(0)
eng.load_pair(w["qbed"], w["sbed"], False)
pin = torch.from_numpy(w["text"]).pin_memory()
dt = eng.upload(pin)
torch.cuda.synchronize()
acc = np.zeros(3)
N = 8
for it in range(N + 3):
    if it == N + 2:
        print("---- last step ----", file=sys.stderr, flush=True)
    if os.environ.get("PROBE_PIPELINE"):
        t0 = time.perf_counter()
        f, s, l = eng.pipeline(dt, want_text=bool(int(os.environ.get("PROBE_TEXT", "0"))), skip_arrays=True, liftover_dist=10)
        t3 = time.perf_counter()
        if it >= 3:
            acc += [t3 - t0, 0, 0]
        continue
    t0 = time.perf_counter()
    r = eng.blastfilter(dt, want_text=False, raw=True, skip_arrays=True)
    t1 = time.perf_counter()
    s = eng.scan_filtered(r)
    eng.free_filter(r)
    t2 = time.perf_counter()
    blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
    l = eng.liftover_text(dt, s.q_rank, s.s_rank, blk, dist=10)
    t3 = time.perf_counter()
    if it >= 3:
        acc += [t1 - t0, t2 - t1, t3 - t2]
print("per step (ms): blastfilter %.3f  scan %.3f  liftover %.3f  total %.3f" % (*(acc / N * 1e3), acc.sum() / N * 1e3))
print("launches per step:", eng.lib.jx_launch_count(eng.ctx) // (N + 3))

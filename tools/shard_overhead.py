#!/usr/bin/env python
"""Single GPU: jx_pipeline vs run_sharded with one (or W emulated) rank(s) on the same table -- what the sharded driver's
orchestration costs without any network.  python tools/shard_overhead.py [hits] [W]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jcvi_b200 import sharded
from jcvi_b200.engine import Engine, get_engine

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1
import tempfile
d = tempfile.mkdtemp()
text, n, qbed, sbed = bench.ss_workload("c3", hits, 1, 0, d)
eng = get_engine(0)
eng.load_pair(qbed, sbed, False)
dev = eng.upload(text)
torch.cuda.synchronize()
def T(f, k=6):
    f(); f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
print("pipeline        %.3f ms" % T(lambda: eng.pipeline(dev, want_text=False, skip_arrays=True, liftover_dist=10)))
engines = [Engine(0) for _ in range(W)]
sl = []
for r in range(W):
    lo, hi = sharded.slice_bounds(text, r, W)
    sl.append(engines[r].upload(text[lo:hi]))
comm = sharded.LocalComm(W)
marks = []
def mark(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))
print("run_sharded W=%d %.3f ms" % (W, T(lambda: sharded.run_sharded(sl, qbed, sbed, comm=comm, engines=engines, want_filtered=False, liftover_dist=10))))
del marks[:]; mark("start")
sharded.run_sharded(sl, qbed, sbed, comm=comm, engines=engines, want_filtered=False, liftover_dist=10, mark=mark)
print({marks[i][0]: round((marks[i][1] - marks[i - 1][1]) * 1e3, 3) for i in range(1, len(marks))})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): sharded.run_sharded(sl, qbed, sbed, comm=comm, engines=engines, want_filtered=False, liftover_dist=10)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

#!/usr/bin/env python
"""Phase times of the end-to-end call (jx_pipeline on the pinned HOST table).  JX_TRACE=1 python tools/e2e_trace.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine
w = bench.make_workload(100, 20_000_000)
eng = get_engine(0); eng.load_pair(w["qbed"], w["sbed"], False)
pin = torch.from_numpy(w["text"]).pin_memory()
for it in range(5):
    if it == 4:
        print("---- last call ----", file=sys.stderr, flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f, s, l = eng.pipeline(pin, want_text=True, skip_arrays=True, liftover_dist=10)
    t1 = time.perf_counter()
    print("call %d: %.2f ms" % (it, (t1 - t0) * 1e3), flush=True)

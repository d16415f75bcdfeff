import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine
w = bench.make_workload(100, 20_000_000)
eng = get_engine(0); eng.load_pair(w["qbed"], w["sbed"], False)
pin = torch.from_numpy(w["text"]).pin_memory()
dt = eng.upload(pin)
for i in range(3):
    r = eng.blastfilter(dt, want_text=False, raw=True, skip_arrays=True); eng.free_filter(r)
print("---- traced call ----", file=sys.stderr, flush=True)
eng.lib.jx_set_trace(eng.ctx, 1) if hasattr(eng.lib, "jx_set_trace") else None
r = eng.blastfilter(dt, want_text=False, raw=True, skip_arrays=True); eng.free_filter(r)

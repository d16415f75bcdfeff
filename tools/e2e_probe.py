import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine
w = bench.make_workload(100, 20_000_000)
eng = get_engine(0); eng.load_pair(w["qbed"], w["sbed"], False)
pin = torch.from_numpy(w["text"]).pin_memory()
def T(f, n=4):
    ts=[]
    for _ in range(n):
        torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
    return min(ts), r
dt = eng.upload(pin)
print("bf notext (held)", T(lambda: eng.blastfilter(dt, want_text=False))[0])
print("bf text (held)", T(lambda: eng.blastfilter(dt, want_text=True))[0])
r = eng.blastfilter(dt, want_text=False, raw=True)
t, s = T(lambda: eng.scan_filtered(r)); print("scan_filtered", t)
blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
def lift_cached():
    rr = eng.blastfilter(dt, want_text=False, raw=True); eng.free_filter(rr)
    t0=time.perf_counter(); l = eng.liftover_text(dt, s.q_rank, s.s_rank, blk, dist=10); torch.cuda.synchronize(); return time.perf_counter()-t0
print("liftover (cached recs)", min(lift_cached() for _ in range(4)))
print("liftover (tokenize)", T(lambda: eng.liftover_text(dt, s.q_rank, s.s_rank, blk, dist=10))[0])
f = eng.blastfilter(dt, want_text=True)
print("scan_text", T(lambda: eng.scan_text(f.text))[0])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    rr = eng.blastfilter(dt, want_text=False, raw=True); s2 = eng.scan_filtered(rr); eng.free_filter(rr)
    l = eng.liftover_text(dt, s.q_rank, s.s_rank, blk, dist=10)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)

import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine
w = bench.make_workload(100, 20_000_000)
eng = get_engine(0); eng.load_pair(w["qbed"], w["sbed"], False)
pin = torch.from_numpy(w["text"]).pin_memory()
def T(f, n=3):
    ts=[]
    for _ in range(n):
        torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t)
    return min(ts), r
print("upload+sync", T(lambda: (eng.upload(pin), eng.lib.jx_tokenizer_timing(eng.ctx,0,None,None,None,None)))[0])
dt = eng.upload(pin)
print("bf notext (held)", T(lambda: eng.blastfilter(dt, want_text=False))[0])
print("bf text (held)", T(lambda: eng.blastfilter(dt, want_text=True))[0])
print("bf raw text (held)", T(lambda: eng.free_filter(eng.blastfilter(dt, want_text=True, raw=True)))[0])
def full():
    d = eng.upload(pin); return eng.blastfilter(d, want_text=True)
print("upload+bf text", T(full)[0])
def full2():
    d = eng.upload(pin); return eng.blastfilter(d, want_text=False)
print("upload+bf notext", T(full2)[0])
print("bf host ptr notext", T(lambda: eng.blastfilter(pin, want_text=False))[0])
print("cpus", os.cpu_count())

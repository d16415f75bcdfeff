#!/usr/bin/env python
"""`formats.blast cscore` on the C2 table: the drop-in's phases on one B200 and
the CPU oracle on a 500 k-line sample of the same table (tuning / reporting aid, not bench.py)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine
from jcvi_b200.formats import blast as gblast

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = bench.make_workload(100, hits)
text = w["text"]
eng = get_engine(0)
pin = torch.from_numpy(text).pin_memory()
for rep in range(3):
    t0 = time.perf_counter()
    dev = eng.upload(pin)
    lines, sides, st = eng.distinct_names(dev, strip_names=True, seed=0)
    t1 = time.perf_counter()
    names = sorted({gblast.gene_name(ln.split(None, 2)[s]) for ln, s in zip(lines, sides.tolist())})
    t2 = time.perf_counter()
    eng.load_names(names)
    t3 = time.perf_counter()
    eng.cscore_begin(dev, strip_names=True)
    assert eng.session_counts()["unmapped"] == 0
    best = eng.best_table(len(names))
    kept, n_kept = eng.cscore_finish(0.9999)
    t4 = time.perf_counter()
    eng.release_text()
    t5 = time.perf_counter()
    pairs = gblast.cscore_pairs(pin.numpy(), 0.9999, True, engine=eng)
    t6 = time.perf_counter()
    print("rep %d: upload+distinct_names %.1f ms (%d names) | host name list %.1f ms | load_names %.1f ms | "
          "cscore_begin+finish %.1f ms (%d kept lines) | whole cscore_pairs() incl. host dict %.1f ms (%d pairs)" % (
              rep, (t1 - t0) * 1e3, len(names), (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, n_kept, (t6 - t5) * 1e3, len(pairs)))
for rep in range(3):
    t0 = time.perf_counter()
    out, _, st = eng.blast_cscore(pin.numpy(), cutoff=0.9999)
    t1 = time.perf_counter()
    print("jx_blast_cscore (one C-ABI call, host buffer in, cscore text out): %.1f ms, %d output lines" % ((t1 - t0) * 1e3, out.count(b"\n")))
print("GPU drop-in: %.3g hits/s through cscore (whole call, host buffers)" % (w["n_hits"] / (t1 - t0)))
# CPU oracle on a sample
import oracle
d = tempfile.mkdtemp()
sample = bench.head_lines(text, 500003)
f = os.path.join(d, "s.last"); open(f, "wb").write(bytes(sample))
t0 = time.perf_counter(); oracle.cscore(f, f + ".cs"); t1 = time.perf_counter()
n = bytes(sample).count(b"\n")
print("CPU oracle (1 thread): %d lines in %.2f s = %.3g hits/s" % (n, t1 - t0, n / (t1 - t0)))

#!/usr/bin/env python
"""`formats.blast filter` (blast.py:620-709) on the C2 table: one C-ABI call per option set on
one B200 (host buffer in, kept rows out; and with the table already resident), next to the CPU oracle on a
500 k-line sample of the same table (tuning / reporting aid, not bench.py)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from jcvi_b200.engine import get_engine

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = bench.make_workload(100, hits)
text = w["text"]
eng = get_engine(0)
pin = torch.from_numpy(text).pin_memory()
devnull = open(os.devnull, 'wb')
sets = (("catalog.ortholog's --pctid=98 --inverse --noself", dict(pctid=98.0, inverse=True, noself=True)),
        ("--pctid=60 --hitlen=150 --score=400 --evalue=1e-30", dict(pctid=60.0, hitlen=150, score=400, evalue=1e-30)))
for tag, kw in sets:
    for rep in range(3):
        t0 = time.perf_counter()
        out, st = eng.blast_filter(pin.numpy(), sink=devnull, **kw)
        t1 = time.perf_counter()
    print("%s | host buffer in, rows out: %.1f ms = %.3g hits/s (%d of %d rows kept)" % (
        tag, (t1 - t0) * 1e3, w["n_hits"] / (t1 - t0), st["kept"], st["rows"]))
    dev = eng.upload(pin)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, st = eng.blast_filter(dev, sink=devnull, **kw)
        t1 = time.perf_counter()
    eng.release_text()
    tt = eng.tokenizer_timing()
    print("%s | table resident: %.1f ms = %.3g hits/s; tokenizer launches %s" % (tag, (t1 - t0) * 1e3, w["n_hits"] / (t1 - t0), tt))
import oracle
d = tempfile.mkdtemp()
sample = bench.head_lines(text, 500003)
f = os.path.join(d, "s.last"); open(f, "wb").write(bytes(sample))
n = bytes(sample).count(b"\n")
for tag, kw in sets:
    t0 = time.perf_counter(); oracle.blast_filter(f, f + ".flt", **kw); t1 = time.perf_counter()
    print("CPU oracle (1 thread) %s: %d lines in %.2f s = %.3g hits/s" % (tag, n, t1 - t0, n / (t1 - t0)))
    got, _ = eng.blast_filter(sample, **kw)
    print("  GPU output on the sample byte-identical:", got == open(f + ".flt", "rb").read())

#!/usr/bin/env python
"""The drop-in CLI functions file to file on the C2 table (20 M hits, 1.43 GB): what a user of
`python -m jcvi_b200.compara.blastfilter` / `... synteny scan --liftover` waits for, BED loading, file reads,
Python and output writing included.  Reporting aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from jcvi_b200.compara import blastfilter, synteny

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
w = bench.make_workload(100, hits)
d = w["dir"]
last = os.path.join(d, "A.B.last")
w["text"].tofile(last)
for rep in range(2):
    t0 = time.perf_counter()
    blastfilter.main([last, "--cscore=0.7"])
    t1 = time.perf_counter()
    out = synteny.scan([last + ".filtered", os.path.join(d, "A.B.anchors"), "--liftover=" + last])
    t2 = time.perf_counter()
    print("rep %d: blastfilter.main %.2f s, synteny.scan --liftover %.2f s, total %.2f s = %.3g hits/s file to file (%s: %d bytes)" % (
        rep, t1 - t0, t2 - t1, t2 - t0, w["n_hits"] / (t2 - t0), os.path.basename(out), os.path.getsize(out)), flush=True)
# the same from a blocked-gzip copy of the table (jx_file_upload inflates the blocks on all host threads)
from jcvi_b200 import bgzf
import shutil
dz = os.path.join(d, "z"); os.makedirs(dz, exist_ok=True)
for b in ("A.bed", "B.bed"):
    shutil.copy(os.path.join(d, b), dz)
lastz = os.path.join(dz, "A.B.last.gz")
t0 = time.perf_counter()
csize = bgzf.write_bgzf(w["text"], lastz, level=4)
print("BGZF copy: %d -> %d bytes (%.2fx) written in %.1f s" % (len(w["text"]), csize, len(w["text"]) / csize, time.perf_counter() - t0), flush=True)
for rep in range(2):
    t0 = time.perf_counter()
    blastfilter.main([lastz, "--cscore=0.7", "--qbed=" + os.path.join(dz, "A.bed"), "--sbed=" + os.path.join(dz, "B.bed")])
    t1 = time.perf_counter()
    out = synteny.scan([lastz + ".filtered", os.path.join(dz, "A.B.anchors"), "--liftover=" + lastz, "--qbed=" + os.path.join(dz, "A.bed"), "--sbed=" + os.path.join(dz, "B.bed")])
    t2 = time.perf_counter()
    print("BGZF rep %d: blastfilter.main %.2f s, synteny.scan --liftover %.2f s, total %.2f s = %.3g hits/s file to file" % (
        rep, t1 - t0, t2 - t1, t2 - t0, w["n_hits"] / (t2 - t0)), flush=True)
assert open(lastz + ".filtered", "rb").read() == open(last + ".filtered", "rb").read()
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
blastfilter.main([last, "--cscore=0.7"])
synteny.scan([last + ".filtered", os.path.join(d, "A.B.anchors"), "--liftover=" + last])
pr.disable()
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(28); print(sio.getvalue()[:6000], flush=True)
# the files against the C-ABI path on the in-memory table (the CLI maps large files instead of reading them)
from jcvi_b200.engine import get_engine
eng = get_engine()
f = eng.blastfilter(w["text"], cscore=0.7, want_text=True)
same = bytes(memoryview(f.text)) == open(last + ".filtered", "rb").read()
print(".filtered written by the CLI == C-ABI result on the in-memory table:", same, flush=True)
assert same


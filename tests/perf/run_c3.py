#!/usr/bin/env python
"""BASELINE.json configs[2] at full size on ONE B200: synthetic hexaploid-vs-diploid pair (110 k vs 35 k genes,
three homeologous copies), 200 M hits (~14 GB of text, > 4 GiB offsets), C-score 0.7.  Runs blastfilter -> scan ->
liftover with the table resident, checks the size-independent properties of tests/test_gpu_fullsize.py, compares
the first 500 k lines with the CPU oracle byte for byte, and prints ms per step.  Reporting aid, not bench.py.

    python tests/perf/run_c3.py [hits] [chunk]
"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from jcvi_b200 import synth
from jcvi_b200.formats.bed import Bed
from jcvi_b200.engine import get_engine

hits = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
GQ, GS = 110_000, 35_000
d = tempfile.mkdtemp(prefix="jcvi_b200_c3_")
rng = np.random.default_rng(3)
gq = synth.Genome("W", GQ, 21, rng)
gs = synth.Genome("H", GS, 7, rng)
t0 = time.perf_counter()
parts, n_hits = [], 0
while n_hits < hits:
    arrays = synth.make_hits(gq, gs, min(chunk, hits - n_hits), rng)
    parts.append(synth.format_hits(gq, gs, arrays, header=not parts))
    n_hits += len(arrays[0])
    del arrays
text = np.concatenate(parts) if len(parts) > 1 else parts[0]
del parts
qbed_path, sbed_path = os.path.join(d, "W.bed"), os.path.join(d, "H.bed")
gq.write_bed(qbed_path, np.random.default_rng(10)); gs.write_bed(sbed_path, np.random.default_rng(11))
qbed, sbed = Bed(qbed_path), Bed(sbed_path)
print("generated %d hits, %.2f GB of text in %.1f s" % (n_hits, len(text) / 1e9, time.perf_counter() - t0), flush=True)

eng = get_engine(0)
eng.load_pair(qbed, sbed, False)
t0 = time.perf_counter()
dev = eng.upload(text)
torch.cuda.synchronize()
print("upload (pageable host memory): %.2f s" % (time.perf_counter() - t0), flush=True)

def step():
    r = eng.blastfilter(dev, cscore=0.7, want_text=False, raw=True, skip_arrays=True)
    try:
        s = eng.scan_filtered(r)
    finally:
        eng.free_filter(r)
    blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
    l = eng.liftover_text(dev, s.q_rank, s.s_rank, blk, dist=10)
    return s, l, blk

step()
eng.tokenizer_timing(reset=True)
ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s, l, blk = step()
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
tok_ms, tok_launches, tok_bytes, tok_lines = eng.tokenizer_timing(reset=True)
ms = 1e3 * min(ts)
print("C3 step (table resident): %.2f ms = %.3g hits/s; tokenizer %.2f ms per pass = %.0f GB/s algorithmic (%.1f%% of 6546.6)" % (
    ms, n_hits / (ms / 1e3), tok_ms / 3, (tok_bytes / 3 + 16.0 * tok_lines / 3) / (tok_ms / 3 * 1e-3) / 1e9,
    100 * (tok_bytes / 3 + 16.0 * tok_lines / 3) / (tok_ms / 3 * 1e-3) / 1e9 / 6546.6), flush=True)
print("anchors %d in %d blocks, lifted %d" % (s.n_anchors, s.n_blocks, l.n_lifted), flush=True)

# ---- properties (as tests/test_gpu_fullsize.py) -------------------------------------------------------------
f = eng.blastfilter(dev, cscore=0.7, want_text=True)
n = f.n
sc, off = f.score.astype(np.float64), f.line_off.astype(np.int64)
assert (np.diff(sc) <= 0).all() and (np.diff(off)[np.diff(sc) == 0] > 0).all(), "order"
assert len(text) < (5 << 30) or off.max() > (1 << 32), "a table this large should exercise offsets beyond 4 GiB"
key = (f.q_rank.astype(np.int64) << 32) | f.s_rank.astype(np.int64)
assert len(np.unique(key)) == n and f.q_rank.min() >= 0 and f.q_rank.max() < GQ and f.s_rank.max() < GS, "pairs"
assert (text[off[off > 0] - 1] == 10).all(), "line offsets"
lines = bytes(memoryview(f.text)).split(b"\n")
assert len(lines) == n + 1 and lines[-1] == b""
qa, sa = qbed.tables()["accns"], sbed.tables()["accns"]
for t in list(np.random.default_rng(1).integers(0, n, size=200)) + [0, n - 1]:
    c = lines[t].split(b"\t")
    assert len(c) == 12 and c[0].decode() == qa[f.q_rank[t]] and c[1].decode() == sa[f.s_rank[t]]
    assert float(c[11]) == float("%.3g" % f.score[t])
    raw = bytes(text[off[t]: off[t] + 200]).split(b"\n")[0].split(b"\t")      # the survivor's own input line
    assert float(np.float32(float(raw[11]))) == float(f.score[t])
g = eng.blastfilter(np.frombuffer(bytes(memoryview(f.text)), dtype=np.uint8), cscore=0, tandem_Nmax=0, strip_names=False)
assert g.n == n and np.array_equal(g.q_rank, f.q_rank) and np.array_equal(g.s_rank, f.s_rank), "fixed point"
s2 = eng.scan_text(f.text)
assert np.array_equal(s.q_rank, s2.q_rank) and np.array_equal(s.block_start, s2.block_start), "in-memory hand-over == text"
akey = (s.q_rank.astype(np.int64) << 32) | s.s_rank.astype(np.int64)
assert np.isin(akey, key).all() and len(np.unique(akey)) == s.n_anchors
lkey = (l.q_rank.astype(np.int64) << 32) | l.s_rank.astype(np.int64)
assert l.n_lifted > 0 and not np.isin(lkey, akey).any() and len(np.unique(lkey)) == l.n_lifted and (np.diff(l.block) >= 0).all()
for t in np.random.default_rng(0).integers(0, l.n_lifted, size=200):
    m = blk == l.block[t]
    dd = np.abs(s.q_rank[m] - l.q_rank[t]) + np.abs(s.s_rank[m] - l.s_rank[t])
    assert 0 < dd.min() < 10
print("properties ok: %d filtered lines, offsets up to %.2f GB" % (n, off.max() / 1e9), flush=True)

# ---- the oracle on a prefix ------------------------------------------------------------------------------
import oracle
from jcvi_b200.compara import blastfilter as bf, synteny
sample = bench.head_lines(text, 500_003)
last = os.path.join(d, "W.H.last")
sample.tofile(last)
t0 = time.perf_counter()
oracle.blastfilter(last, qbed_path, sbed_path, out_path=last + ".orc")
oracle.scan(last + ".orc", os.path.join(d, "orc.anchors"), qbed_path, sbed_path)
oracle.liftover(last, os.path.join(d, "orc.anchors"), qbed_path, sbed_path, out_path=os.path.join(d, "orc.lifted"), dist=10)
t_cpu = time.perf_counter() - t0
bf.main([last, "--qbed=" + qbed_path, "--sbed=" + sbed_path])
out = synteny.scan([last + ".filtered", os.path.join(d, "W.H.anchors"), "--qbed=" + qbed_path, "--sbed=" + sbed_path, "--liftover=" + last])
rd = lambda p: open(p, "rb").read()
ok = (rd(last + ".filtered") == rd(last + ".orc") and rd(os.path.join(d, "W.H.anchors")) == rd(os.path.join(d, "orc.anchors"))
      and rd(out) == rd(os.path.join(d, "orc.lifted")))
print("oracle on the first %d lines: %.2f s = %.3g hits/s (1 thread); GPU files byte-identical: %s" % (
    bytes(sample).count(b"\n"), t_cpu, bytes(sample).count(b"\n") / t_cpu, ok), flush=True)
assert ok

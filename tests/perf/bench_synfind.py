#!/usr/bin/env python
"""`compara.synfind` on one B200: the drop-in's C-ABI call on (a) a LAST-like sparse table
(40 k + 40 k genes, about 10 hits per gene) and (b) the dense C2 table (20 M raw hits: windows of several thousand points),
with the CPU oracle on (a) and on a prefix of (b) beside it (tuning / reporting aid, not bench.py)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from jcvi_b200 import synth
from jcvi_b200.engine import get_engine
from jcvi_b200.compara import synfind
from jcvi_b200.compara.synteny import _load_bed

d = tempfile.mkdtemp()
eng = get_engine(0)
cases = [("sparse", 40000, 40000, int(sys.argv[1]) if len(sys.argv) > 1 else 600_000),
         ("dense", 40000, 40000, int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000)]
for tag, Gq, Gs, n in cases:
    sub = os.path.join(d, tag)
    qb, sb, last = synth.write_dataset(sub, "A", "B", 7, Gq, Gs, n, Kq=12, Ks=11)
    qbed, sbed = _load_bed(qb), _load_bed(sb)
    eng.load_pair(qbed, sbed, False)
    text = np.fromfile(last, dtype=np.uint8)
    for rep in range(3):
        t0 = time.perf_counter()
        res = eng.synfind_text(text, window=40, cutoff=0.1)
        t1 = time.perf_counter()
        print("%s rep %d: jx_synfind_text %.1f ms, %d distinct pairs, %d + %d regions, %d launches so far" % (
            tag, rep, (t1 - t0) * 1e3, res.n_points, res.n_forward, res.n_mirror, eng.launch_count()))
    t0 = time.perf_counter()
    synfind.main([last, "--qbed=" + qb, "--sbed=" + sb, "-o", last + ".gpu"])
    t1 = time.perf_counter()
    print("%s: drop-in main() file to file %.2f s" % (tag, t1 - t0))
    if tag == "sparse" or os.environ.get("ORACLE_DENSE"):
        t0 = time.perf_counter()
        oracle.synfind(last, qb, sb, last + ".orc")
        t1 = time.perf_counter()
        same = open(last + ".gpu", "rb").read() == open(last + ".orc", "rb").read()
        print("%s: CPU oracle (1 thread) %.2f s; outputs identical: %s" % (tag, t1 - t0, same))

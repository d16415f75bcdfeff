#!/usr/bin/env python
"""Dense-window stress of the scan kernels: `synteny scan` on point sets where a point's xdist window
holds hundreds to thousands of earlier points (a --cscore=0 table of a polyploid, dist=40) next to the usual sparse case.
Prints the JX_TRACE phase times of jx_scan_points.  Reporting aid.
    JX_TRACE=1 python tests/perf/scan_dense.py [n_points] [genes] [chromosomes]"""
import os, sys, time
os.environ.setdefault("JX_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from jcvi_b200.engine import get_engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 40_000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 12
rng = np.random.default_rng(5)
eng = get_engine(0)
per = G // K
for tag, npts, dist in (("sparse", max(n // 16, 1000), 20), ("dense", n, 40)):
    # uniform noise over the K x K chromosome pairs + collinear runs along the diagonals (the part that survives min_size)
    nd = npts // 10
    q = rng.integers(0, G, npts - nd); s = rng.integers(0, G, npts - nd)
    dq = rng.integers(0, G, nd); ds = np.clip(dq + rng.integers(-3, 4, nd), 0, G - 1)
    q = np.concatenate([q, dq]); s = np.concatenate([s, ds])
    key = np.unique(q.astype(np.int64) * G + s)
    q = (key // G).astype(np.int32); s = (key % G).astype(np.int32)
    grp = (q // per).astype(np.int64) * (K + 1) + (s // per)
    sc = np.ones(len(q), dtype=np.float32)
    for rep in range(2):
        print("---- %s: %d points, dist %d, rep %d" % (tag, len(q), dist, rep), file=sys.stderr, flush=True)
        t0 = time.perf_counter()
        res = eng.scan_points(q, s, sc, grp, xdist=dist, ydist=dist, N=4)
        t1 = time.perf_counter()
        print("%s: %d points -> %d anchors in %d blocks, %.1f ms (whole call, host arrays in and out)" % (
            tag, len(q), res.n_anchors, res.n_blocks, (t1 - t0) * 1e3), flush=True)

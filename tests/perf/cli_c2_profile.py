#!/usr/bin/env python
"""cProfile of the second `synteny.scan --liftover` call of tests/perf/cli_c2.py (where the CLI time goes)."""
import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from jcvi_b200.compara import blastfilter, synteny
w = bench.make_workload(100, int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000)
d = w["dir"]; last = os.path.join(d, "A.B.last"); w["text"].tofile(last)
blastfilter.main([last, "--cscore=0.7"])
synteny.scan([last + ".filtered", os.path.join(d, "A.B.anchors"), "--liftover=" + last])
pr = cProfile.Profile(); pr.enable()
blastfilter.main([last, "--cscore=0.7"])
synteny.scan([last + ".filtered", os.path.join(d, "A.B.anchors"), "--liftover=" + last])
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)

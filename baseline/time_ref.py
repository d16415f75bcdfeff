#!/usr/bin/env python
"""Time the reference's own path on one table: blastfilter.main -> synteny.scan --liftover (file to file), the way
`catalog.ortholog` runs them (src/jcvi/compara/catalog.py:720-750).  Run as a subprocess of bench.py with the built
baseline/_ref on sys.path; prints one JSON line {"blastfilter": s, "scan_liftover": s, "total": s, "lines": n}."""
import json
import logging
import os
import sys
import time

last, qbed, sbed = sys.argv[1:4]
dist = sys.argv[4] if len(sys.argv) > 4 else "10"
sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "shims"),
                os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")]
import jcvi.compara.blastfilter as bf
import jcvi.compara.synteny as syn
from jcvi.formats.blast import BlastLine
assert BlastLine.__module__ == "jcvi.formats.cblast", BlastLine.__module__      # the compiled parser, not pyblast
logging.getLogger("jcvi").setLevel(logging.ERROR)
logging.disable(logging.WARNING)
n = sum(1 for ln in open(last, "rb") if not ln.startswith(b"#"))
stem = last[:-len(".last")] if last.endswith(".last") else last
beds = ["--qbed=" + qbed, "--sbed=" + sbed]
t0 = time.perf_counter()
bf.main([last] + beds)
t1 = time.perf_counter()
err = os.dup(2); dn = os.open(os.devnull, os.O_WRONLY); os.dup2(dn, 2)        # summary() prints to stderr
try:
    syn.scan([last + ".filtered", stem + ".ref.anchors"] + beds + ["--liftover=" + last, "--liftover_dist=" + dist])
finally:
    os.dup2(err, 2)
t2 = time.perf_counter()
print(json.dumps({"blastfilter": t1 - t0, "scan_liftover": t2 - t1, "total": t2 - t0, "lines": n}))

#!/usr/bin/env python
"""Time the tokenizer kernel of several builds of libjcvi_b200.so on the same synthetic table.

    python tools/tok_variants.py [--hits N] lib1.so lib2.so ...

One subprocess per library (JCVI_B200_LIB selects it); prints ms per 64 MiB launch as measured by the
library's own CUDA events (jx_tokenizer_timing).  Tuning aid, not a benchmark.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def child(hits):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from jcvi_b200.engine import get_engine
    w = bench.make_workload(100, hits)
    eng = get_engine(0)
    eng.load_pair(w["qbed"], w["sbed"], False)
    pin = torch.from_numpy(w["text"]).pin_memory()
    dev = eng.upload(pin)
    torch.cuda.synchronize()
    for it in range(8):
        if it == 3:
            eng.tokenizer_timing(reset=True)
        try:
            r = eng.blastfilter(dev, want_text=False, raw=True, skip_arrays=True)
            eng.free_filter(r)
        except ZeroDivisionError:          # experiment builds that skip the best-score table
            pass
    ms, launches, nbytes, lines = eng.tokenizer_timing(reset=True)
    full = nbytes / 5 // (64 << 20)          # launches over a full 64 MiB chunk per pass
    print("RESULT lib=%s ms_total_per_pass=%.4f launches_per_pass=%d MBps=%.1f ms_per_64MiB=%.4f" % (
        os.environ.get("JCVI_B200_LIB", "default"), ms / 5, launches // 5, nbytes / 1e6 / (ms / 1e3) / 1e3,
        (ms / 5) / (nbytes / 5 / (64 << 20))), flush=True)


def main():
    args = sys.argv[1:]
    hits = 6000000
    if args and args[0] == "--hits":
        hits = int(args[1]); args = args[2:]
    if args and args[0] == "--child":
        return child(int(args[1]))
    for lib in args or [""]:
        env = dict(os.environ)
        if lib:
            env["JCVI_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(hits)], env=env, check=False)


if __name__ == "__main__":
    main()

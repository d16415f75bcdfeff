#!/usr/bin/env python
"""One genome pair sharded by chromosome pair over the ranks of a torchrun launch (BASELINE configs[2] / [4] shapes):

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/run_shard_cp.py [--hits H] [--config c3|c5] [--verify] [--steps K]

The table is the concatenation of 8 seeded chunks; rank r of N generates (only) chunks [8 r / N, 8 (r + 1) / N), so the
table is the same for every N.  A step = jcvi_b200.sharded.run_sharded over the resident slices; time = max over ranks
between barriers.  --verify: rank 0 also builds the whole table, runs the single-GPU pipeline and compares every array.
Prints one JSON line on rank 0.  Reporting aid (bench.py carries the same measurement as `strong_scaling`)."""
import argparse, hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--hits", type=int, default=200_000_000)
ap.add_argument("--config", default="c3")
ap.add_argument("--verify", action="store_true")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--trace", action="store_true")
ap.add_argument("--profile", action="store_true", help="cProfile of rank 0 around the timed steps (host-side cost of the driver)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
acc = {}
seen = [0]
if args.profile:
    # wall time spent inside every driver-level call (engine stage calls, collectives, synchronisations): where the host
    # of rank 0 blocks.  No extra synchronisation is added.
    from jcvi_b200 import sharded as _sh
    from jcvi_b200.engine import Engine as _Eng

    def _wrap(obj, name, label):
        f = getattr(obj, name)

        def g(*a, **k):
            if label.startswith("run_sharded"):
                seen[0] += 1
                if seen[0] == 4:                      # warm-up steps (allocations, NCCL set-up) are not counted
                    acc.clear()
            t0 = time.perf_counter()
            try:
                return f(*a, **k)
            finally:
                e = acc.setdefault(label, [0, 0.0]); e[0] += 1; e[1] += time.perf_counter() - t0
        setattr(obj, name, g)
    for n in ("allreduce_max_", "allgather_fixed", "alltoallv", "allgatherv"):
        _wrap(_sh.DistComm, n, "comm." + n)
    for n in ("cscore_begin", "shard_candidates", "shard_partition", "shard_pairs", "shard_survivors", "scan_cands", "shard_near",
              "shard_lift", "shard_reset", "best_table_tensor", "load_pair", "shard_slots"):
        _wrap(_Eng, n, "engine." + n)
    _wrap(torch.cuda, "synchronize", "torch.cuda.synchronize")
    _wrap(_sh, "run_sharded", "run_sharded (total)")
out = bench.strong_scaling(args.config, args.hits, args.steps, 2, world, rank, local, verify=args.verify, trace=args.trace)
if args.profile and rank == 0:
    calls = max(acc.get("run_sharded (total)", [1])[0], 1)
    for k, (n, t) in sorted(acc.items(), key=lambda x: -x[1][1]):
        print("%-32s %5.1f calls/step %8.3f ms/step" % (k, n / calls, t / calls * 1e3), file=sys.stderr, flush=True)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

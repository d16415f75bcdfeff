#!/usr/bin/env python
"""BASELINE configs[3] as a JOB LIST: pair files of mixed size (one `.last` per genome pair of a synthetic pangenome)
dealt to the ranks by jcvi_b200.parallel.run_pairs (greedy size balancing, no data-path collective), each job the
drop-in CLI path file to file: blastfilter.main -> synteny.scan --liftover (src/jcvi/compara/catalog.py:720-750).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/run_c4.py [--genomes 12] [--hits-total 120000000]

Rank r generates the files j with j % N == r into /dev/shm; after a barrier every rank runs its jobs.  Rank 0 prints one
JSON line: hits/s, per-rank busy seconds, balance efficiency (mean busy / max busy)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from jcvi_b200 import synth, parallel

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=12)
ap.add_argument("--hits-total", type=int, default=120_000_000, dest="hits_total")
ap.add_argument("--genes", type=int, default=25_000)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["JCVI_B200_DEVICE"] = str(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
shm = "/dev/shm/jcvi_b200_c4"
os.makedirs(shm, exist_ok=True)
G = args.genomes
pairs = [(a, b) for a in range(G) for b in range(a + 1, G)]
rng = np.random.default_rng(44)
w = rng.uniform(0.2, 3.0, size=len(pairs)); w /= w.sum()
hits = np.maximum((w * args.hits_total).astype(np.int64), 100_000)          # mixed sizes, 15x spread
genomes = []
grng = np.random.default_rng(4)
for g in range(G):
    genomes.append(synth.Genome("P%02d" % g, args.genes + int(grng.integers(-3000, 3000)), int(grng.integers(8, 20)), grng))
t0 = time.perf_counter()
for g in range(G):
    if g % world == rank:
        genomes[g].write_bed(os.path.join(shm, "P%02d.bed" % g), np.random.default_rng(100 + g))
jobs = []
for j, (a, b) in enumerate(pairs):
    last = os.path.join(shm, "P%02d.P%02d.last" % (a, b))
    jobs.append((last, os.path.join(shm, "P%02d.bed" % a), os.path.join(shm, "P%02d.bed" % b)))
    if j % world == rank:
        arrays = synth.make_hits(genomes[a], genomes[b], int(hits[j]), np.random.default_rng(1000 + j))
        synth.format_hits(genomes[a], genomes[b], arrays, nthreads=max(2, 32 // world)).tofile(last)
gen_s = time.perf_counter() - t0
if world > 1:
    dist.barrier()
# warm-up: CUDA context, library, one small job per rank
from jcvi_b200.compara import blastfilter, synteny
small = min(range(len(jobs)), key=lambda j: hits[j])
blastfilter.main([jobs[small][0], "--qbed=" + jobs[small][1], "--sbed=" + jobs[small][2]])
import bench
sampler = bench.ClockSampler(local)
sampler.start()
t0 = time.perf_counter()
mine, outs, wall = parallel.run_pairs(jobs, liftover_dist=10)
busy = parallel.run_pairs.last_busy
clocks = sampler.summary()
tb = torch.tensor([busy], dtype=torch.float64, device="cuda")
allb = [torch.zeros_like(tb) for _ in range(world)]
if world > 1:
    dist.all_gather(allb, tb)
else:
    allb = [tb]
lifted = sum(open(p, "rb").read().count(b"\n") for p in outs.values())
tl = torch.tensor([lifted], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(tl)
if rank == 0:
    b = [float(x[0]) for x in allb]
    tot = int(hits.sum())
    print(json.dumps({"workload": "C4: synthetic pangenome, %d genomes, %d pair files of mixed size (%.1f M .. %.1f M hits), drop-in CLI file to file"
                      % (G, len(jobs), hits.min() / 1e6, hits.max() / 1e6), "n_gpus": world, "jobs": len(jobs), "hits": tot,
                      "wall_s": wall, "value": tot / wall, "unit": "hits/s", "busy_s_per_rank": [round(x, 3) for x in b],
                      "balance_efficiency": round(sum(b) / len(b) / max(b), 4), "lifted_anchor_lines": int(tl[0]),
                      "generation_s_rank0": round(gen_s, 1), "clocks": clocks}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
for f in os.listdir(shm):
    if rank == 0:
        try:
            os.remove(os.path.join(shm, f))
        except OSError:
            pass

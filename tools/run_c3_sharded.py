#!/usr/bin/env python
"""BASELINE.json configs[2] ("200 M hits, C-score 0.7, 1/2/4 B200"): ONE genome pair sharded over the ranks
(jcvi_b200/sharded.py: byte-range slices cut at line boundaries, ncclAllReduce(max) of the per-name best-score
table, NCCL gathers of the surviving lines), slices resident in HBM, strong scaling.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/run_c3_sharded.py [hits]

Rank 0 generates the table once into /dev/shm; every rank maps it and uploads only its slice.  A step =
blastfilter -> scan -> liftover of the whole pair; time = max over ranks between barriers.  After timing, rank 0
runs the unsharded path on the whole table and compares every output array.  Reporting aid, not bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from jcvi_b200 import synth
from jcvi_b200.formats.bed import Bed
from jcvi_b200.engine import get_engine
from jcvi_b200.sharded import slice_bounds, _gather_bytes, _gather_device

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hits = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
chunk = 20_000_000
GQ, GS = 110_000, 35_000
shm = "/dev/shm/jcvi_b200_c3"
os.makedirs(shm, exist_ok=True)
last, qbed_path, sbed_path = os.path.join(shm, "W.H.last"), os.path.join(shm, "W.bed"), os.path.join(shm, "H.bed")
n_hits = 0
if rank == 0:
    t0 = time.perf_counter()
    rng = np.random.default_rng(3)
    gq = synth.Genome("W", GQ, 21, rng)
    gs = synth.Genome("H", GS, 7, rng)
    with open(last, "wb") as fw:
        first = True
        while n_hits < hits:
            arrays = synth.make_hits(gq, gs, min(chunk, hits - n_hits), rng)
            synth.format_hits(gq, gs, arrays, header=first).tofile(fw)
            n_hits += len(arrays[0]); first = False
            del arrays
    gq.write_bed(qbed_path, np.random.default_rng(10)); gs.write_bed(sbed_path, np.random.default_rng(11))
    print("generated %d hits in %.1f s" % (n_hits, time.perf_counter() - t0), flush=True)
nh = torch.tensor([n_hits], dtype=torch.int64, device="cuda")
dist.all_reduce(nh); n_hits = int(nh.item())
qbed, sbed = Bed(qbed_path), Bed(sbed_path)
text = np.memmap(last, dtype=np.uint8, mode="r")
lo, hi = slice_bounds(text, rank, world)
mine = np.array(text[lo:hi])                         # this rank's slice only
eng = get_engine(local)
eng.load_pair(qbed, sbed, False)
dev = eng.upload(mine)
torch.cuda.synchronize()
print("rank %d: slice [%d, %d) = %.2f GB resident" % (rank, lo, hi, (hi - lo) / 1e9), flush=True)


marks = []


def mark(label):
    if os.environ.get("C3_TRACE"):
        torch.cuda.synchronize(); marks.append((label, time.perf_counter()))


def step():
    del marks[:]
    mark("start")
    ptr, n = eng.cscore_begin(dev, strip_names=True)
    mark("cscore_begin (tokenize slice)")
    if world > 1:
        best = eng.best_table_tensor(ptr, n)
        dist.all_reduce(best, op=dist.ReduceOp.MAX)   # ncclAllReduce(max) over NVLink, in place
        torch.cuda.synchronize()
    mark("all-reduce max")
    if world > 1:
        lines, _ = eng.cscore_finish(0.7, device=True)       # survivors stay on the device ...
        merged = _gather_device(lines, device=local)         # ... and travel GPU to GPU
    else:
        merged, _ = eng.cscore_finish(0.7)
    mark("cscore_finish + gather")
    a = [None, None, None]
    f = s = None
    if rank == 0:
        f = eng.blastfilter(merged, cscore=0, tandem_Nmax=10, want_text=True)
        mark("rank 0: blastfilter on survivors")
        s = eng.scan_text(f.text)
        blk = np.repeat(np.arange(s.n_blocks, dtype=np.int32), np.diff(s.block_start))
        a = [s.q_rank, s.s_rank, blk]
    if world > 1:
        sizes = torch.tensor([0 if a[0] is None else len(a[0])], dtype=torch.int64, device="cuda")
        dist.broadcast(sizes, src=0)
        na = int(sizes.item())
        buf = torch.empty((3, na), dtype=torch.int32, device="cuda")
        if rank == 0:
            buf.copy_(torch.from_numpy(np.stack(a)).cuda())
        dist.broadcast(buf, src=0)                    # anchors to every rank, GPU to GPU
        a = list(buf.cpu().numpy())
    mark("scan + anchors broadcast")
    if world > 1:
        near, _ = eng.near_lines(dev, a[0], a[1], strip_names=True, dist=10, device=True)   # re-uses the slice's records
        merged = _gather_device(near, device=local)
    else:
        merged, _ = eng.near_lines(dev, a[0], a[1], strip_names=True, dist=10)
    mark("near_lines + gather")
    l = None
    if rank == 0:
        l = eng.liftover_text(merged, a[0], a[1], a[2], dist=10)
    mark("rank 0: liftover on near lines")
    return f, s, l, a


step()
ts = []
for _ in range(3):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    f, s, l, a = step()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
ms = 1e3 * min(ts)
if os.environ.get("C3_TRACE"):
    print("rank %d phases (ms): " % rank + "; ".join("%s %.2f" % (marks[i][0], 1e3 * (marks[i][1] - marks[i - 1][1])) for i in range(1, len(marks))), flush=True)
if rank == 0:
    print("C3 sharded over %d GPU(s), slices resident: %.2f ms per step = %.3g hits/s (filtered %d, anchors %d in %d blocks, lifted %d)" % (
        world, ms, n_hits / (ms / 1e3), f.n, s.n_anchors, s.n_blocks, l.n_lifted), flush=True)
    # ---- the unsharded path on the whole table, same GPU ---------------------------------------------------
    whole = eng.upload(np.asarray(text))
    t0 = time.perf_counter()
    f1 = eng.blastfilter(whole, cscore=0.7, want_text=True)
    s1 = eng.scan_text(f1.text)
    blk1 = np.repeat(np.arange(s1.n_blocks, dtype=np.int32), np.diff(s1.block_start))
    l1 = eng.liftover_text(whole, s1.q_rank, s1.s_rank, blk1, dist=10)
    torch.cuda.synchronize()
    same = (bytes(memoryview(f.text)) == bytes(memoryview(f1.text)) and np.array_equal(s.q_rank, s1.q_rank)
            and np.array_equal(s.s_rank, s1.s_rank) and np.array_equal(s.block_start, s1.block_start) and np.array_equal(s.score, s1.score)
            and all(np.array_equal(getattr(l, k), getattr(l1, k)) for k in ("q_rank", "s_rank", "block", "accepted", "score")))
    print("identical to the single-GPU outputs (.filtered bytes, anchors, lifted pairs): %s" % same, flush=True)
    assert same
dist.barrier()
if rank == 0:
    for p in (last, qbed_path, sbed_path):
        os.remove(p)
dist.destroy_process_group()

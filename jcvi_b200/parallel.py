"""Multi-GPU sharding of the path: genome pairs are independent jobs
(blastfilter's C-score tables are per `.last` file), so they are dealt to ranks by greedy size
balancing and need NO data-path collective.  torch.distributed is only used for the barrier /
max-over-ranks timing and for gathering per-rank summaries."""
import os


def assign_pairs(sizes, world):
    """Longest-processing-time greedy: -> list (per rank) of job indices; deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    return out


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def run_pairs(jobs, cscore=0.7, tandem_Nmax=10, dist=20, min_size=4, liftover_dist=None, backend=None):
    """Run blastfilter -> scan(--liftover) for every job = (last_path, qbed_path, sbed_path) owned by
    this rank (one process per GPU, LOCAL_RANK selects the device).  Returns
    (my job indices, {job index: lifted anchors path}, max-over-ranks seconds)."""
    import time
    import torch
    import torch.distributed as dist_
    from .compara import blastfilter, synteny

    rank, world = rank_world()
    sizes = [os.path.getsize(j[0]) for j in jobs]
    mine = assign_pairs(sizes, world)[rank]
    if world > 1 and not dist_.is_initialized():
        dist_.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"))
    if world > 1:
        dist_.barrier()
    t0 = time.perf_counter()
    outs = {}
    for i in mine:
        last, qbed, sbed = jobs[i]
        beds = ["--qbed=" + qbed, "--sbed=" + sbed]
        blastfilter.main([last] + beds + ["--cscore=%g" % cscore, "--tandem_Nmax=%d" % tandem_Nmax])
        anchors = last.rsplit(".", 1)[0] + ".anchors"
        args = [last + ".filtered", anchors] + beds + ["--dist=%d" % dist, "--min_size=%d" % min_size, "--liftover=" + last]
        if liftover_dist:
            args.append("--liftover_dist=%d" % liftover_dist)
        outs[i] = synteny.scan(args)
    dt = time.perf_counter() - t0
    run_pairs.last_busy = dt                     # this rank's own busy time (the return value is the max over ranks)
    return mine, outs, max_over_ranks(dt)


def max_over_ranks(value):
    import torch
    import torch.distributed as dist_
    if not dist_.is_available() or not dist_.is_initialized():
        return value
    dev = "cuda" if dist_.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
    return float(t[0])

"""One genome pair sharded over several GPUs (BASELINE config C3).

Every rank tokenizes a byte range of the raw table cut at line boundaries.  blastfilter needs one
global quantity, filter_cscore's per-name best score (src/jcvi/compara/blastfilter.py:227-232):
the rank-local tables are max-reduced in place with an NCCL all-reduce over NVLink.  Each rank then
keeps the lines of its slice that pass the C-score filter (~5 % of the table); their concatenation
in rank order is in file order, and the remaining stages (pair de-duplication, tandem collapse,
ordering, writer) on rank 0 see exactly what they would have seen unsharded.  liftover is sharded
the same way with the "within dist of an anchor" predicate.  Outputs are byte-identical to the
single-GPU path (tests/test_gpu_sharded.py; tools/check_sharded.py on 2 GPUs).
"""
import numpy as np

from .engine import get_engine
from .parallel import rank_world


def slice_bounds(text, rank, world):
    """[lo, hi) of rank's slice; cuts fall just after a '\\n' (lines never straddle ranks)."""
    n = len(text)

    def cut(r):
        if r <= 0:
            return 0
        if r >= world:
            return n
        p = n * r // world
        nl = np.flatnonzero(text[p:p + (1 << 20)] == 10)
        while len(nl) == 0 and p + (1 << 20) < n:
            p += 1 << 20
            nl = np.flatnonzero(text[p:p + (1 << 20)] == 10)
        return n if len(nl) == 0 else p + int(nl[0]) + 1

    return cut(rank), cut(rank + 1)


def _gather_bytes(arr, dst=0, device=None):
    """Gather variable-length uint8 arrays on rank `dst` (list in rank order; None elsewhere).
    With the NCCL backend the payload travels GPU to GPU over NVLink."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    if world == 1 or not dist.is_initialized():
        return [arr]
    if dist.get_backend() != "nccl":
        out = [None] * world if rank == dst else None
        dist.gather_object(arr, out, dst=dst)
        return out
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = len(arr)
    dist.all_reduce(sizes)
    sizes = sizes.tolist()
    pad = max(max(sizes), 1)
    mine = torch.zeros(pad, dtype=torch.uint8, device=dev)
    if len(arr):
        mine[: len(arr)] = torch.from_numpy(np.ascontiguousarray(arr)).to(dev, non_blocking=True)
    bufs = [torch.empty(pad, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(mine, bufs, dst=dst)
    if rank != dst:
        return None
    return [bufs[r][: sizes[r]].cpu().numpy() for r in range(world)]


def _gather_device(t, dst=0, device=None):
    """Gather variable-length CUDA uint8 tensors on rank `dst` into ONE contiguous CUDA tensor (rank order) with
    point-to-point NCCL transfers straight into their final place: the lines never touch host memory.
    Returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device())
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = t.numel()
    dist.all_reduce(sizes)
    sizes = sizes.tolist()
    ops, out = [], None
    if rank == dst:
        out = torch.empty(max(sum(sizes), 1), dtype=torch.uint8, device=dev)
        off = 0
        for r in range(world):
            if r == rank:
                out[off:off + sizes[r]].copy_(t)
            elif sizes[r]:
                ops.append(dist.P2POp(dist.irecv, out[off:off + sizes[r]], r))
            off += sizes[r]
    elif t.numel():
        ops.append(dist.P2POp(dist.isend, t, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    torch.cuda.current_stream(dev).synchronize()      # the library reads `out` on its own stream next
    return out[: sum(sizes)] if rank == dst else None


def _use_device_exchange():
    import torch.distributed as dist
    _, world = rank_world()
    return world > 1 and dist.is_initialized() and dist.get_backend() == "nccl"


def _as_u8(text):
    """numpy view of the text; torch CPU tensors (pinned or not) are viewed without a copy."""
    if hasattr(text, "numpy") and hasattr(text, "is_cuda") and not text.is_cuda:
        return text.numpy().view(np.uint8).reshape(-1), text
    return np.asarray(text).view(np.uint8).reshape(-1), None


def _slice_for_engine(text_np, owner, lo, hi):
    """The rank's slice as something Engine accepts; a slice of a pinned torch tensor stays pinned
    (full PCIe speed for jx_text_upload)."""
    if owner is not None:
        return owner.view(-1)[lo:hi] if owner.dtype == __import__("torch").uint8 else text_np[lo:hi]
    return text_np[lo:hi]


def blastfilter_sharded(text, qbed, sbed, is_self, strip_names=True, cscore=0.7, tandem_Nmax=10, exclude_keys=None,
                        want_text=True, engine=None):
    """Collective over all ranks; returns the FilterResult on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    rank, world = rank_world()
    eng = engine or get_engine()
    eng.load_pair(qbed, sbed, is_self)
    text, owner = _as_u8(text)
    lo, hi = slice_bounds(text, rank, world)
    ptr, n = eng.cscore_begin(_slice_for_engine(text, owner, lo, hi), strip_names=strip_names, exclude_keys=exclude_keys)
    if world > 1 and dist.is_initialized():
        best = eng.best_table_tensor(ptr, n)
        dist.all_reduce(best, op=dist.ReduceOp.MAX)          # ncclAllReduce(max) over NVLink, in place
        torch.cuda.synchronize(eng.device)
    if _use_device_exchange():                               # survivors go GPU to GPU over NVLink
        lines, cnt = eng.cscore_finish(cscore, device=True)
        merged = _gather_device(lines, device=eng.device)
        if rank != 0:
            return None
    else:
        lines, cnt = eng.cscore_finish(cscore)
        parts = _gather_bytes(lines, device=eng.device)
        if rank != 0:
            return None
        merged = np.concatenate(parts) if len(parts) > 1 else parts[0]
    # survivors only: the C-score filter is done, exclusions were applied while tokenizing
    return eng.blastfilter(merged, strip_names=strip_names, cscore=0, tandem_Nmax=tandem_Nmax, want_text=want_text)


def liftover_sharded(text, a_qi, a_si, a_block, qbed, sbed, is_self, strip_names=True, dist=10, engine=None):
    """Collective; LiftoverResult on rank 0 (None elsewhere)."""
    rank, world = rank_world()
    eng = engine or get_engine()
    eng.load_pair(qbed, sbed, is_self)
    text, owner = _as_u8(text)
    lo, hi = slice_bounds(text, rank, world)
    if _use_device_exchange():
        lines, cnt = eng.near_lines(_slice_for_engine(text, owner, lo, hi), a_qi, a_si, strip_names=strip_names, dist=dist, device=True)
        merged = _gather_device(lines, device=eng.device)
        if rank != 0:
            return None
    else:
        lines, cnt = eng.near_lines(_slice_for_engine(text, owner, lo, hi), a_qi, a_si, strip_names=strip_names, dist=dist)
        parts = _gather_bytes(lines, device=eng.device)
        if rank != 0:
            return None
        merged = np.concatenate(parts) if len(parts) > 1 else parts[0]
    return eng.liftover_text(merged, a_qi, a_si, a_block, strip_names=strip_names, dist=dist)


# =====================================================================================================================
# Chromosome-pair sharding (SURVEY.md 8(e) stages B / C): every stage after the C-score predicate runs on the rank that
# OWNS the chromosome pair.  Data path collectives: all-reduce(max) of the best-score table, all-to-all of 16-byte
# candidates (twice: blastfilter survivors-to-be, liftover near-hits), all-gather of tandem edges and of the anchors.
# The same driver runs over torch.distributed (one rank per process, NCCL) and over LocalComm (all ranks in one process
# on one GPU: how the GPU tests exercise the whole exchange logic without a second device).
# =====================================================================================================================
class DistComm(object):
    """One rank per process, torch.distributed (NCCL on GPUs)."""

    def __init__(self):
        import torch.distributed as dist
        self.rank, self.world = rank_world()
        self.local = [self.rank]
        self.dist = dist if self.world > 1 and dist.is_initialized() else None

    def allreduce_max_(self, tensors):
        if self.dist:
            self.dist.all_reduce(tensors[0], op=self.dist.ReduceOp.MAX)

    def allreduce_sum_np(self, arrays):
        import torch
        if not self.dist:
            return [arrays[0].copy()]
        t = torch.from_numpy(arrays[0].astype(np.int64)).to(self._dev())
        self.dist.all_reduce(t)
        return [t.cpu().numpy()]

    def allgather_np(self, arrays):
        """small host arrays: -> [list over all ranks] per local rank"""
        if not self.dist:
            return [[arrays[0]]]
        out = [None] * self.world
        self.dist.all_gather_object(out, arrays[0])
        return [out]

    def _dev(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def alltoallv(self, sends, counts):
        """sends[0]: [n, W] tensor grouped by destination, counts[0]: int64[world] -> (received [m, W] tensor, recv counts)"""
        import torch
        if not self.dist:
            return [sends[0]], [counts[0].copy()]
        c = torch.from_numpy(np.ascontiguousarray(counts[0], dtype=np.int64)).to(sends[0].device)
        rc = torch.empty_like(c)
        self.dist.all_to_all_single(rc, c)
        rcounts = rc.cpu().numpy()
        recv = torch.empty((int(rcounts.sum()), sends[0].shape[1]), dtype=sends[0].dtype, device=sends[0].device)
        self.dist.all_to_all_single(recv, sends[0].contiguous(), output_split_sizes=[int(x) for x in rcounts],
                                    input_split_sizes=[int(x) for x in counts[0]])
        return [recv], [rcounts]

    def allgatherv(self, tensors):
        """[n_r, W] device tensors -> the concatenation over all ranks (rank order) on every rank, + sizes"""
        import torch
        t = tensors[0]
        if not self.dist:
            return [t], [np.array([t.shape[0]], dtype=np.int64)]
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = torch.empty(self.world, dtype=torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(sizes, n)
        sizes = sizes.cpu().numpy()
        pad = int(max(sizes.max(), 1))
        mine = torch.zeros((pad, t.shape[1]), dtype=t.dtype, device=t.device)
        mine[: t.shape[0]] = t
        allp = torch.empty((self.world * pad, t.shape[1]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(allp, mine)
        parts = [allp[r * pad: r * pad + int(sizes[r])] for r in range(self.world)]
        return [torch.cat(parts, 0) if parts else t], [sizes]

    def barrier(self):
        if self.dist:
            self.dist.barrier()


class LocalComm(object):
    """All ranks in this process (engines on one device): the collectives are plain tensor / array operations."""

    def __init__(self, world):
        self.world = world
        self.rank = 0
        self.local = list(range(world))
        self.dist = None

    def allreduce_max_(self, tensors):
        import torch
        m = tensors[0].clone()
        for t in tensors[1:]:
            m = torch.maximum(m, t)
        for t in tensors:
            t.copy_(m)

    def allreduce_sum_np(self, arrays):
        s = np.sum([a.astype(np.int64) for a in arrays], axis=0)
        return [s.copy() for _ in arrays]

    def allgather_np(self, arrays):
        return [list(arrays) for _ in arrays]

    def alltoallv(self, sends, counts):
        import torch
        W = self.world
        offs = [np.concatenate([[0], np.cumsum(c)]) for c in counts]
        recv, rcounts = [], []
        for d in range(W):
            parts = [sends[r][int(offs[r][d]): int(offs[r][d + 1])] for r in range(W)]
            recv.append(torch.cat(parts, 0).clone())
            rcounts.append(np.array([int(counts[r][d]) for r in range(W)], dtype=np.int64))
        return recv, rcounts

    def allgatherv(self, tensors):
        import torch
        cat = torch.cat(list(tensors), 0)
        sizes = np.array([t.shape[0] for t in tensors], dtype=np.int64)
        return [cat.clone() for _ in tensors], [sizes.copy() for _ in tensors]

    def barrier(self):
        pass


def assign_owners(hist, world):
    """Chromosome pair -> owning rank: longest-processing-time greedy on the global candidate histogram (deterministic:
    every rank computes the same table)."""
    order = sorted(range(len(hist)), key=lambda k: (-int(hist[k]), k))
    load = [0] * world
    owner = np.zeros(len(hist), dtype=np.int32)
    for k in order:
        r = min(range(world), key=lambda x: (load[x], x))
        owner[k] = r
        load[r] += int(hist[k])
    return owner


class ShardedResult(object):
    """What rank 0 holds after a sharded run: the `.filtered` columns (output order), the anchors block by block in the
    reference's order, the lifted pairs with `accepted` / `raw_score` per anchor line -- the same arrays as the
    single-GPU FilterResult / ScanResult / LiftoverResult."""

    def __init__(self):
        self.filtered = None
        self.scan = None
        self.lifted = None
        self.timing = {}


class _Scan(object):
    pass


class _Lift(object):
    pass


def run_sharded(slices, qbed, sbed, is_self=False, strip_names=True, cscore=0.7, tandem_Nmax=10, xdist=20, ydist=20, N=4,
                intrabound=300, liftover_dist=10, comm=None, engines=None, want_filtered=True, mark=None):
    """blastfilter -> scan -> liftover of ONE genome pair whose raw table is cut into `slices` (one per LOCAL rank: a list
    of length 1 under torch.distributed, of length world under LocalComm; each slice ends at a line boundary; host arrays,
    pinned tensors or DeviceText).  Returns a ShardedResult (complete on rank 0)."""
    import torch
    comm = comm or DistComm()
    W = comm.world
    engines = engines or [get_engine()]
    assert len(engines) == len(comm.local) == len(slices)
    if is_self:
        raise NotImplementedError("chromosome-pair sharding of a self comparison (pairs are re-oriented after the exchange)")
    mark = mark or (lambda name: None)
    L = range(len(engines))
    tq = qbed.tables()
    ts = sbed.tables()
    n_chr_s = int(ts["n_chr"])
    n_cp = int(tq["n_chr"]) * n_chr_s
    qchr = np.asarray(tq["chr_lex"]); schr = np.asarray(ts["chr_lex"])
    for i in L:
        engines[i].load_pair(qbed, sbed, False)
        engines[i].shard_reset()
    # ---- stage A: tokenize the slice, best-score table, all-reduce(max) -------------------------------------------------
    bests = []
    for i in L:
        ptr, n_ids = engines[i].cscore_begin(slices[i], strip_names=strip_names)
        bests.append(engines[i].best_table_tensor(ptr, n_ids))
    mark("tokenize")
    torch.cuda.synchronize()
    comm.allreduce_max_(bests)
    torch.cuda.synchronize()
    slots = [np.array([engines[i].shard_slots()], dtype=np.int64) for i in L]
    all_slots = comm.allgather_np(slots)
    base = []
    for k, i in enumerate(L):
        sizes = [int(x[0]) for x in all_slots[k]]
        r = comm.local[k]
        base.append(int(sum(sizes[:r])))
        total_slots = int(sum(sizes))
    mark("allreduce")
    # ---- stage B: candidates -> owners of their chromosome pair ------------------------------------------------------------
    cands, hists = [], []
    for i in L:
        c, h = engines[i].shard_candidates(cscore, base[i], n_cp)
        cands.append(c); hists.append(h)
    ghist = comm.allreduce_sum_np(hists)
    owner = assign_owners(ghist[0], W)
    parts, counts = [], []
    for i in L:
        p, c = engines[i].shard_partition(cands[i], owner, W)
        parts.append(p); counts.append(c)
    torch.cuda.synchronize()
    recv, _ = comm.alltoallv(parts, counts)
    torch.cuda.synchronize()
    mark("select+exchange")
    # ---- stage C: per owner -- pair best, tandem edges; all-gather the edges; mothers; survivors ------------------------------
    qe, se = [], []
    for i in L:
        a, b = engines[i].shard_pairs(recv[i], tandem_Nmax)
        qe.append(a); se.append(b)
    if tandem_Nmax:
        qall, _ = comm.allgatherv(qe)
        sall, _ = comm.allgatherv(se)
    else:
        qall, sall = qe, se
    torch.cuda.synchronize()
    surv = []
    for i in L:
        s_, mq, ms = engines[i].shard_survivors(qall[i].contiguous(), sall[i].contiguous(), tandem_Nmax)
        surv.append(s_)
    mark("pairs+tandem")
    res = ShardedResult()
    if want_filtered:
        allsurv, _ = comm.allgatherv(surv)
        torch.cuda.synchronize()
        if comm.rank == 0:
            res.filtered = engines[0].shard_order(allsurv[0].contiguous())
        mark("filtered order")
    # ---- scan per owner; the blocks of all owners in sorted chromosome-pair order --------------------------------------------
    scans = [engines[i].scan_cands(surv[i], xdist=xdist, ydist=ydist, N=N, intrabound=intrabound) for i in L]
    mark("scan")
    anchors_dev, meta = [], []
    for i in L:
        s_ = scans[i]
        first = s_.block_start[:-1]
        bcp = (qchr[s_.q_rank[first]].astype(np.int64) * n_chr_s + schr[s_.s_rank[first]]) if s_.n_blocks else np.zeros(0, np.int64)
        meta.append(np.stack([bcp, np.diff(s_.block_start)]) if s_.n_blocks else np.zeros((2, 0), np.int64))
        a = np.stack([s_.q_rank, s_.s_rank, s_.score.view(np.int32)], axis=1) if s_.n_anchors else np.zeros((0, 3), np.int32)
        anchors_dev.append(torch.from_numpy(np.ascontiguousarray(a)).to("cuda:%d" % engines[i].device))
    all_meta = comm.allgather_np(meta)
    all_anchors, _ = comm.allgatherv(anchors_dev)
    torch.cuda.synchronize()
    # global block order: sorted chromosome pair (each pair lives on one rank), then the rank's own cluster order
    m = all_meta[0]
    cps = np.concatenate([x[0] for x in m]) if m else np.zeros(0, np.int64)
    sizes = np.concatenate([x[1] for x in m]) if m else np.zeros(0, np.int64)
    order = np.argsort(cps, kind="stable")
    starts = np.concatenate([[0], np.cumsum(sizes)])[:-1]
    A = all_anchors[0].cpu().numpy()
    nb = len(order)
    new_sizes = sizes[order]
    new_start = np.concatenate([[0], np.cumsum(new_sizes)]).astype(np.int64)
    if nb:
        idx = np.concatenate([np.arange(starts[b], starts[b] + sizes[b]) for b in order]) if A.shape[0] else np.zeros(0, np.int64)
    else:
        idx = np.zeros(0, np.int64)
    G = A[idx] if len(idx) else np.zeros((0, 3), np.int32)
    sc = _Scan()
    sc.q_rank = np.ascontiguousarray(G[:, 0]); sc.s_rank = np.ascontiguousarray(G[:, 1])
    sc.score = np.ascontiguousarray(G[:, 2]).view(np.float32)
    sc.block_start = new_start; sc.n_blocks = nb; sc.n_anchors = int(G.shape[0])
    res.scan = sc
    a_block = np.repeat(np.arange(nb, dtype=np.int32), new_sizes) if nb else np.zeros(0, np.int32)
    mark("anchors gathered")
    # ---- liftover: every rank tests its raw records against all anchors, near hits go to the owners ---------------------------
    near = [engines[i].shard_near(sc.q_rank, sc.s_rank, base[i], strip_names=strip_names, dist=liftover_dist) for i in L]
    parts, counts = [], []
    for i in L:
        p, c = engines[i].shard_partition(near[i], owner, W)
        parts.append(p); counts.append(c)
    torch.cuda.synchronize()
    recv2, _ = comm.alltoallv(parts, counts)
    torch.cuda.synchronize()
    mark("near+exchange")
    lifts = [engines[i].shard_lift(recv2[i].contiguous(), total_slots, sc.q_rank, sc.s_rank, a_block, dist=liftover_dist) for i in L]
    mark("lift")
    # lifted pairs and the per-anchor columns to rank 0: a block's lifted pairs all come from its owner
    lt = []
    for i in L:
        l = lifts[i]
        t = np.stack([l.q_rank, l.s_rank, l.score.view(np.int32), l.block], axis=1) if l.n_lifted else np.zeros((0, 4), np.int32)
        acc = np.stack([l.accepted.astype(np.int32), l.raw_score.view(np.int32)], axis=1) if len(l.accepted) else np.zeros((0, 2), np.int32)
        lt.append((torch.from_numpy(np.ascontiguousarray(t)).to("cuda:%d" % engines[i].device),
                   torch.from_numpy(np.ascontiguousarray(acc)).to("cuda:%d" % engines[i].device)))
    all_l, _ = comm.allgatherv([x[0] for x in lt])
    all_a, asz = comm.allgatherv([x[1] for x in lt])
    torch.cuda.synchronize()
    Lh = all_l[0].cpu().numpy()
    o = np.argsort(Lh[:, 3], kind="stable") if Lh.shape[0] else np.zeros(0, np.int64)
    Lh = Lh[o]
    li = _Lift()
    li.q_rank = np.ascontiguousarray(Lh[:, 0]); li.s_rank = np.ascontiguousarray(Lh[:, 1])
    li.score = np.ascontiguousarray(Lh[:, 2]).view(np.float32); li.block = np.ascontiguousarray(Lh[:, 3])
    li.n_lifted = int(Lh.shape[0])
    Ah = all_a[0].cpu().numpy().reshape(W, -1, 2) if sc.n_anchors else np.zeros((W, 0, 2), np.int32)
    accm = Ah[:, :, 0]
    li.accepted = (accm.max(axis=0) if accm.size else np.zeros(0, np.int32)).astype(np.uint8)
    who = accm.argmax(axis=0) if accm.size else np.zeros(0, np.int64)
    raw = Ah[who, np.arange(Ah.shape[1]), 1] if accm.size else np.zeros(0, np.int32)
    li.raw_score = np.ascontiguousarray(raw).view(np.float32)
    res.lifted = li
    mark("lifted gathered")
    return res

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


_STATUS = ((NotImplementedError, 1), (ZeroDivisionError, 2), (MemoryError, 3))


def _status_code(exc):
    if exc is None:
        return 0
    for cls, code in _STATUS:
        if isinstance(exc, cls):
            return code
    return 4


def _agree(err, device=None):
    """All ranks learn whether any rank failed (one tiny all-reduce) and stop together: without it a rank-local
    NotImplementedError / ZeroDivisionError would leave the other ranks waiting in the next collective."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if err is not None:
            raise err
        return
    nccl = dist.get_backend() == "nccl"
    t = torch.tensor([_status_code(err)], dtype=torch.int32, device=("cuda:%d" % device) if nccl else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    code = int(t.item())
    if code:
        if err is not None:
            raise err
        cls = {c: k for k, c in _STATUS}.get(code, RuntimeError)
        raise cls("another rank failed in its slice of the table (see its log); all ranks stop together")


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
    err = None
    try:
        ptr, n = eng.cscore_begin(_slice_for_engine(text, owner, lo, hi), strip_names=strip_names, exclude_keys=exclude_keys)
    except Exception as e:                                   # noqa: BLE001 -- agreed on below
        err = e
    _agree(err, eng.device)
    if world > 1 and dist.is_initialized():
        best = eng.best_table_tensor(ptr, n)
        dist.all_reduce(best, op=dist.ReduceOp.MAX)          # ncclAllReduce(max) over NVLink, in place
        torch.cuda.synchronize(eng.device)
    dev_x = _use_device_exchange()
    try:
        lines, cnt = eng.cscore_finish(cscore, device=True) if dev_x else eng.cscore_finish(cscore)
    except Exception as e:                                   # noqa: BLE001 -- ZeroDivisionError on one rank only
        err = e
    _agree(err, eng.device)
    if dev_x:                                                # survivors go GPU to GPU over NVLink
        merged = _gather_device(lines, device=eng.device)
        if rank != 0:
            return None
    else:
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
    dev_x = _use_device_exchange()
    err = None
    try:
        lines, cnt = eng.near_lines(_slice_for_engine(text, owner, lo, hi), a_qi, a_si, strip_names=strip_names, dist=dist,
                                    device=dev_x)
    except Exception as e:                                   # noqa: BLE001 -- agreed on below
        err = e
    _agree(err, eng.device)
    if dev_x:
        merged = _gather_device(lines, device=eng.device)
        if rank != 0:
            return None
    else:
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

    def _dev(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def allreduce_max_(self, tensors):
        if self.dist:
            self.dist.all_reduce(tensors[0], op=self.dist.ReduceOp.MAX)

    def allgather_fixed(self, arrays):
        """equal-shape small host arrays -> [world, ...] int64 numpy (the same on every rank), per local rank"""
        import torch
        a = np.ascontiguousarray(arrays[0], dtype=np.int64)
        if not self.dist:
            return [a[None]]
        t = torch.from_numpy(a).to(self._dev())
        out = torch.empty((self.world,) + tuple(a.shape), dtype=torch.int64, device=t.device)
        self.dist.all_gather_into_tensor(out, t.unsqueeze(0) if a.ndim == 0 else t)
        return [out.cpu().numpy().reshape((self.world,) + tuple(a.shape))]

    def alltoallv(self, sends, matrix):
        """sends[0]: [n, W] device tensor grouped by destination; matrix[r][d] = rows rank r sends to rank d (known to
        every rank) -> the rows received, grouped by source"""
        import torch
        if not self.dist:
            return [sends[0]]
        r = self.rank
        recv = torch.empty((int(matrix[:, r].sum()), sends[0].shape[1]), dtype=sends[0].dtype, device=sends[0].device)
        self.dist.all_to_all_single(recv, sends[0].contiguous(), output_split_sizes=[int(x) for x in matrix[:, r]],
                                    input_split_sizes=[int(x) for x in matrix[r, :]])
        return [recv]

    def allgatherv_known(self, tensors, sizes):
        """allgatherv when every rank already knows every rank's row count (no size exchange)"""
        import torch
        t = tensors[0]
        if not self.dist:
            return [t]
        sizes = np.asarray(sizes, dtype=np.int64)
        pad = int(max(sizes.max(), 1))
        mine = torch.zeros((pad, t.shape[1]), dtype=t.dtype, device=t.device)
        mine[: t.shape[0]] = t
        allp = torch.empty((self.world * pad, t.shape[1]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(allp, mine)
        if (sizes == pad).all():
            return [allp]
        return [torch.cat([allp[r * pad: r * pad + int(sizes[r])] for r in range(self.world)], 0)]

    def allgatherv(self, tensors):
        """[n_r, W] device tensors -> (their concatenation over all ranks in rank order, sizes[world]) on every rank"""
        import torch
        t = tensors[0]
        if not self.dist:
            return [t], np.array([t.shape[0]], dtype=np.int64)
        sizes = self.allgather_fixed([np.array([t.shape[0]])])[0][:, 0]
        pad = int(max(sizes.max(), 1))
        mine = torch.zeros((pad, t.shape[1]), dtype=t.dtype, device=t.device)
        mine[: t.shape[0]] = t
        allp = torch.empty((self.world * pad, t.shape[1]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(allp, mine)
        if (sizes == pad).all():
            return [allp], sizes
        return [torch.cat([allp[r * pad: r * pad + int(sizes[r])] for r in range(self.world)], 0)], sizes

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

    def allgather_fixed(self, arrays):
        m = np.stack([np.ascontiguousarray(a, dtype=np.int64) for a in arrays])
        return [m.copy() for _ in arrays]

    def alltoallv(self, sends, matrix):
        import torch
        W = self.world
        offs = np.concatenate([np.zeros((W, 1), np.int64), np.cumsum(matrix, axis=1)], axis=1)
        return [torch.cat([sends[r][int(offs[r][d]): int(offs[r][d + 1])] for r in range(W)], 0).clone() for d in range(W)]

    def allgatherv(self, tensors):
        import torch
        cat = torch.cat(list(tensors), 0)
        return [cat.clone() for _ in tensors], np.array([t.shape[0] for t in tensors], dtype=np.int64)

    def allgatherv_known(self, tensors, sizes):
        return self.allgatherv(tensors)[0]

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


def _send_matrix(H, owner, world):
    """H[r][cp] = candidates of rank r per chromosome pair -> matrix[r][d] = rows rank r sends to rank d"""
    M = np.zeros((H.shape[0], world), dtype=np.int64)
    for d in range(world):
        M[:, d] = H[:, owner == d].sum(axis=1)
    return M


class ShardedResult(object):
    """What rank 0 holds after a sharded run: the `.filtered` columns (output order), the anchors block by block in the
    reference's order, the lifted pairs with `accepted` / `raw_score` per anchor line -- the same arrays as the
    single-GPU FilterResult / ScanResult / LiftoverResult."""

    def __init__(self):
        self.filtered = None
        self.scan = None
        self.lifted = None
        self.mode = "owners"           # how the stages behind the C-score predicate were run (run_sharded's `mode`)
        self.n_candidates = 0          # records that passed the C-score predicate, over all ranks


class _Scan(object):
    pass


class _Lift(object):
    pass


def _raise_agreed(codes, failed):
    """codes: one status per rank (0 = fine).  Every rank raises the exception class of the lowest failing rank; a rank
    that failed itself re-raises its own exception (it carries the message)."""
    codes = [int(c) for c in np.asarray(codes).reshape(-1)]
    bad = [r for r, c in enumerate(codes) if c]
    if not bad:
        return
    for e in failed:
        if e is not None:
            raise e
    cls = {code: c for c, code in _STATUS}.get(codes[bad[0]], RuntimeError)
    raise cls("rank %d failed in its slice of the table (see its log); all ranks stop together" % bad[0])


# candidates in total up to which every rank takes all of them ("auto" mode of run_sharded).  Measured on 8 B200s: C3
# (200 M hits, <= 1.6 M candidates) gather 6.1 ms vs owners 7.8 ms; C5 (500 M hits, ~14 M candidates) owners 13.9 ms vs gather
# 18.0 ms -- the stages behind the predicate cost ~0.65 ms per million candidates on one GPU, the owners' extra exchanges ~3 ms
GATHER_MAX = 4_000_000


def _run_gathered(n_candidates, comm, engines, L, W, cands, sizes, base, total_slots, strip_names, tandem_Nmax, xdist, ydist, N, intrabound,
                  liftover_dist, want_filtered, mark):
    """run_sharded, mode "gather" (from stage B on; `sizes`: candidates per rank, known from the histogram all-gather)."""
    import torch
    allc = comm.allgatherv_known(cands, sizes)              # every rank: all candidates, in rank order
    torch.cuda.synchronize()
    mark("select+exchange")
    rows, survs = [], []
    for i in L:
        a, b = engines[i].shard_pairs(allc[i], tandem_Nmax)
        surv = engines[i].shard_survivors(a, b, tandem_Nmax)[0]
        survs.append(surv)
        rows.append(engines[i].scan_cands(surv, xdist=xdist, ydist=ydist, N=N, intrabound=intrabound, device=True))   # {q, s, score, block}
    mark("pairs+tandem+scan")
    res = ShardedResult()
    res.mode = "gather"
    res.n_candidates = n_candidates
    anchors = []
    for i in L:
        r = rows[i]
        anchors.append((r[:, 0].contiguous(), r[:, 1].contiguous(), r[:, 3].contiguous()))
    # ---- liftover: every rank tests its raw records against the anchors; the near hits are gathered ---------------------------
    near = [engines[i].shard_near(anchors[i][0], anchors[i][1], base[i], strip_names=strip_names, dist=liftover_dist) for i in L]
    alln, _ = comm.allgatherv(near)
    torch.cuda.synchronize()
    mark("near+exchange")
    if comm.rank == 0:
        if want_filtered:
            res.filtered = engines[0].shard_order(survs[0])
        R = rows[0].cpu().numpy()
        sc = _Scan()
        sc.q_rank = np.ascontiguousarray(R[:, 0]); sc.s_rank = np.ascontiguousarray(R[:, 1])
        sc.score = np.ascontiguousarray(R[:, 2]).view(np.float32)
        blk = R[:, 3]
        hs = np.flatnonzero(np.concatenate([[1], blk[1:] != blk[:-1]])) if len(blk) else np.zeros(0, np.int64)
        sc.block_start = np.concatenate([hs, [len(blk)]]).astype(np.int64)
        sc.n_blocks = len(hs); sc.n_anchors = int(len(blk))
        res.scan = sc
        res.lifted = engines[0].shard_lift(alln[0].contiguous(), total_slots, anchors[0][0], anchors[0][1], anchors[0][2],
                                           dist=liftover_dist, device=False)
    mark("lift")
    return res


def run_sharded(slices, qbed, sbed, is_self=False, strip_names=True, cscore=0.7, tandem_Nmax=10, xdist=20, ydist=20, N=4,
                intrabound=300, liftover_dist=10, comm=None, engines=None, want_filtered=True, mark=None, mode="auto"):
    """blastfilter -> scan -> liftover of ONE genome pair whose raw table is cut into `slices` (one per LOCAL rank: a list
    of length 1 under torch.distributed, of length world under LocalComm; each slice ends at a line boundary; host arrays,
    pinned tensors or DeviceText).  Returns a ShardedResult (complete on rank 0).

    mode: what happens after the C-score predicate (the part that is O(raw hits) -- tokenizing, the predicate, the
    liftover prefilter -- is sharded either way):
      "owners"  every chromosome pair has an owner rank; all-to-all of the candidates, every later stage per owner;
      "gather"  the candidates (a few % of the raw hits) are all-gathered and EVERY rank runs pair de-duplication, tandem
                collapse and scan on all of them (identical results, nothing to exchange afterwards); the raw records near
                an anchor are gathered and lifted on rank 0.  The stages behind the predicate are latency-sized on such a
                set (2 - 3 ms on one GPU whatever the split), so splitting them buys nothing and every exchange costs;
      "auto"    "gather" up to GATHER_MAX candidates in total, "owners" beyond (large or unfiltered tables)."""
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
    devs = ["cuda:%d" % e.device for e in engines]
    for i in L:
        engines[i].load_pair(qbed, sbed, False)
        engines[i].shard_reset()
    # every rank owns a fixed stride of the global order (no exchange of slice sizes)
    stride = (1 << 31) // W
    base = [comm.local[i] * stride for i in L]
    total_slots = (1 << 31) - 1
    # ---- stage A: tokenize the slice, best-score table, all-reduce(max) -------------------------------------------------
    # A failure that only one rank sees (an unsupported line in its slice, ZeroDivisionError from a score <= 0, an
    # allocation) must not leave the others waiting in a collective: the rank keeps taking part with empty data, its
    # status travels with the histogram all-gather below, and every rank raises the same exception after it.
    bests = []
    failed = [None] * len(engines)
    for i in L:
        try:
            ptr, n_ids = engines[i].cscore_begin(slices[i], strip_names=strip_names)
            bests.append(engines[i].best_table_tensor(ptr, n_ids))
            if engines[i].shard_slots() >= stride:
                raise NotImplementedError("a slice with more than 2^31 / world record slots")
        except Exception as e:                               # noqa: BLE001 -- re-raised on every rank below
            failed[i] = e
            bests.append(torch.zeros(max(engines[i].n_name_ids, 1), dtype=torch.int32, device=devs[i]))
    mark("tokenize")
    comm.allreduce_max_(bests)
    torch.cuda.synchronize()
    mark("allreduce")
    # ---- stage B: candidates -> owners of their chromosome pair ------------------------------------------------------------
    cands, hists = [], []
    for i in L:
        c = h = None
        if failed[i] is None:
            try:
                c, h = engines[i].shard_candidates(cscore, base[i], n_cp)
            except Exception as e:                           # noqa: BLE001
                failed[i] = e
        if failed[i] is not None:
            c, h = torch.zeros((0, 4), dtype=torch.int32, device=devs[i]), np.zeros(n_cp, dtype=np.uint32)
        cands.append(c)
        hists.append(np.concatenate([np.asarray(h, dtype=np.uint32), [_status_code(failed[i])]]).astype(np.uint32))
    H = comm.allgather_fixed(hists)[0]                      # [world, n_cp + 1]: every rank's send counts + its status
    _raise_agreed(H[:, n_cp], failed)
    H = np.ascontiguousarray(H[:, :n_cp])
    hists = [h[:n_cp] for h in hists]
    if mode == "gather" or (mode == "auto" and int(H.sum()) <= GATHER_MAX):
        return _run_gathered(int(H.sum()), comm, engines, L, W, cands, H.sum(axis=1), base, total_slots, strip_names, tandem_Nmax, xdist, ydist, N,
                             intrabound, liftover_dist, want_filtered, mark)
    owner = assign_owners(H.sum(axis=0), W)
    M = _send_matrix(H, owner, W)
    parts = [engines[i].shard_partition(cands[i], owner, W, hist_in=hists[i])[0] for i in L]
    recv = comm.alltoallv(parts, M)
    torch.cuda.synchronize()
    mark("select+exchange")
    # ---- stage C: per owner -- pair best, tandem edges; all-gather the edges; mothers; survivors ------------------------------
    edges, nq = [], []
    for i in L:
        a, b = engines[i].shard_pairs(recv[i], tandem_Nmax)
        edges.append(torch.cat([a, b], 0) if tandem_Nmax else a[:0]); nq.append(np.array([a.shape[0], b.shape[0]]))
    surv = []
    if tandem_Nmax:
        E = comm.allgather_fixed(nq)[0]                     # [world, 2]
        eall, esz = comm.allgatherv(edges)
        torch.cuda.synchronize()
        offs = np.concatenate([[0], np.cumsum(esz)])
        for i in L:
            qs = torch.cat([eall[i][int(offs[r]): int(offs[r]) + int(E[r, 0])] for r in range(W)], 0).contiguous()
            ss = torch.cat([eall[i][int(offs[r]) + int(E[r, 0]): int(offs[r + 1])] for r in range(W)], 0).contiguous()
            surv.append(engines[i].shard_survivors(qs, ss, tandem_Nmax)[0])
    else:
        for i in L:
            z = edges[i]
            surv.append(engines[i].shard_survivors(z, z, 0)[0])
    mark("pairs+tandem")
    res = ShardedResult()
    res.n_candidates = int(H.sum())
    if want_filtered:
        allsurv, _ = comm.allgatherv(surv)
        torch.cuda.synchronize()
        if comm.rank == 0:
            res.filtered = engines[0].shard_order(allsurv[0].contiguous())
        mark("filtered order")
    # ---- scan per owner; the blocks of all owners in sorted chromosome-pair order (assembled on the device) -----------------
    per = []
    for i in L:
        d = devs[i]
        rows = engines[i].scan_cands(surv[i], xdist=xdist, ydist=ydist, N=N, intrabound=intrabound, device=True)   # {q, s, score, block}
        if not hasattr(engines[i], "_chr_t") or engines[i]._chr_t[0] is not qbed:
            engines[i]._chr_t = (qbed, torch.from_numpy(np.asarray(tq["chr_lex"], dtype=np.int32)).to(d),
                                 torch.from_numpy(np.asarray(ts["chr_lex"], dtype=np.int32)).to(d))
        _, qc, sc_ = engines[i]._chr_t
        if rows.shape[0]:
            cp = qc[rows[:, 0].long()] * n_chr_s + sc_[rows[:, 1].long()]
            per.append(torch.cat([rows[:, :3], cp.unsqueeze(1), rows[:, 3:4]], dim=1))
        else:
            per.append(torch.zeros((0, 5), dtype=torch.int32, device=d))
    mark("scan")
    allA, _ = comm.allgatherv(per)
    anchors = []
    for i in L:
        A = allA[i]
        if A.shape[0]:
            # every chromosome pair lives on one rank and a rank's anchors are ordered (pair, cluster, member): a stable
            # sort by (pair, cluster) of the concatenation is the reference's block order
            key = A[:, 3].long() * (int(A[:, 4].max()) + 1) + A[:, 4].long()
            ks, o = torch.sort(key, stable=True)
            A = A[o]
            head = torch.ones(A.shape[0], dtype=torch.int32, device=A.device)
            head[1:] = (ks[1:] != ks[:-1]).to(torch.int32)
            blk = (torch.cumsum(head, 0) - 1).to(torch.int32)
        else:
            head = blk = A[:, 0]
        anchors.append((A[:, 0].contiguous(), A[:, 1].contiguous(), blk.contiguous(), A[:, 2].contiguous(), head))
    if comm.rank == 0:
        aq, as_, ab, asb, head = anchors[0]
        sc = _Scan()
        sc.q_rank = aq.cpu().numpy(); sc.s_rank = as_.cpu().numpy(); sc.score = asb.cpu().numpy().view(np.float32)
        hs = np.flatnonzero(head.cpu().numpy()) if aq.shape[0] else np.zeros(0, np.int64)
        sc.block_start = np.concatenate([hs, [aq.shape[0]]]).astype(np.int64)
        sc.n_blocks = len(hs); sc.n_anchors = int(aq.shape[0])
        res.scan = sc
    mark("anchors gathered")
    # ---- liftover: every rank tests its raw records against all anchors, near hits go to the owners ---------------------------
    parts, cnts = [], []
    for i in L:
        near = engines[i].shard_near(anchors[i][0], anchors[i][1], base[i], strip_names=strip_names, dist=liftover_dist)
        p, c = engines[i].shard_partition(near, owner, W)
        parts.append(p); cnts.append(c)
    M2 = comm.allgather_fixed(cnts)[0]
    recv2 = comm.alltoallv(parts, M2)
    torch.cuda.synchronize()
    mark("near+exchange")
    lt, accs = [], []
    for i in L:
        cols, acc = engines[i].shard_lift(recv2[i].contiguous(), total_slots, anchors[i][0], anchors[i][1], anchors[i][2],
                                          dist=liftover_dist, device=True)
        lt.append(cols.t().contiguous())
        # accepted / raw score of an anchor line come from the one rank that owns its pair: max over ranks of (flag, bits)
        accs.append(acc.clone())
    mark("lift")
    all_l, _ = comm.allgatherv(lt)
    comm.allreduce_max_(accs)
    torch.cuda.synchronize()
    if comm.rank == 0:
        Lt = all_l[0]
        if Lt.shape[0]:                                    # a block's lifted pairs all come from its owner, already in order
            Lt = Lt[torch.sort(Lt[:, 3], stable=True)[1]]
        Lh = Lt.cpu().numpy()
        li = _Lift()
        li.q_rank = np.ascontiguousarray(Lh[:, 0]); li.s_rank = np.ascontiguousarray(Lh[:, 1])
        li.score = np.ascontiguousarray(Lh[:, 2]).view(np.float32); li.block = np.ascontiguousarray(Lh[:, 3])
        li.n_lifted = int(Lh.shape[0])
        ah = accs[0].cpu().numpy()
        li.accepted = (ah >> 32).astype(np.uint8)
        li.raw_score = (ah & 0xFFFFFFFF).astype(np.uint32).view(np.float32)
        res.lifted = li
    mark("lifted gathered")
    return res

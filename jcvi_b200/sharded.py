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

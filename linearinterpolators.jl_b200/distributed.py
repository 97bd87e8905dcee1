"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch; gloo for CPU tests).

The interpolation path shards over independent units -- the non-interpolated (post) dimension, batch images
(BASELINE config #3) or observed frames (config #5) -- so forward and per-image adjoint need no communication.  The
one exchange step is the adjoint *accumulated into a shared model* during iterative fitting, ``m = sum_k A_k' y_k``:
every rank accumulates its frames into a private model-sized buffer and a single sum-allreduce combines them.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of ``range(n)`` over ``world`` ranks: rank r owns [lo, hi).  The first
    ``n % world`` ranks get one extra unit, so the blocks differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    q, r = divmod(int(n), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def world_info():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:  # pragma: no cover
        pass
    return 0, 1


def _dense(t):
    """A contiguous view of the same memory (collectives need one): column-major arrays are permuted views."""
    if t.is_contiguous():
        return t
    tt = t.permute(*reversed(range(t.ndim)))
    if not tt.is_contiguous():
        raise ValueError("accumulator must be a dense row- or column-major tensor")
    return tt


def accumulate_adjoint(nframes: int, adjoint_into: Callable[[int, object], None], acc, *, groups: int = 1,
                       group=None, comm_stream=None):
    """``acc <- sum over ALL ranks and frames k of A_k' y_k``.

    ``adjoint_into(k, acc)`` must add frame k's adjoint into ``acc`` in place (e.g.
    ``apply_b(1, Adjoint, A_k, y_k, False, 1, acc)``); ``acc`` (a torch tensor, zero-initialised by the caller or
    holding the prior model) is the rank-private accumulator.  Frames are block-sharded over the ranks and visited
    in ascending order, so the result is run-to-run deterministic for a given world size.

    ``groups > 1`` splits the local frames into that many groups accumulated into separate buffers; the
    all-reduce of group g runs on ``comm_stream`` while group g+1 is being computed (overlap), and the groups are
    added up at the end.  ``groups == 1`` is a single all-reduce after all local frames.
    """
    import torch
    import torch.distributed as dist
    rank, world = world_info()
    lo, hi = shard_range(nframes, rank, world)
    if groups <= 1 or world == 1:
        for k in range(lo, hi):
            adjoint_into(k, acc)
        if world > 1:
            dist.all_reduce(_dense(acc), op=dist.ReduceOp.SUM, group=group)
        return acc
    # overlapped variant
    n_local = hi - lo
    bounds = [lo + (n_local * g) // groups for g in range(groups + 1)]
    parts, works = [], []
    on_gpu = acc.is_cuda
    cur = torch.cuda.current_stream(acc.device) if on_gpu else None
    comm = comm_stream if comm_stream is not None else (torch.cuda.Stream(acc.device) if on_gpu else None)
    for g in range(groups):
        part = acc if g == 0 else torch.zeros_like(acc)
        for k in range(bounds[g], bounds[g + 1]):
            adjoint_into(k, part)
        if on_gpu:
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                works.append(dist.all_reduce(_dense(part), op=dist.ReduceOp.SUM, group=group, async_op=True))
        else:
            works.append(dist.all_reduce(_dense(part), op=dist.ReduceOp.SUM, group=group, async_op=True))
        parts.append(part)
    for w in works:
        w.wait()
    if on_gpu:
        cur.wait_stream(comm)
    for part in parts[1:]:
        acc.add_(part)
    return acc

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


def bind_host_to_device(device_index: int) -> dict:
    """Pins the calling process to the CPU cores NVML reports as local to GPU ``device_index`` (its NUMA node), so that
    pinned host buffers allocated afterwards are first-touched on that node and the host-buffer entry points
    (``li_*_host``) do not cross the inter-socket link: with one process per GPU on a two-socket box, eight ranks that
    all stage through one node's memory share its bandwidth.  Returns what was done (for the bench line); never raises.
    """
    import os
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES renumbers devices: go through the PCI bus id of the CUDA device
        handle = None
        try:
            import torch
            bus = torch.cuda.get_device_properties(device_index)
            pci = f"{getattr(bus, 'pci_domain_id', 0):08x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(pci.encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"bound": True, "cpus": len(cpus), "first_cpu": min(cpus), "allowed_before": len(allowed)}
    except Exception as exc:  # pragma: no cover - depends on the box
        info["why"] = f"{type(exc).__name__}: {exc}"[:120]
    return info


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


# ---- the same exchange step behind the C ABI (NCCL inside libli_b200.so) ----------------------------------------
class Comm:
    """``li_comm``: an NCCL communicator owned by the library, one rank per GPU.  Nothing here needs torch; a Julia
    or C host does the same three calls (``li_comm_unique_id`` on rank 0, hand the 128 bytes to every rank,
    ``li_comm_init_rank`` on each)."""

    def __init__(self, handle: int):
        import ctypes as C
        from . import _lib
        self.handle = handle
        r, n, d = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib.li_comm_info(handle, C.byref(r), C.byref(n), C.byref(d)))
        self.rank, self.nranks, self.device = int(r.value), int(n.value), int(d.value)

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _lib
        buf = (C.c_char * _lib.COMM_ID_BYTES)()
        _lib.check(_lib.lib.li_comm_unique_id(buf, _lib.COMM_ID_BYTES))
        return bytes(buf)

    @classmethod
    def init_rank(cls, uid: bytes, nranks: int, rank: int) -> "Comm":
        """Collective: every rank calls it with the id made by rank 0; binds to the current CUDA device."""
        import ctypes as C
        from . import _lib
        if len(uid) != _lib.COMM_ID_BYTES:
            raise ValueError("the communicator id must be LI_COMM_ID_BYTES long")
        h = C.c_void_p()
        _lib.check(_lib.lib.li_comm_init_rank(uid, int(nranks), int(rank), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_torch_distributed(cls, group=None) -> Optional["Comm"]:
        """One library communicator over the ranks of an initialised ``torch.distributed`` group (which only carries
        the 128-byte id, over its own backend); None when there is a single process."""
        import torch
        import torch.distributed as dist
        rank, world = world_info()
        if world == 1:
            return None
        ids = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, group=group)
        return cls.init_rank(ids[0], world, rank)

    def allreduce_sum(self, x):
        """In-place sum of a device array over the ranks (``li_allreduce_sum``)."""
        import numpy as np
        from . import _arrays as ar
        from . import _lib
        n = int(np.prod(ar.shape_of(x), dtype=np.int64))
        _lib.check(_lib.lib.li_allreduce_sum(self.handle, ar.li_dtype(ar.np_dtype_of(x)), ar.data_ptr(x), n,
                                             ar.current_stream(x)))
        return x

    def destroy(self):
        from . import _lib
        if getattr(self, "handle", None):
            _lib.lib.li_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:  # pragma: no cover
            pass


class FrameBatch:
    """The local frames of an iterative fit, checked and marshalled ONCE: the table handles and frame pointers of
    ``li_adjoint_accumulate_allreduce`` as C arrays.  ``adjoint_accumulate(batch, model)`` then costs one library call
    per iteration instead of a Python loop over the frames (512 frames: 1.5 ms of argument checks per call, as long as
    the device work itself).  The frames and tables must stay alive and in place while the batch is used."""

    def __init__(self, tabs_list, srcs):
        import ctypes as C
        from . import _arrays as ar
        from . import _lib
        n = len(tabs_list)
        if len(srcs) != n:
            raise ValueError("one source array per frame")
        self.n = n
        self.dtype = None
        self.handles = (C.c_void_p * max(2 * n, 1))()
        self.ptrs = (C.c_void_p * max(n, 1))()
        self._keep = (list(tabs_list), list(srcs))
        self.stream = None
        for k, (tabs, x) in enumerate(zip(tabs_list, srcs)):
            if len(tabs) != 2 or tabs[0] is None or tabs[1] is None:
                raise _lib.ErrorException("every frame needs one table per dimension")
            if not ar.is_device(x) or not ar.is_colmajor(x):
                raise _lib.ErrorException("frames must be column-major device arrays")
            dt = ar.np_dtype_of(x)
            if self.dtype is None:
                self.dtype = dt
                self.stream = ar.current_stream(x)
            elif dt != self.dtype:
                raise _lib.ErrorException("frames must share one element type")
            if ar.shape_of(x) != (tabs[0].nrows, tabs[1].nrows):
                raise _lib.DimensionMismatch(f"frame {k} has size {ar.shape_of(x)}, expected {(tabs[0].nrows, tabs[1].nrows)}")
            self.handles[2 * k], self.handles[2 * k + 1] = tabs[0].handle, tabs[1].handle
            self.ptrs[k] = ar.data_ptr(x)

    def __len__(self):
        return self.n


def adjoint_accumulate(tabs_list, srcs=None, model=None, alpha=1.0, beta=0.0, comm: Optional[Comm] = None, nbands: int = 1):
    """``model <- beta*model + alpha * sum over all ranks and local frames k of (A1_k*A2_k)' * y_k`` in one library
    call (``li_adjoint_accumulate_allreduce``): the multi-frame accumulate kernel on the local frames, band by band,
    with the NCCL all-reduce of each band overlapping the next band's computation.  ``tabs_list[k]`` = the two
    per-dimension tables (``A1_k.table, A2_k.table``) of local frame k, ``srcs[k]`` its device array; ``model`` a
    column-major device array [m1, m2].  ``comm=None``: this GPU only.  The first argument may be a prepared
    :class:`FrameBatch` (then ``srcs`` is omitted: ``adjoint_accumulate(batch, model)``)."""
    from . import _arrays as ar
    from . import _lib
    if isinstance(tabs_list, FrameBatch):
        batch = tabs_list
        if model is None:
            model = srcs
    else:
        batch = FrameBatch(tabs_list, srcs)
    if model is None or not ar.is_device(model) or not ar.is_colmajor(model) or len(ar.shape_of(model)) != 2:
        raise _lib.ErrorException("the model must be a column-major 2-D device array")
    dt = ar.np_dtype_of(model)
    if batch.n and batch.dtype != dt:
        raise _lib.ErrorException("frames must be column-major device arrays of the model's element type")
    m1, m2 = ar.shape_of(model)
    _lib.check(_lib.lib.li_adjoint_accumulate_allreduce(comm.handle if comm is not None else None, ar.li_dtype(dt), m1, m2,
                                                        batch.n, batch.handles, batch.ptrs, float(alpha), float(beta),
                                                        ar.data_ptr(model), int(nbands), ar.current_stream(model)))
    return model

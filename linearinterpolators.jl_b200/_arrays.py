"""Array plumbing: Julia-style (column-major) views of numpy / torch / raw device memory.

Shapes are Julia ``size`` tuples; memory must be column-major (Fortran order), as every array of
the reference is.  numpy arrays are host memory (staged through the device by the library's
``*_host`` entry points or explicit copies); torch CUDA tensors and :class:`DeviceArray` are device
memory and are used in place.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Tuple

import numpy as np

from . import _lib

try:  # torch is plumbing only (device memory, streams); the package works without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def li_dtype(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _lib.F32
    if dtype == np.float64:
        return _lib.F64
    raise TypeError(f"unsupported floating-point type {dtype}")


def np_dtype_of(x):
    if torch is not None and isinstance(x, torch.Tensor):
        return {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
                torch.complex64: np.dtype(np.complex64), torch.complex128: np.dtype(np.complex128),
                torch.int64: np.dtype(np.int64), torch.int32: np.dtype(np.int32),
                torch.float16: np.dtype(np.float16)}[x.dtype]
    return np.dtype(x.dtype)


class DeviceArray:
    """Minimal column-major device array backed by ``li_malloc`` (no torch needed)."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        _lib.check(_lib.lib.li_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value or 0
        self._owns = True

    @classmethod
    def view(cls, ptr: int, shape, dtype) -> "DeviceArray":
        """Non-owning column-major view of device memory (e.g. a buffer handed to a callback by the library)."""
        d = object.__new__(cls)
        d.shape = tuple(int(s) for s in shape)
        d.dtype = np.dtype(dtype)
        d.nbytes = int(np.prod(d.shape, dtype=np.int64)) * d.dtype.itemsize
        d.ptr, d._owns = int(ptr), False
        return d

    @classmethod
    def from_numpy(cls, a: np.ndarray, stream=0) -> "DeviceArray":
        a = np.asfortranarray(a)
        d = cls(a.shape, a.dtype)
        _lib.check(_lib.lib.li_memcpy_h2d(d.ptr, a.ctypes.data, d.nbytes, stream))
        _lib.check(_lib.lib.li_stream_sync(stream))
        return d

    def numpy(self, stream=0) -> np.ndarray:
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        _lib.check(_lib.lib.li_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream))
        _lib.check(_lib.lib.li_stream_sync(stream))
        return out

    def __del__(self):
        try:
            if getattr(self, "ptr", 0) and getattr(self, "_owns", True):
                _lib.lib.li_free(self.ptr)
            self.ptr = 0
        except Exception:  # pragma: no cover
            pass


def is_device(x) -> bool:
    if isinstance(x, DeviceArray):
        return True
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def shape_of(x) -> Tuple[int, ...]:
    return tuple(int(s) for s in x.shape)


def current_stream(x=None) -> int:
    if torch is not None and is_torch(x) and x.is_cuda:
        return int(torch.cuda.current_stream(x.device).cuda_stream)
    return 0


def _colmajor_strides(shape):
    st, acc = [], 1
    for n in shape:
        st.append(acc)
        acc *= n
    return tuple(st)


def is_colmajor(x) -> bool:
    shape = shape_of(x)
    if isinstance(x, DeviceArray):
        return True
    if is_torch(x):
        want = _colmajor_strides(shape)
        return all(n <= 1 or s == w for n, s, w in zip(shape, x.stride(), want))
    return bool(x.flags.f_contiguous)


def data_ptr(x) -> int:
    if isinstance(x, DeviceArray):
        return x.ptr
    if is_torch(x):
        return int(x.data_ptr())
    return int(x.ctypes.data)


def empty_like_kind(like, shape, dtype):
    """Uninitialised column-major array of the same kind (numpy / torch-cuda / DeviceArray) as ``like``:
    the ``Array{T}(undef, dims)`` of the reference's vcreate."""
    shape = tuple(int(s) for s in shape)
    dtype = np.dtype(dtype)
    if isinstance(like, DeviceArray):
        return DeviceArray(shape, dtype)
    if is_torch(like):
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
               np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[dtype]
        t = torch.empty(shape[::-1], dtype=tdt, device=like.device)
        return t.permute(*reversed(range(len(shape)))) if len(shape) > 1 else t
    if isinstance(like, np.ndarray) and getattr(like, "_li_pinned", False):
        return pinned_empty(shape, dtype)
    return np.empty(shape, dtype=dtype, order="F")


class _Pinned(np.ndarray):
    _li_pinned = True


def pinned_empty(shape, dtype) -> np.ndarray:
    """Column-major numpy array in page-locked host memory (``li_host_alloc``)."""
    shape = tuple(int(s) for s in shape)
    dtype = np.dtype(dtype)
    nbytes = max(1, int(np.prod(shape, dtype=np.int64)) * dtype.itemsize)
    p = C.c_void_p()
    _lib.check(_lib.lib.li_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * nbytes).from_address(p.value)
    root = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64)))
    weakref.finalize(root, _lib.lib.li_host_free, p.value)  # runs once every view of `root` is gone
    return root.reshape(shape, order="F").view(_Pinned)


def to_colmajor(x, dtype):
    """Host arrays: convert to a column-major numpy array of ``dtype`` (Julia's implicit promotion of the
    data to the operator's floating-point type).  Device arrays must already match."""
    dtype = np.dtype(dtype)
    if is_device(x):
        if np_dtype_of(x) != dtype:
            raise TypeError(f"device array has element type {np_dtype_of(x)}, operator needs {dtype}")
        if not is_colmajor(x):
            raise _lib.ErrorException("device arrays must be column-major (Julia layout); "
                                      "allocate them with li_b200.device_empty / vcreate")
        return x
    if is_torch(x):
        x = x.detach().cpu().numpy()
    if isinstance(x, np.ndarray) and x.dtype == dtype and x.flags.f_contiguous:
        return x  # (keeps the pinned-memory subclass)
    a = np.asarray(x)
    if a.dtype == dtype and a.flags.f_contiguous:
        return a
    return np.asfortranarray(a, dtype=dtype)


def device_empty(shape, dtype=np.float32, device=None):
    """Column-major device array (a torch CUDA tensor when torch is present)."""
    shape = tuple(int(s) for s in shape)
    if torch is None:
        return DeviceArray(shape, dtype)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    like = torch.empty(0, device=dev)
    return empty_like_kind(like, shape, dtype)


def to_device(a, device=None):
    """Copy a host array to a new column-major device array."""
    a = np.asfortranarray(np.asarray(a))
    if torch is None:
        return DeviceArray.from_numpy(a)
    d = device_empty(a.shape, a.dtype, device)
    if a.size:
        src = torch.from_numpy(np.ascontiguousarray(a.T))  # same bytes as the column-major array
        d.permute(*reversed(range(a.ndim))).copy_(src) if a.ndim > 1 else d.copy_(src)
    return d


def wrap_device_ptr(ptr: int, shape, dtype, like):
    """Column-major device array of the same kind as ``like`` over raw device memory (non-owning)."""
    shape = tuple(int(s) for s in shape)
    dtype = np.dtype(dtype)
    if is_torch(like):
        class _Ext:
            pass
        e = _Ext()
        n = int(np.prod(shape, dtype=np.int64))
        e.__cuda_array_interface__ = {"shape": (n,), "typestr": dtype.str, "data": (int(ptr), False), "version": 2}
        t = torch.as_tensor(e, device=like.device)
        return t.view(shape[::-1]).permute(*reversed(range(len(shape)))) if len(shape) > 1 else t
    return DeviceArray.view(ptr, shape, dtype)


def to_host(x) -> np.ndarray:
    if isinstance(x, DeviceArray):
        return x.numpy()
    if is_torch(x):
        if x.ndim > 1:
            return np.asfortranarray(x.detach().permute(*reversed(range(x.ndim))).contiguous().cpu().numpy().T)
        return x.detach().cpu().numpy()
    return np.asarray(x)


# ---- asynchronous host-staged calls -------------------------------------------------------------------------
class HostStream:
    """A CUDA stream for host-staged (numpy) applies.  Inside ``with async_host(stream):`` every apply on host
    arrays only *enqueues* its H2D copy, kernels and D2H copy on that stream and returns; the arrays must be pinned
    (:func:`pinned_empty`) and must not be touched before :meth:`synchronize`.  Applies issued on different streams
    overlap, e.g. the download of a forward result with the upload of an adjoint's input."""

    def __init__(self):
        p = C.c_void_p()
        _lib.check(_lib.lib.li_stream_create(C.byref(p)))
        self.handle = p.value

    def synchronize(self):
        _lib.check(_lib.lib.li_stream_sync(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib.li_stream_destroy(self.handle)
                self.handle = None
        except Exception:  # pragma: no cover
            pass


import contextlib
import threading

_tls = threading.local()


def host_stream_state():
    """(stream handle, asynchronous?) used by host-staged applies on this thread."""
    st = getattr(_tls, "stream", None)
    return (st.handle, True) if st is not None else (0, False)


@contextlib.contextmanager
def async_host(stream: HostStream):
    prev = getattr(_tls, "stream", None)
    _tls.stream = stream
    try:
        yield stream
    finally:
        _tls.stream = prev

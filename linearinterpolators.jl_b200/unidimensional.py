"""1-D interpolation with coefficients evaluated on the fly (src/unidimensional.jl)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as ar
from . import _lib
from .mapping import Adjoint, Direct, check_dst, check_same_kind


def _result_type(ker, src):
    # promote_type(eltype(ker), eltype(src)) (src/unidimensional.jl:44,54) restricted to the kernel types
    sdt = ar.np_dtype_of(src) if hasattr(src, "dtype") else np.asarray(src).dtype
    if sdt == np.float64 and ker.T == np.float32:
        return np.dtype(np.float64)
    return ker.T


def apply_kernel(P, ker, x, src, length=None):
    """``apply([P=Direct,] ker, x, src)`` and ``apply(Adjoint, ker, x, src, len)`` (src/unidimensional.jl:33-56)."""
    T = _result_type(ker, src)
    if P is Direct:
        like = src if ar.is_device(src) else np.empty(0)
        dst = ar.empty_like_kind(like, ar.shape_of(x) if hasattr(x, "shape") else np.asarray(x).shape, T)
    else:
        if length is None:
            raise TypeError("apply(Adjoint, ker, x, src, len)")
        like = src if ar.is_device(src) else np.empty(0)
        dst = ar.empty_like_kind(like, (int(length),), T)
    return otf1d_apply(1, P, ker, x, src, 0, dst)


def otf1d_apply(alpha, P, ker, x, src, beta, dst, variant=_lib.VARIANT_AUTO):
    """``apply!(alpha, P, ker, x, src, beta, dst)`` (src/unidimensional.jl:95-245)."""
    if callable(x) and not hasattr(x, "shape"):
        # positions given as a function of the (1-based, column-major linear) index (:136-175, :217-245)
        shape = ar.shape_of(dst if P is Direct else src)
        x = np.array([x(i) for i in range(1, int(np.prod(shape)) + 1)]).reshape(shape, order="F")
    T = ar.np_dtype_of(dst)
    if T not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("destination must be Float32 or Float64")
    ker = ker.with_eltype(T)
    adjoint = P is Adjoint
    src = ar.to_colmajor(src, T)
    xshape = ar.shape_of(x) if hasattr(x, "shape") else np.asarray(x).shape
    if adjoint:
        if ar.shape_of(src) != tuple(xshape):                                   # @assert, :197
            raise AssertionError("size(src) == size(x)")
        if len(ar.shape_of(dst)) != 1:
            raise TypeError("destination of the adjoint must be a vector")
        length = ar.shape_of(dst)[0]
    else:
        if ar.shape_of(dst) != tuple(xshape):                                   # @assert, :105
            raise AssertionError("size(dst) == size(x)")
        if len(ar.shape_of(src)) != 1:
            raise TypeError("source must be a vector")
        length = ar.shape_of(src)[0]
    check_dst(dst, ar.shape_of(dst), T)
    check_same_kind(src, dst)
    n = int(np.prod(xshape, dtype=np.int64))
    dev = ar.is_device(dst)
    if ar.is_device(x):
        if ar.np_dtype_of(x) != T:
            raise TypeError("device positions must have the kernel's floating-point type")
        dx = x
    else:
        dx = ar.to_device(np.ascontiguousarray(np.asarray(x).ravel(order="F"), dtype=T),
                          device=dst.device if ar.is_torch(dst) else None)
    if dev:
        dsrc, ddst = src, dst
    else:
        dsrc = ar.to_device(src)
        ddst = ar.empty_like_kind(dsrc, ar.shape_of(dst), T) if float(beta) == 0.0 else ar.to_device(dst)
    _lib.check(_lib.lib.li_otf1d_apply(int(adjoint), C.byref(ker.desc()), ar.data_ptr(dx), n, int(length),
                                       float(alpha), ar.data_ptr(dsrc), float(beta), ar.data_ptr(ddst), variant,
                                       ar.current_stream(ddst)))
    if not dev:
        dst[...] = ar.to_host(ddst).reshape(ar.shape_of(dst), order="F")
    return dst

"""Separable operators: fused product of table-backed factors (``li_sep_apply``) and the 2-D on-the-fly
``apply!(alpha, P, ker1, x1, ker2, x2, src, beta, dst)`` of src/separable.jl:56-162."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as ar
from . import _lib
from .mapping import Adjoint, Direct, check_dst, check_same_kind

_WS = {}


def _workspace(like, nbytes: int):
    """Grow-only device scratch for the unfused multi-pass path (one per device/kind)."""
    if nbytes == 0:
        return None, 0
    key = ("torch", str(like.device)) if ar.is_torch(like) else ("raw",)
    ws = _WS.get(key)
    if ws is None or ws[1] < nbytes:
        buf = ar.empty_like_kind(like, (nbytes // 4 + 1,), np.float32) if ar.is_device(like) else None
        ws = (buf, nbytes)
        _WS[key] = ws
    return ws[0], nbytes


def sep_apply_tables(adjoint: bool, tabs, alpha, x, beta, y, owner=None, nbatch_dims: int = 0,
                     variant=_lib.VARIANT_AUTO):
    """``y <- alpha * (A_1*...*A_n or its adjoint) * x + beta*y`` for per-dimension tables ``tabs`` (``None`` leaves a
    dimension untouched).  Trailing ``nbatch_dims`` dimensions of ``x`` are batch dimensions."""
    dtype = next(t.dtype for t in tabs if t is not None)
    x = ar.to_colmajor(x, dtype)
    xs = ar.shape_of(x)
    n = len(tabs)
    if len(xs) != n + nbatch_dims:
        raise _lib.DimensionMismatch("array rank does not match the separable operator")
    oshape = list(xs)
    other = (C.c_int64 * n)()
    for k, t in enumerate(tabs):
        other[k] = xs[k]
        if t is not None:
            nin, nout = (t.nrows, t.ncols) if adjoint else (t.ncols, t.nrows)
            if xs[k] != nin:
                raise _lib.DimensionMismatch(f"dimension {k + 1} of source must be {nin}")
            oshape[k] = nout
    nbatch = int(np.prod(xs[n:], dtype=np.int64))
    check_dst(y, tuple(oshape), dtype)
    check_same_kind(x, y)
    handles = (C.c_void_p * n)(*[t.handle if t is not None else None for t in tabs])
    if ar.is_device(y):
        # two tables on dimensions 1 and 2 (plus untouched / batch dimensions), or three tables on dimensions 1, 2, 3,
        # normally run as ONE fused kernel that needs no workspace: try that first, and only size the scratch of the
        # per-dimension passes if it declines
        if variant != _lib.VARIANT_ATOMIC and n >= 2 and tabs[0] is not None and tabs[1] is not None \
                and (all(t is None for t in tabs[2:]) or (n == 3 and tabs[2] is not None)):
            status = _lib.lib.li_sep_apply(int(adjoint), n, handles, other, nbatch, float(alpha), ar.data_ptr(x),
                                           float(beta), ar.data_ptr(y), None, 0, variant, ar.current_stream(y))
            if status == _lib.OK:
                return y
            if status != _lib.ERR_ARG or b"workspace too small" not in _lib.lib.li_last_error():
                _lib.check(status)
        need = C.c_size_t()
        _lib.check(_lib.lib.li_sep_workspace_bytes(int(adjoint), n, handles, other, nbatch, C.byref(need)))
        ws, wsb = _workspace(y, int(need.value))
        _lib.check(_lib.lib.li_sep_apply(int(adjoint), n, handles, other, nbatch, float(alpha), ar.data_ptr(x),
                                         float(beta), ar.data_ptr(y), ar.data_ptr(ws) if ws is not None else None,
                                         wsb, variant, ar.current_stream(y)))
    else:
        stream, asynchronous = ar.host_stream_state()
        fn = _lib.lib.li_sep_apply_host_async if asynchronous else _lib.lib.li_sep_apply_host
        _lib.check(fn(int(adjoint), n, handles, other, nbatch, float(alpha), ar.data_ptr(x), float(beta),
                      ar.data_ptr(y), variant, stream))
    return y


def sep_apply_batch(adjoint: bool, tabs_list, alpha, srcs, beta, dsts, variant=_lib.VARIANT_AUTO):
    """Many separable applies in ONE library call (`li_sep_apply_batch`): apply k uses the per-dimension tables
    ``tabs_list[k]``, ``srcs[k]`` and ``dsts[k]`` (device arrays).  Entries of ``dsts`` may be the same array: with
    ``beta = 1`` the applies accumulate in order, e.g. ``m += A_k' y_k`` over the frames of iterative fitting."""
    n = len(tabs_list)
    if n == 0:
        return dsts
    ndim = len(tabs_list[0])
    dtype = tabs_list[0][0].dtype
    handles = (C.c_void_p * (n * ndim))()
    sp, dp = (C.c_void_p * n)(), (C.c_void_p * n)()
    for k, (tabs, x, y) in enumerate(zip(tabs_list, srcs, dsts)):
        if len(tabs) != ndim or any(t is None for t in tabs):
            raise _lib.ErrorException("every apply needs one table per dimension")
        if not (ar.is_device(x) and ar.is_device(y)):
            raise _lib.ErrorException("sep_apply_batch works on device arrays")
        x = ar.to_colmajor(x, dtype)
        xin = tuple((t.nrows if adjoint else t.ncols) for t in tabs)
        yout = tuple((t.ncols if adjoint else t.nrows) for t in tabs)
        if ar.shape_of(x) != xin:
            raise _lib.DimensionMismatch(f"source {k} has size {ar.shape_of(x)}, expected {xin}")
        check_dst(y, yout, dtype)
        for d, t in enumerate(tabs):
            handles[k * ndim + d] = t.handle
        sp[k], dp[k] = ar.data_ptr(x), ar.data_ptr(y)
    _lib.check(_lib.lib.li_sep_apply_batch(int(adjoint), ndim, n, handles, None, 1, float(alpha), sp, float(beta), dp,
                                           variant, ar.current_stream(dsts[0])))
    return dsts


def sep2d_otf_apply(alpha, P, ker1, x1, ker2, x2, src, beta, dst):
    """src/separable.jl:56-162.  Coefficients are evaluated on the device for this call only."""
    T = ker1.T
    if ker2.T != T:
        raise TypeError("kernels must have the same floating-point type")
    adjoint = P is Adjoint
    src = ar.to_colmajor(src, T)
    n1 = int(np.prod(ar.shape_of(x1)))
    n2 = int(np.prod(ar.shape_of(x2)))
    ss, ds = ar.shape_of(src), ar.shape_of(dst)
    if len(ss) != 2 or len(ds) != 2:
        raise TypeError("source and destination must be 2-dimensional")
    if adjoint:
        if ss != (n1, n2):                                                      # @assert, :143
            raise AssertionError("size(src) == (length(x1), length(x2))")
        m1, m2 = ds
    else:
        if ds != (n1, n2):                                                      # @assert, :78
            raise AssertionError("size(dst) == (length(x1), length(x2))")
        m1, m2 = ss
    check_dst(dst, ds, T)
    check_same_kind(src, dst)
    dev = ar.is_device(dst)
    stream = ar.current_stream(dst) if dev else 0

    def dev_vec(x):
        if ar.is_device(x):
            if ar.np_dtype_of(x) != T:
                raise TypeError("device positions must have the kernel's floating-point type")
            return x
        return ar.to_device(np.ascontiguousarray(np.asarray(x).ravel(), dtype=T),
                            device=dst.device if ar.is_torch(dst) else None)

    dx1, dx2 = dev_vec(x1), dev_vec(x2)
    if dev:
        dsrc, ddst = src, dst
    else:
        dsrc = ar.to_device(src)
        ddst = ar.empty_like_kind(dsrc, ds, T) if float(beta) == 0.0 else ar.to_device(dst)
    _lib.check(_lib.lib.li_sep2d_otf_apply(int(adjoint), C.byref(ker1.desc()), ar.data_ptr(dx1), n1,
                                           C.byref(ker2.desc()), ar.data_ptr(dx2), n2, m1, m2, 1, float(alpha),
                                           ar.data_ptr(dsrc), float(beta), ar.data_ptr(ddst),
                                           ar.current_stream(ddst) if not dev else stream))
    if not dev:
        dst[...] = ar.to_host(ddst)
    return dst


def sep_otf_apply(alpha, P, kers, xs, src, beta, dst):
    """N-d (1 to 3 dimensions) separable interpolation with on-the-fly coefficients: one kernel and one vector of
    positions per dimension -- the generalisation of src/separable.jl:56-162 the reference plans (src/separable.jl:14
    "use a table for the innermost dimensions").  Tables are generated on the device for this call
    (``li_sep_otf_apply``), then the fused 2-D / 3-D kernels run."""
    nd = len(kers)
    if nd != len(xs) or not 1 <= nd <= 3:
        raise TypeError("one kernel and one position vector per dimension (1 to 3 dimensions)")
    T = kers[0].T
    if any(k.T != T for k in kers):
        raise TypeError("kernels must have the same floating-point type")
    adjoint = P is Adjoint
    src = ar.to_colmajor(src, T)
    ns = tuple(int(np.prod(ar.shape_of(x))) for x in xs)
    ss, ds = ar.shape_of(src), ar.shape_of(dst)
    if len(ss) != nd or len(ds) != nd:
        raise TypeError(f"source and destination must be {nd}-dimensional")
    if (ss if adjoint else ds) != ns:
        raise AssertionError("size(src) == map(length, x)" if adjoint else "size(dst) == map(length, x)")
    ms = ds if adjoint else ss
    check_dst(dst, ds, T)
    check_same_kind(src, dst)
    dev = ar.is_device(dst)

    def dev_vec(x):
        if ar.is_device(x):
            if ar.np_dtype_of(x) != T:
                raise TypeError("device positions must have the kernel's floating-point type")
            return x
        return ar.to_device(np.ascontiguousarray(np.asarray(x).ravel(), dtype=T), device=dst.device if ar.is_torch(dst) else None)

    dxs = [dev_vec(x) for x in xs]
    if dev:
        dsrc, ddst = src, dst
    else:
        dsrc = ar.to_device(src)
        ddst = ar.empty_like_kind(dsrc, ds, T) if float(beta) == 0.0 else ar.to_device(dst)
    descs = [k.desc() for k in kers]
    kp = (C.c_void_p * nd)(*[C.cast(C.pointer(d), C.c_void_p) for d in descs])
    xp = (C.c_void_p * nd)(*[ar.data_ptr(x) for x in dxs])
    na, ma = (C.c_int64 * nd)(*ns), (C.c_int64 * nd)(*ms)
    _lib.check(_lib.lib.li_sep_otf_apply(int(adjoint), nd, kp, xp, na, ma, 1, float(alpha), ar.data_ptr(dsrc), float(beta),
                                         ar.data_ptr(ddst), ar.current_stream(ddst)))
    if not dev:
        dst[...] = ar.to_host(ddst)
    return dst

"""Vector operations and the linear conjugate-gradient driver of iterative fitting: LazyAlgebra's ``vdot``,
``vupdate!``, ``vcombine!`` and ``conjgrad`` as the reference uses them in ``x = conjgrad(H'*H, H'*y; ftol=1e-8,
maxiter=50)`` (test/benchmarks.jl:31-40; purpose: README.md:36-37).  Every operation is a call through the C ABI
(``li_vdot`` ...); the scalars of the recurrence stay on the device, so an iteration costs no host round trip."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _arrays as ar
from . import _lib
from .mapping import Adjoint, Composition, Direct, Mapping, AdjointMapping


def _n(x) -> int:
    return int(np.prod(ar.shape_of(x), dtype=np.int64))


def _dev(x, what):
    if not ar.is_device(x):
        raise _lib.ErrorException(f"{what}: device arrays are required (li_b200.to_device / device_empty)")
    if not ar.is_colmajor(x):
        raise _lib.ErrorException(f"{what}: arrays must be column-major")
    return x


def vdot_device(x, y) -> float:
    """``vdot(x, y)`` of two real device arrays, accumulated in Float64 on the device in a fixed order."""
    _dev(x, "vdot"), _dev(y, "vdot")
    if ar.shape_of(x) != ar.shape_of(y) or ar.np_dtype_of(x) != ar.np_dtype_of(y):
        raise _lib.DimensionMismatch("vdot: arguments must have the same size and element type")
    out = C.c_double()
    _lib.check(_lib.lib.li_vdot(ar.li_dtype(ar.np_dtype_of(x)), ar.data_ptr(x), ar.data_ptr(y), _n(x), C.byref(out),
                                ar.current_stream(x)))
    return float(out.value)


def vupdate(y, alpha, x):
    """``vupdate!(y, alpha, x)``: ``y += alpha*x`` (device arrays)."""
    _dev(x, "vupdate!"), _dev(y, "vupdate!")
    if ar.shape_of(x) != ar.shape_of(y) or ar.np_dtype_of(x) != ar.np_dtype_of(y):
        raise _lib.DimensionMismatch("vupdate!: arguments must have the same size and element type")
    _lib.check(_lib.lib.li_vupdate(ar.li_dtype(ar.np_dtype_of(y)), ar.data_ptr(y), float(alpha), ar.data_ptr(x), _n(y),
                                   ar.current_stream(y)))
    return y


def vcombine(dst, alpha, x, beta, y):
    """``vcombine!(dst, alpha, x, beta, y)``: ``dst = alpha*x + beta*y`` (device arrays; ``dst`` may alias)."""
    for a in (dst, x, y):
        _dev(a, "vcombine!")
        if ar.shape_of(a) != ar.shape_of(dst) or ar.np_dtype_of(a) != ar.np_dtype_of(dst):
            raise _lib.DimensionMismatch("vcombine!: arguments must have the same size and element type")
    _lib.check(_lib.lib.li_vcombine(ar.li_dtype(ar.np_dtype_of(dst)), ar.data_ptr(dst), float(alpha), ar.data_ptr(x),
                                    float(beta), ar.data_ptr(y), _n(dst), ar.current_stream(dst)))
    return dst


class CgResult:
    """What :func:`conjgrad` reports besides the solution."""
    STATUS = {_lib.CG_MAXITER: "too many iterations", _lib.CG_FTEST: "function reduction test satisfied",
              _lib.CG_GTEST: "gradient test satisfied", _lib.CG_NOT_POSITIVE: "operator is not positive definite"}

    def __init__(self, info: _lib.CgInfo):
        self.iters, self.status = int(info.iters), int(info.status)
        self.rho0, self.rho = float(info.rho0), float(info.rho)

    def __repr__(self):
        return f"CgResult(iters={self.iters}, status={self.STATUS.get(self.status)!r}, rho0={self.rho0:.6g}, rho={self.rho:.6g})"


def _options(maxiter, ftol, gtol, check_every):
    g = (gtol, 0.0) if np.isscalar(gtol) else tuple(gtol)
    return _lib.CgOptions(int(maxiter), int(check_every), float(ftol), float(g[0]), float(g[1]))


def _gram_of_separable(A):
    """``H'*H`` (or ``Gram(H)``) with H a product of table-backed interpolators along distinct dimensions ->
    the per-dimension tables of H, else None."""
    from .sparse import SparseUnidimensionalInterpolator
    from .tabulated import TabulatedInterpolator
    if not isinstance(A, Composition):
        return None
    f = A.factors
    if len(f) % 2 or not f:
        return None
    half = len(f) // 2
    left, right = f[:half], f[half:]
    if not all(isinstance(m, AdjointMapping) for m in left) or any(isinstance(m, AdjointMapping) for m in right):
        return None
    if [m.parent for m in left] != list(reversed(right)):
        return None
    dims = {}
    for m in right:
        if not isinstance(m, (TabulatedInterpolator, SparseUnidimensionalInterpolator)) or not isinstance(m.d, int):
            return None
        if m.d in dims:
            return None
        dims[m.d] = m.table
    return dims


def conjgrad(A: Mapping, b, x0=None, *, ftol=1e-8, gtol=(0.0, 0.0), maxiter=50, check_every=5, verb=False, info=False):
    """``conjgrad(A, b[, x0]; ftol, gtol, maxiter, verb)`` of LazyAlgebra on device arrays: solves ``A*x = b`` for a
    symmetric positive definite mapping, e.g. ``conjgrad(H.adj*H, H.adj*y)`` (test/benchmarks.jl:40).  The loop runs
    inside the library (``li_conjgrad``); ``A`` is applied through a callback, and when ``A`` is the Gram operator of
    a product of table-backed interpolators the callback never leaves the library either."""
    _dev(b, "conjgrad")
    dt = ar.np_dtype_of(b)
    x = ar.empty_like_kind(b, ar.shape_of(b), dt) if x0 is None else _dev(x0, "conjgrad")
    if ar.shape_of(x) != ar.shape_of(b) or ar.np_dtype_of(x) != dt:
        raise _lib.DimensionMismatch("conjgrad: x0 and b must have the same size and element type")
    opts = _options(maxiter, ftol, gtol, check_every)
    res = _lib.CgInfo()
    stream = ar.current_stream(b)
    shape = ar.shape_of(b)
    err = []

    def cb(_ctx, xp, yp, _stream):
        try:
            A._apply(1, Direct, ar.wrap_device_ptr(xp, shape, dt, b), False, 0, ar.wrap_device_ptr(yp, shape, dt, b))
            return _lib.OK
        except Exception as exc:  # reported after the C call returns
            err.append(exc)
            return _lib.ERR_ARG

    fn = _lib.APPLY_FN(cb)
    status = _lib.lib.li_conjgrad(fn, None, ar.li_dtype(dt), _n(b), ar.data_ptr(b), ar.data_ptr(x), int(x0 is None),
                                  C.byref(opts), C.byref(res), stream)
    if err:
        raise err[0]
    _lib.check(status)
    out = CgResult(res)
    if verb:
        print(f"# conjgrad: {out}")
    return (x, out) if info else x


def fit_separable(tabs, y, mu=0.0, x0=None, *, ftol=1e-8, gtol=(0.0, 0.0), maxiter=50, check_every=5, info=False):
    """``argmin |H*x - y|^2 + mu*|x|^2`` for ``H = A_1*...*A_n`` given by per-dimension tables (``None`` = untouched
    dimension of that length in ``y``): ``conjgrad(H'*H + mu*I, H'*y)`` run entirely inside the library
    (``li_conjgrad_sep``).  ``y`` is a device array; returns the device array ``x``."""
    _dev(y, "fit_separable")
    n = len(tabs)
    ys = ar.shape_of(y)
    if len(ys) != n:
        raise _lib.DimensionMismatch("array rank does not match the separable operator")
    dt = next(t.dtype for t in tabs if t is not None)
    if ar.np_dtype_of(y) != dt:
        raise TypeError("data must have the operator's floating-point type")
    xshape = tuple(t.ncols if t is not None else ys[k] for k, t in enumerate(tabs))
    for k, t in enumerate(tabs):
        if t is not None and t.nrows != ys[k]:
            raise _lib.DimensionMismatch(f"dimension {k + 1} of the data must be {t.nrows}")
    x = ar.empty_like_kind(y, xshape, dt) if x0 is None else _dev(x0, "fit_separable")
    if ar.shape_of(x) != xshape:
        raise _lib.DimensionMismatch(f"x0 has size {ar.shape_of(x)}, expected {xshape}")
    handles = (C.c_void_p * n)(*[t.handle if t is not None else None for t in tabs])
    other = (C.c_int64 * n)(*[int(s) for s in ys])
    opts = _options(maxiter, ftol, gtol, check_every)
    res = _lib.CgInfo()
    _lib.check(_lib.lib.li_conjgrad_sep(n, handles, other, 1, float(mu), ar.data_ptr(y), ar.data_ptr(x), int(x0 is None),
                                        C.byref(opts), C.byref(res), ar.current_stream(y)))
    return (x, CgResult(res)) if info else x

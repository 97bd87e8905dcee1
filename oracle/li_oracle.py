"""ctypes front-end of the CPU oracle (``oracle/libli_oracle.so``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
CPU-baseline legs of ``bench.py``; never by the product package.  See ``li_oracle.h`` for
the parity status ("weights: parity unpinned").

Conventions follow the reference (Julia): arrays are column-major, indices in tables are
1-based Int64, shapes are given as Julia ``size`` tuples, ``d`` is a 1-based dimension.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libli_oracle.so")

RECTANGULAR, LINEAR, QUADRATIC, CUBIC, CATMULLROM, CARDINALCUBIC, KEYS, MITCHELLNETRAVALI, LANCZOS = range(9)
FLAT, SAFEFLAT = 0, 1
OK, ERR_ARG, ERR_INEXACT = 0, 1, 3
MAX_S = 16


class InexactError(ArithmeticError):
    """Julia's InexactError from trunc(Int, f) (src/boundaries.jl:81,105,147)."""


class KernelDesc(C.Structure):
    _fields_ = [("family", C.c_int32), ("S", C.c_int32), ("boundary", C.c_int32),
                ("prime", C.c_int32), ("p", C.c_double * 2)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("li_oracle.c", "li_oracle_impl.inc", "li_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libli_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lio_sin_f64.restype = C.c_double
        _lib.lio_sin_f64.argtypes = [C.c_double]
        _lib.lio_sin_f32.restype = C.c_float
        _lib.lio_sin_f32.argtypes = [C.c_float]
    return _lib


def kernel(family: int, S: int, boundary: int = FLAT, params=(), prime: bool = False) -> KernelDesc:
    """`prime`: the derivative kernel (`*Prime` of src/LinearInterpolators.jl:23-41)."""
    k = KernelDesc()
    k.family, k.S, k.boundary, k.prime = family, S, boundary, int(bool(prime))
    p = list(params) + [0.0, 0.0]
    k.p[0], k.p[1] = float(p[0]), float(p[1])
    return k


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"unsupported floating-point type {dtype}")


def _ct(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _farr(a, dtype) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=dtype))


def _check(st: int):
    if st == ERR_INEXACT:
        raise InexactError("trunc(Int, f): position is NaN or out of Int64 range")
    if st != OK:
        raise ValueError(f"oracle status {st}")


def sin(x, dtype=np.float64):
    l = lib()
    if np.dtype(dtype) == np.float32:
        return np.float32(l.lio_sin_f32(float(np.float32(x))))
    return np.float64(l.lio_sin_f64(float(x)))


def limits(k: KernelDesc, length: int, dtype):
    ct = _ct(dtype)
    inf, sup = ct(), ct()
    getattr(lib(), "lio_limits_" + _sfx(dtype))(C.byref(k), C.c_int64(length), C.byref(inf), C.byref(sup))
    return np.dtype(dtype).type(inf.value), np.dtype(dtype).type(sup.value)


def getweights(k: KernelDesc, t, dtype) -> np.ndarray:
    w = np.zeros(k.S, dtype=dtype)
    _check(getattr(lib(), "lio_getweights_" + _sfx(dtype))(C.byref(k), _ct(dtype)(float(np.dtype(dtype).type(t))), _p(w)))
    return w


def getcoefs(k: KernelDesc, length: int, x, dtype):
    """-> (j[S] int64 1-based, w[S]) -- src/boundaries.jl:77-156."""
    j = np.zeros(k.S, dtype=np.int64)
    w = np.zeros(k.S, dtype=dtype)
    xs = np.dtype(dtype).type(x)  # convert_coordinate(T, x), src/boundaries.jl:166
    _check(getattr(lib(), "lio_getcoefs_" + _sfx(dtype))(C.byref(k), C.c_int64(length), _ct(dtype)(float(xs)), _p(j), _p(w)))
    return j, w


def maketable(k: KernelDesc, x, ncols: int, dtype):
    """-> (J[S, nrows] int64 1-based, W[S, nrows]), column-major -- src/tabulated.jl:137-162."""
    x = np.ascontiguousarray(np.asarray(x).ravel(order="F"), dtype=dtype)
    nrows = x.size
    J = np.zeros((k.S, nrows), dtype=np.int64, order="F")
    W = np.zeros((k.S, nrows), dtype=dtype, order="F")
    _check(getattr(lib(), "lio_maketable_" + _sfx(dtype))(C.byref(k), _p(x), C.c_int64(nrows), C.c_int64(ncols), _p(J), _p(W)))
    return J, W


def fractional_index(pos, first, step, last, length: int, dtype) -> np.ndarray:
    """src/sparse.jl:494-502."""
    pos = np.ascontiguousarray(np.asarray(pos).ravel(order="F"), dtype=dtype)
    out = np.empty_like(pos)
    ct = _ct(dtype)
    t = np.dtype(dtype).type
    getattr(lib(), "lio_fractional_index_" + _sfx(dtype))(
        _p(pos), C.c_int64(pos.size), ct(float(t(first))), ct(float(t(step))), ct(float(t(last))),
        C.c_int64(length), _p(out))
    return out


def _split(shape, d: int):
    npre = int(np.prod(shape[: d - 1], dtype=np.int64))
    npost = int(np.prod(shape[d:], dtype=np.int64))
    return npre, npost


def tab_direct(J, W, ncols: int, d: int, src, alpha=1, beta=0, dst=None, nthreads: int = 1):
    """dst = alpha*A*src + beta*dst along (1-based) dimension d -- src/tabulated.jl:329-439."""
    dtype = W.dtype
    S, nrows = W.shape
    src = _farr(src, dtype)
    assert src.shape[d - 1] == ncols
    npre, npost = _split(src.shape, d)
    oshape = tuple(nrows if i == d - 1 else n for i, n in enumerate(src.shape))
    if dst is None:
        dst = np.full(oshape, np.nan, dtype=dtype, order="F")
    assert dst.shape == oshape and dst.flags.f_contiguous and dst.dtype == dtype
    ct = _ct(dtype)
    getattr(lib(), "lio_tab_direct_" + _sfx(dtype))(
        C.c_int(S), C.c_int64(nrows), _p(np.asfortranarray(J)), _p(np.asfortranarray(W)), C.c_int64(npre),
        C.c_int64(ncols), C.c_int64(npost), ct(alpha), _p(src), ct(beta), _p(dst), C.c_int(nthreads))
    return dst


def tab_adjoint(J, W, ncols: int, d: int, src, alpha=1, beta=0, dst=None, skip_zero: bool = True,
                nthreads: int = 1):
    """dst = alpha*A'*src + beta*dst -- src/tabulated.jl:362-392,441-478."""
    dtype = W.dtype
    S, nrows = W.shape
    src = _farr(src, dtype)
    assert src.shape[d - 1] == nrows
    npre, npost = _split(src.shape, d)
    oshape = tuple(ncols if i == d - 1 else n for i, n in enumerate(src.shape))
    if dst is None:
        dst = np.full(oshape, np.nan, dtype=dtype, order="F")
    assert dst.shape == oshape and dst.flags.f_contiguous and dst.dtype == dtype
    ct = _ct(dtype)
    getattr(lib(), "lio_tab_adjoint_" + _sfx(dtype))(
        C.c_int(S), C.c_int64(nrows), _p(np.asfortranarray(J)), _p(np.asfortranarray(W)), C.c_int64(npre),
        C.c_int64(ncols), C.c_int64(npost), ct(alpha), _p(src), ct(beta), _p(dst), C.c_int(int(skip_zero)),
        C.c_int(nthreads))
    return dst


def otf1d_direct(k, x, src, dtype, alpha=1, beta=0, dst=None):
    """src/unidimensional.jl:95-134."""
    x = _farr(x, dtype)
    src = np.ascontiguousarray(src, dtype=dtype)
    if dst is None:
        dst = np.full(x.shape, np.nan, dtype=dtype, order="F")
    assert dst.shape == x.shape and dst.flags.f_contiguous and dst.dtype == np.dtype(dtype)
    ct = _ct(dtype)
    _check(getattr(lib(), "lio_otf1d_direct_" + _sfx(dtype))(
        C.byref(k), _p(x), C.c_int64(x.size), C.c_int64(src.size), ct(alpha), _p(src), ct(beta), _p(dst)))
    return dst


def otf1d_adjoint(k, x, src, length: int, dtype, alpha=1, beta=0, dst=None):
    """src/unidimensional.jl:188-215."""
    x = _farr(x, dtype)
    src = _farr(src, dtype)
    assert src.shape == x.shape
    if dst is None:
        dst = np.full(length, np.nan, dtype=dtype)
    assert dst.shape == (length,) and dst.dtype == np.dtype(dtype)
    ct = _ct(dtype)
    _check(getattr(lib(), "lio_otf1d_adjoint_" + _sfx(dtype))(
        C.byref(k), _p(x), C.c_int64(x.size), C.c_int64(length), ct(alpha), _p(src), ct(beta), _p(dst)))
    return dst


def sep2d(adjoint: bool, k1, x1, k2, x2, src, other_shape, dtype, alpha=1, beta=0, dst=None):
    """src/separable.jl:56-162.  Direct: src is (m1,m2) -> dst (n1,n2); Adjoint: src is (n1,n2) -> dst (m1,m2).
    ``other_shape`` is the shape of the result."""
    x1 = np.ascontiguousarray(x1, dtype=dtype)
    x2 = np.ascontiguousarray(x2, dtype=dtype)
    src = _farr(src, dtype)
    n1, n2 = x1.size, x2.size
    if adjoint:
        assert src.shape == (n1, n2)
        m1, m2 = other_shape
        oshape = (m1, m2)
    else:
        m1, m2 = src.shape
        oshape = (n1, n2)
        assert tuple(other_shape) == oshape
    if dst is None:
        dst = np.full(oshape, np.nan, dtype=dtype, order="F")
    assert dst.shape == oshape and dst.flags.f_contiguous
    ct = _ct(dtype)
    fn = getattr(lib(), ("lio_sep2d_adjoint_" if adjoint else "lio_sep2d_direct_") + _sfx(dtype))
    _check(fn(C.byref(k1), _p(x1), C.c_int64(n1), C.byref(k2), _p(x2), C.c_int64(n2), C.c_int64(m1),
              C.c_int64(m2), ct(alpha), _p(src), ct(beta), _p(dst)))
    return dst


def affine2d(adjoint: bool, k1, k2, R, rows, cols, src, dtype, alpha=1, beta=0, dst=None, nthreads: int = 1):
    """src/nonseparable.jl:147-264.  R = (xx, xy, x, yx, yy, y); rows = output size of the direct
    operator, cols = its input size."""
    R = np.ascontiguousarray(R, dtype=dtype)
    assert R.size == 6
    src = _farr(src, dtype)
    n1, n2 = rows
    m1, m2 = cols
    oshape = (m1, m2) if adjoint else (n1, n2)
    assert src.shape == ((n1, n2) if adjoint else (m1, m2))
    if dst is None:
        dst = np.full(oshape, np.nan, dtype=dtype, order="F")
    assert dst.shape == oshape and dst.flags.f_contiguous
    ct = _ct(dtype)
    if adjoint:
        st = getattr(lib(), "lio_affine2d_adjoint_" + _sfx(dtype))(
            C.byref(k1), C.byref(k2), _p(R), C.c_int64(n1), C.c_int64(n2), C.c_int64(m1), C.c_int64(m2),
            ct(alpha), _p(src), ct(beta), _p(dst))
    else:
        st = getattr(lib(), "lio_affine2d_direct_" + _sfx(dtype))(
            C.byref(k1), C.byref(k2), _p(R), C.c_int64(n1), C.c_int64(n2), C.c_int64(m1), C.c_int64(m2),
            ct(alpha), _p(src), ct(beta), _p(dst), C.c_int(nthreads))
    _check(st)
    return dst


class weight_mode:
    """Context manager for the error-scale helpers of the tests: inside it, the affine loops use |w| (mode 1) or
    (w != 0) (mode 2) in place of every weight."""

    def __init__(self, mode: int):
        self.mode = int(mode)

    def __enter__(self):
        lib().lio_set_weight_mode(self.mode)
        return self

    def __exit__(self, *exc):
        lib().lio_set_weight_mode(0)
        return False


def max_threads() -> int:
    return int(lib().lio_max_threads())


def AtWA(J, W, ncols: int, w=None):
    """AtA! / AtWA! -- src/sparse.jl:299-356: dense A'*diag(w)*A (w None: A'*A) of the table (J, W)."""
    dtype = W.dtype
    S, nrows = W.shape
    dst = np.empty((ncols, ncols), dtype=dtype, order="F")
    wv = None if w is None else np.ascontiguousarray(np.asarray(w, dtype=dtype).reshape(-1, order="F"))
    st = getattr(lib(), "lio_AtWA_" + _sfx(dtype))(C.c_int(S), C.c_int64(nrows), _p(np.asfortranarray(J)), _p(np.asfortranarray(W)),
                                                   C.c_int64(ncols), _p(wv) if wv is not None else None, _p(dst))
    if st != 0:
        raise RuntimeError("corrupted interpolator table")
    return dst


def regularize(A, eps, mu):
    """regularize! -- src/sparse.jl:451-490, on a copy."""
    A = np.array(A, order="F", copy=True)
    st = getattr(lib(), "lio_regularize_" + _sfx(A.dtype))(_p(A), C.c_int64(A.shape[0]), _ct(A.dtype)(eps), _ct(A.dtype)(mu))
    if st != 0:
        raise RuntimeError("we have a problem!")
    return A

"""Host-side mirror of the LazyAlgebra mapping API the reference extends
(src/LinearInterpolators.jl:62-63): ``Direct``/``Adjoint`` tags, ``LinearMapping``, ``vcreate``,
``apply``, ``apply!`` (spelled :func:`apply_b`, alias ``apply_``), composition ``A*B`` and ``A'``
(spelled ``A.adj``).  No arithmetic happens here: every method ends in a call through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _arrays as ar
from . import _lib
from .kernels import Kernel


class Direct:
    """Tag type: apply the mapping itself."""


class Adjoint:
    """Tag type: apply the adjoint of the mapping."""


Operations = (Direct, Adjoint)


def _is_op(x) -> bool:
    return isinstance(x, type) and x in Operations


def _is_array(x) -> bool:
    return isinstance(x, (np.ndarray, ar.DeviceArray, list, tuple)) or ar.is_torch(x)


class Table:
    """Owner of a device-resident ``li_table`` (the J / W arrays of the reference's operators)."""

    def __init__(self, handle: int):
        self.handle = handle
        info = _lib.TableInfo()
        _lib.check(_lib.lib.li_table_get_info(handle, C.byref(info)))
        self.S, self.nrows, self.ncols = int(info.S), int(info.nrows), int(info.ncols)
        self.dtype = np.dtype(np.float32 if info.dtype == _lib.F32 else np.float64)
        self.windowed, self.monotone = bool(info.windowed), bool(info.monotone)
        self._J = self._W = None

    @classmethod
    def from_positions(cls, ker: Kernel, x, ncols: int) -> "Table":
        """``__maketables`` (src/tabulated.jl:137-162) / ``_sparsecoefs`` (src/sparse.jl:140-164)."""
        h = C.c_void_p()
        if ar.is_device(x):
            if ar.np_dtype_of(x) != ker.T:
                raise TypeError("device positions must have the kernel's floating-point type")
            n = int(np.prod(ar.shape_of(x), dtype=np.int64))
            _lib.check(_lib.lib.li_table_create(C.byref(ker.desc()), ar.data_ptr(x), 1, n, int(ncols),
                                                ar.current_stream(x), C.byref(h)))
        else:
            xs = np.ascontiguousarray(np.asarray(x).ravel(order="F"), dtype=ker.T)  # convert(T, X[i])
            _lib.check(_lib.lib.li_table_create(C.byref(ker.desc()), xs.ctypes.data, 0, xs.size, int(ncols), 0,
                                                C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, J, W, ncols: int) -> "Table":
        """Adopt reference-layout tables: J Int64 1-based [S, nrows], W [S, nrows]."""
        W = np.asfortranarray(W)
        if W.dtype not in (np.float32, np.float64):
            raise TypeError("weights must be Float32 or Float64")
        J = np.asfortranarray(J, dtype=np.int64)
        if W.ndim == 1:
            raise _lib.ErrorException("weights must be an S x nrows matrix")
        if J.shape != W.shape:
            raise _lib.ErrorException("corrupted interpolation table")  # src/tabulated.jl:202
        S, nrows = W.shape
        h = C.c_void_p()
        _lib.check(_lib.lib.li_table_import(ar.li_dtype(W.dtype), S, nrows, int(ncols), J.ctypes.data, W.ctypes.data,
                                            0, C.byref(h)))
        return cls(h.value)

    def _export(self):
        if self._J is None:
            J = np.empty((self.S, self.nrows), dtype=np.int64, order="F")
            W = np.empty((self.S, self.nrows), dtype=self.dtype, order="F")
            _lib.check(_lib.lib.li_table_export(self.handle, J.ctypes.data, W.ctypes.data, 0))
            self._J, self._W = J, W

    @property
    def J(self) -> np.ndarray:
        self._export()
        return self._J

    @property
    def W(self) -> np.ndarray:
        self._export()
        return self._W

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib.li_table_destroy(self.handle)
                self.handle = None
        except Exception:  # pragma: no cover
            pass


# ---------------------------------------------------------------------------------------------
class Mapping:
    """LazyAlgebra ``Mapping``: callable, ``A*x``, ``A*B`` and ``A.adj`` (Julia ``A'``)."""

    def __call__(self, x):
        return apply(self, x)

    def __mul__(self, other):
        if isinstance(other, Mapping):
            return Composition(_factors(self) + _factors(other))
        return apply(Direct, self, other)

    def __matmul__(self, other):
        return self.__mul__(other)

    @property
    def adj(self):
        return AdjointMapping(self)

    T = adj
    H = adj

    # subclasses implement:
    def _vcreate(self, P, x):
        raise NotImplementedError

    def _apply(self, alpha, P, x, scratch, beta, y):
        raise NotImplementedError


class LinearMapping(Mapping):
    pass


def _factors(A):
    return list(A.factors) if isinstance(A, Composition) else [A]


class AdjointMapping(LinearMapping):
    """``A'``."""

    def __init__(self, parent: Mapping):
        self.parent = parent

    @property
    def adj(self):
        return self.parent

    T = adj
    H = adj

    @staticmethod
    def _flip(P):
        return Adjoint if P is Direct else Direct

    def _vcreate(self, P, x, **kw):      # (d=...: Tabulated without a stored dimension, src/tabulated.jl:236-257)
        return self.parent._vcreate(self._flip(P), x, **kw)

    def _apply(self, alpha, P, x, scratch, beta, y, **kw):
        return self.parent._apply(alpha, self._flip(P), x, scratch, beta, y, **kw)


class Composition(LinearMapping):
    """``A1*A2*...``: ``H*x = A1*(A2*(...x))`` with one vcreate'd temporary per factor; when every factor
    is a table-backed interpolator along a distinct dimension of a 2-D/3-D array the whole product goes
    through ``li_sep_apply`` (one fused pass for 2-D) instead."""

    def __init__(self, factors: Sequence[Mapping]):
        self.factors = list(factors)
        self._ws = None

    @property
    def adj(self):
        return Composition([AdjointMapping(f) if not isinstance(f, AdjointMapping) else f.parent
                            for f in reversed(self.factors)])

    T = adj
    H = adj

    def _order(self, P):
        # Direct: rightmost factor first.  Adjoint of a product: leftmost factor's adjoint first.
        return list(reversed(self.factors)) if P is Direct else list(self.factors)

    def _vcreate(self, P, x):
        y = x
        for f in self._order(P):
            y = f._vcreate(P, y)
        return y

    @staticmethod
    def _sep_group(factors, N):
        """``factors`` (any order) are table-backed interpolators along distinct dimensions of an N-d array, all
        direct or all adjoint -> (tabs, flipped), else None."""
        from .tabulated import TabulatedInterpolator
        from .sparse import SparseUnidimensionalInterpolator
        if N < 2 or N > 3 or len(factors) < 2:
            return None
        tabs = [None] * N
        flips = set()
        for f in factors:
            g, flip = (f.parent, True) if isinstance(f, AdjointMapping) else (f, False)
            if not isinstance(g, (TabulatedInterpolator, SparseUnidimensionalInterpolator)):
                return None
            d = g.d
            if not isinstance(d, int) or not (1 <= d <= N) or tabs[d - 1] is not None:
                return None
            tabs[d - 1] = g.table
            flips.add(flip)
        if len(flips) != 1:
            return None
        return tabs, flips.pop()

    def _separable(self, P, x):
        g = self._sep_group(self.factors, len(ar.shape_of(x)))
        if g is None:
            return None
        tabs, flip = g
        return tabs, (P is Adjoint) != flip

    def _runs(self, P, N):
        """The factors in application order, cut into maximal runs that one fused separable apply can take
        (e.g. ``H'*H`` with ``H = A1*A2`` is two runs: H, then H')."""
        runs, cur = [], []
        for f in self._order(P):
            if self._is_tab(f, N):
                if cur and self._sep_group(cur + [f], N) is not None:
                    cur.append(f)
                else:
                    if cur:
                        runs.append(cur)
                    cur = [f]
            else:
                if cur:
                    runs.append(cur)
                    cur = []
                runs.append([f])
        if cur:
            runs.append(cur)
        return runs

    @staticmethod
    def _is_tab(f, N):
        from .tabulated import TabulatedInterpolator
        from .sparse import SparseUnidimensionalInterpolator
        g = f.parent if isinstance(f, AdjointMapping) else f
        return isinstance(g, (TabulatedInterpolator, SparseUnidimensionalInterpolator)) and isinstance(g.d, int) \
            and 1 <= g.d <= N

    def _apply(self, alpha, P, x, scratch, beta, y):
        from .separable import sep_apply_tables
        # (complex data: each factor handles Complex{T} as T with a leading dimension of 2, src/tabulated.jl:196;
        # the fused kernels work on real arrays of the table's type)
        real = _is_array(x) and not isinstance(x, (list, tuple)) and ar.np_dtype_of(x).kind != "c"
        sep = self._separable(P, x) if real else None
        if sep is not None:
            tabs, adjoint = sep
            return sep_apply_tables(adjoint, tabs, alpha, x, beta, y, owner=self)
        N = len(ar.shape_of(x)) if _is_array(x) and not isinstance(x, (list, tuple)) else 0
        runs = self._runs(P, N) if real and 2 <= N <= 3 else [[f] for f in self._order(P)]
        cur = x
        for k, run in enumerate(runs):
            last = k == len(runs) - 1
            a, b_ = (alpha, beta) if last else (1, 0)
            if len(run) >= 2:
                tabs, flip = self._sep_group(run, N)
                adjoint = (P is Adjoint) != flip
                if last:
                    out = y
                else:
                    out = cur
                    for f in run:
                        out = f._vcreate(P, out)
                sep_apply_tables(adjoint, tabs, a, cur, b_, out, owner=self)
            else:
                f = run[0]
                out = y if last else f._vcreate(P, cur)
                f._apply(a, P, cur, scratch, b_, out)
            cur = out
        return y


class Gram(Composition):
    """``Gram(H) == H'*H`` (LazyAlgebra; test/benchmarks.jl:31-40)."""

    def __init__(self, H: Mapping):
        super().__init__(_factors(H.adj) + _factors(H))


# ---------------------------------------------------------------------------------------------
def vcreate(P, A, *args):
    """``vcreate(P, A, x, scratch=false)`` and ``vcreate(P, A, d, x, scratch=false)``
    (src/tabulated.jl:218-244,284-291; src/sparse.jl:187-199,603-617; src/nonseparable.jl:65-75)."""
    if not _is_op(P):
        raise TypeError("vcreate(P, A, x[, scratch]): P must be Direct or Adjoint")
    args = list(args)
    if args and isinstance(args[-1], bool):
        args.pop()
    if len(args) == 2 and isinstance(args[0], (int, np.integer)):
        return A._vcreate(P, args[1], d=int(args[0]))
    if len(args) != 1:
        raise TypeError("vcreate(P, A, [d,] x[, scratch])")
    return A._vcreate(P, args[0])


def apply(*args):
    """``apply(A, x)``, ``apply(P, A, x[, scratch])`` (LazyAlgebra) and the operator-less
    ``apply([P,] ker, x, src[, len])`` (src/unidimensional.jl:33-56)."""
    args = list(args)
    P = Direct
    if args and _is_op(args[0]):
        P = args.pop(0)
    if args and isinstance(args[0], Kernel):
        from .unidimensional import apply_kernel
        return apply_kernel(P, *args)
    if not args or not isinstance(args[0], Mapping):
        raise TypeError("apply([P,] A, x[, scratch])")
    A = args.pop(0)
    if args and isinstance(args[-1], bool):
        args.pop()
    if len(args) == 2 and isinstance(args[0], (int, np.integer)):  # Tabulated with explicit dimension
        d, x = int(args[0]), args[1]
        y = A._vcreate(P, x, d=d)
        return A._apply(1, P, x, False, 0, y, d=d)
    (x,) = args
    y = A._vcreate(P, x)
    A._apply(1, P, x, False, 0, y)
    return y


def apply_b(*args):
    """``apply!``.  Accepted forms (argument order exactly as in the reference):

    * ``apply!(alpha, P, A, x, scratch, beta, y)`` and LazyAlgebra's 6-argument ``apply!(alpha, P, A, x, beta, y)``
    * ``apply!(alpha, P, A, d, x, scratch, beta, y)`` -- Tabulated without a stored dimension (src/tabulated.jl:246-257)
    * ``apply!(y, [P,] A, x)``
    * ``apply!(alpha, P, ker, x, src, beta, dst)`` / ``apply!(dst, [P,] ker, x, src)`` -- src/unidimensional.jl:58-245
    * ``apply!(alpha, P, ker1, x1, ker2, x2, src, beta, dst)`` / ``apply!(dst, [P,] ker1, x1, [ker2,] x2, src)`` -- src/separable.jl
    * ``apply!(alpha, P, ker1, [ker2,] R, src, beta, dst)`` / ``apply!(dst, [P,] ker1, [ker2,] R, src)`` -- src/nonseparable.jl
    * ``apply!(alpha, P, ker1, x1, ker2, x2, ker3, x3, src, beta, dst)`` -- three dimensions on the fly (``sep_otf_apply``)
    """
    from .nonseparable import AffineTransform2D, affine_apply
    from .separable import sep2d_otf_apply
    from .unidimensional import otf1d_apply
    a = list(args)
    if not a:
        raise TypeError("apply!: no arguments")
    if isinstance(a[0], (int, float, np.number)) and not isinstance(a[0], bool):
        alpha, P = a[0], a[1]
        if not _is_op(P):
            raise TypeError("apply!(alpha, P, ...): P must be Direct or Adjoint")
        rest = a[2:]
        beta, y = rest[-2], rest[-1]
        mid = rest[:-2]
    else:
        y = a[0]
        rest = a[1:]
        P = Direct
        if rest and _is_op(rest[0]):
            P = rest.pop(0)
        alpha, beta, mid = 1, 0, rest
    if not mid:
        raise TypeError("apply!: missing operator")
    head = mid[0]
    if isinstance(head, Mapping):
        tail = mid[1:]
        if tail and isinstance(tail[-1], bool):
            tail = tail[:-1]
        if len(tail) == 2 and isinstance(tail[0], (int, np.integer)):
            return head._apply(alpha, P, tail[1], False, beta, y, d=int(tail[0]))
        if len(tail) != 1:
            raise TypeError("apply!(alpha, P, A, [d,] x, [scratch,] beta, y)")
        return head._apply(alpha, P, tail[0], False, beta, y)
    if isinstance(head, Kernel):
        # affine forms contain an AffineTransform2D
        if any(isinstance(m, AffineTransform2D) for m in mid):
            if isinstance(mid[1], Kernel):
                ker1, ker2, R, src = mid
            else:
                (ker1, R, src), ker2 = mid, mid[0]
            return affine_apply(alpha, P, ker1, ker2, R, src, beta, y)
        if len(mid) == 3:
            ker, x, src = mid
            return otf1d_apply(alpha, P, ker, x, src, beta, y)
        if len(mid) == 5:
            ker1, x1, ker2, x2, src = mid
            return sep2d_otf_apply(alpha, P, ker1, x1, ker2, x2, src, beta, y)
        if len(mid) == 4:
            ker, x1, x2, src = mid
            return sep2d_otf_apply(alpha, P, ker, x1, ker, x2, src, beta, y)
        if len(mid) == 7 and all(isinstance(m, Kernel) for m in mid[0:6:2]):   # three dimensions (beyond the reference)
            from .separable import sep_otf_apply
            return sep_otf_apply(alpha, P, mid[0:6:2], mid[1:6:2], mid[6], beta, y)
    raise TypeError("apply!: no method matching these arguments")


apply_ = apply_b


def vdot(x, y) -> float:
    """LazyAlgebra ``vdot`` for real arrays, accumulated in Float64: on the device (``li_vdot``, fixed reduction
    order) when both arguments are device arrays, else on the host."""
    if ar.is_device(x) and ar.is_device(y) and ar.np_dtype_of(x) == ar.np_dtype_of(y) and ar.np_dtype_of(x).kind == "f" \
            and ar.is_colmajor(x) and ar.is_colmajor(y):
        from .solvers import vdot_device
        return vdot_device(x, y)
    return float(np.vdot(ar.to_host(x).astype(np.float64).ravel(order="F"),
                         ar.to_host(y).astype(np.float64).ravel(order="F")))


# ---- shared helpers for the operator modules ------------------------------------------------------
def real_view_factor(dtype) -> int:
    """Complex{T} data are handled as T data with a leading dimension of 2 (src/tabulated.jl:196)."""
    return 2 if np.dtype(dtype).kind == "c" else 1


def real_dtype(dtype) -> np.dtype:
    dtype = np.dtype(dtype)
    return np.dtype(np.float32) if dtype == np.complex64 else (np.dtype(np.float64) if dtype == np.complex128 else dtype)


def check_same_kind(x, y):
    if ar.is_device(x) != ar.is_device(y):
        raise _lib.ErrorException("source and destination must both be host arrays or both be device arrays")


def check_dst(y, shape, dtype, what="destination"):
    if not _is_array(y) or isinstance(y, (list, tuple)):
        raise TypeError(f"{what} must be an array")
    if ar.shape_of(y) != tuple(shape):
        raise _lib.DimensionMismatch(f"{what} has size {ar.shape_of(y)}, expected {tuple(shape)}")
    if ar.np_dtype_of(y) != np.dtype(dtype):
        raise TypeError(f"{what} has element type {ar.np_dtype_of(y)}, expected {np.dtype(dtype)}")
    if not ar.is_colmajor(y):
        raise _lib.ErrorException(f"{what} must be column-major (Julia layout)")

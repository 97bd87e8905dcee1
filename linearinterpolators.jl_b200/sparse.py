"""``SparseInterpolator`` and ``SparseUnidimensionalInterpolator`` -- host mirror of src/sparse.jl."""
from __future__ import annotations

import warnings

import numpy as np

from . import _arrays as ar
from . import _lib
from .kernels import Kernel
from .mapping import Adjoint, Direct, LinearMapping, Table, check_dst
from .tabulated import run_tab


class Range:
    """An ``AbstractRange`` (``first:step:last`` / ``range(first, last, length=n)``)."""

    def __init__(self, first, last=None, length=None, step=None):
        if isinstance(first, range):
            r = first
            self.first, self.step, self.length = float(r.start), float(r.step), len(r)
            self.last = float(r.start + (len(r) - 1) * r.step)
            return
        if length is None:
            step = 1.0 if step is None else float(step)
            length = int(np.floor((float(last) - float(first)) / step + 1e-12)) + 1
            last = float(first) + (length - 1) * step
        elif step is None:
            step = (float(last) - float(first)) / (int(length) - 1) if int(length) > 1 else 1.0
        self.first, self.last, self.step, self.length = float(first), float(last), float(step), int(length)

    def __len__(self):
        return self.length


def _as_range(grd) -> Range:
    if isinstance(grd, Range):
        return grd
    if isinstance(grd, range):
        return Range(grd)
    raise TypeError("grid must be a Range (or a Python range)")


def fractional_index(T, pos, grd=None):
    """src/sparse.jl:494-507: positions -> fractional (1-based) indices, evaluated in ``T`` on the host
    (each operation rounded separately, like the Julia closure)."""
    T = np.dtype(T).type
    p = np.asarray(pos).astype(T)
    if grd is None:
        return p
    g = _as_range(grd)
    c = (T(g.first) + T(g.last)) / T(2)
    q = T(1) / T(g.step)
    r = T(1 + g.length) / T(2)
    return q * (p - c) + r


class AbstractSparseInterpolator(LinearMapping):
    @property
    def C(self):
        """``coefficients(A)``: flat ``Vector{T}`` of length ``S*nrows`` (src/sparse.jl:43,521)."""
        return self.table.W.reshape(-1, order="F")

    @property
    def J(self):
        """``columns(A)``: flat ``Vector{Int}`` (1-based)."""
        return self.table.J.reshape(-1, order="F")


def _parse_T(args, T):
    args = list(args)
    if args and isinstance(args[0], (type, np.dtype)) and not isinstance(args[0], Kernel):
        # deprecated SparseInterpolator(T, ker, ...) form (src/sparse.jl:107,559; test/runtests.jl:134-136)
        warnings.warn("`Sparse*Interpolator(T, ker, ...)` is deprecated, use the `T=` keyword "
                      "(Julia: `Sparse*Interpolator{T}(ker, ...)`)", DeprecationWarning, stacklevel=3)
        T = args.pop(0)
    return args, T


class SparseInterpolator(AbstractSparseInterpolator):
    """``SparseInterpolator{T}(ker, pos, grd)`` / ``(ker, pos, len)`` (src/sparse.jl:91-138): vector in,
    array shaped like ``pos`` out."""

    def __init__(self, *args, T=None, variant=_lib.VARIANT_AUTO):
        args, T = _parse_T(args, T)
        if len(args) != 3 or not isinstance(args[0], Kernel):
            raise TypeError("SparseInterpolator(ker, pos, grd_or_len; T)")
        ker, pos, grd = args
        T = np.dtype(T) if T is not None else ker.T
        ker = ker.with_eltype(T)                                       # :109-110
        pos = np.asarray(ar.to_host(pos) if ar.is_device(pos) else pos)
        if isinstance(grd, (int, np.integer)):
            ncols, f = int(grd), fractional_index(T, pos)
        else:
            g = _as_range(grd)
            ncols, f = len(g), fractional_index(T, pos, g)
        if pos.size < 1 or min(pos.shape or (1,)) < 1:
            raise AssertionError("minimum(dims) >= 1")                 # :53
        self.table = Table.from_positions(ker, f, ncols)               # _sparsecoefs, :140-164
        self.dims = tuple(pos.shape)
        self.nrows, self.ncols, self.S, self.eltype = self.table.nrows, ncols, ker.S, ker.T
        self.variant = variant

    def output_size(self):
        return self.dims

    def input_size(self):
        return (self.ncols,)

    def rows(self):
        """``rows(A)`` (src/sparse.jl:69-85), 1-based."""
        return np.repeat(np.arange(1, self.nrows + 1, dtype=np.int64), self.S)

    def sparse(self):
        """``sparse(A)`` (src/sparse.jl:88-89) as a scipy CSR matrix (duplicates summed)."""
        import scipy.sparse as sp
        return sp.coo_matrix((self.C, (self.rows() - 1, self.J - 1)), shape=(self.nrows, self.ncols)).tocsr()

    def _vcreate(self, P, x):
        x = ar.to_colmajor(x, self.eltype)
        return ar.empty_like_kind(x, self.dims if P is Direct else (self.ncols,), self.eltype)  # :187-199

    def _apply(self, alpha, P, x, scratch, beta, y):
        x = ar.to_colmajor(x, self.eltype)
        inp, out = (x, y) if P is Direct else (y, x)
        if len(ar.shape_of(inp)) != 1 or ar.shape_of(inp)[0] != self.ncols:
            raise _lib.ErrorException(f"bad vector length (expecting {self.ncols}, got {ar.shape_of(inp)})")   # :176-177
        if ar.shape_of(out) != self.dims:
            raise _lib.ErrorException(f"bad output array size (expecting {self.dims}, got {ar.shape_of(out)})")  # :178-179
        check_dst(y, ar.shape_of(y), self.eltype)
        return run_tab(self.table, P is Adjoint, 1, 1, alpha, x, beta, y, self.variant)


class SparseUnidimensionalInterpolator(AbstractSparseInterpolator):
    """``SparseUnidimensionalInterpolator{T}(ker, d, pos, grd)`` / ``(ker, d, pos, len)``
    (src/sparse.jl:536-601): interpolates dimension ``d`` of an N-d array."""

    def __init__(self, *args, T=None, variant=_lib.VARIANT_AUTO):
        args, T = _parse_T(args, T)
        if len(args) != 4 or not isinstance(args[0], Kernel):
            raise TypeError("SparseUnidimensionalInterpolator(ker, d, pos, grd_or_len; T)")
        ker, d, pos, grd = args
        T = np.dtype(T) if T is not None else ker.T
        ker = ker.with_eltype(T)
        if isinstance(grd, (int, np.integer)):
            if grd < 1:
                raise _lib.ArgumentError("invalid dimension length")                  # :577
            grd = Range(range(1, int(grd) + 1))
        g = _as_range(grd)
        if not isinstance(d, (int, np.integer)):
            raise _lib.ArgumentError("invalid type for dimension of interpolation")   # :594
        if d < 1:
            raise _lib.ArgumentError("invalid dimension of interpolation")            # :595
        pos = np.asarray(ar.to_host(pos) if ar.is_device(pos) else pos)
        if pos.ndim != 1:
            raise TypeError("positions must be a vector")
        self.table = Table.from_positions(ker, fractional_index(T, pos, g), len(g))
        self.d = int(d)
        self.nrows, self.ncols, self.S, self.eltype = self.table.nrows, len(g), ker.S, ker.T
        self.variant = variant

    def size(self, i=None):
        if i is None:
            return (self.nrows, self.ncols)
        if i in (1, 2):
            return (self.nrows, self.ncols)[i - 1]
        raise _lib.ErrorException("out of bounds dimension")

    def _shape(self, P, x):
        xd = ar.shape_of(x)
        D = self.d
        nx, ny = (self.ncols, self.nrows) if P is Direct else (self.nrows, self.ncols)
        if not (1 <= D <= len(xd)):
            raise _lib.DimensionMismatch("out of range dimension of interpolation")   # :624, :806
        if xd[D - 1] != nx:
            raise _lib.DimensionMismatch(f"dimension {D} of `x` must be {nx}")        # :626, :646-651
        return xd[:D - 1] + (ny,) + xd[D:]

    def _vcreate(self, P, x):
        x = ar.to_colmajor(x, self.eltype)
        return ar.empty_like_kind(x, self._shape(P, x), self.eltype)

    def _apply(self, alpha, P, x, scratch, beta, y):
        x = ar.to_colmajor(x, self.eltype)
        yshape = self._shape(P, x)
        yd = ar.shape_of(y)
        if len(yd) != len(yshape):
            raise _lib.DimensionMismatch("`x` and `y` have incompatible dimensions")
        if yd[self.d - 1] != yshape[self.d - 1]:
            raise _lib.DimensionMismatch(f"dimension {self.d} of `y` must be {yshape[self.d - 1]}")
        if yd != yshape:
            raise _lib.DimensionMismatch("`x` and `y` have incompatible dimensions")  # :652-655
        check_dst(y, yshape, self.eltype)
        xd = ar.shape_of(x)
        npre = int(np.prod(xd[:self.d - 1], dtype=np.int64))
        npost = int(np.prod(xd[self.d:], dtype=np.int64))
        return run_tab(self.table, P is Adjoint, npre, npost, alpha, x, beta, y, self.variant)

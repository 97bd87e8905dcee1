"""``TabulatedInterpolator`` -- host mirror of src/tabulated.jl; tables and loops live on the device."""
from __future__ import annotations

import numpy as np

from . import _arrays as ar
from . import _lib
from .kernels import Kernel
from .mapping import (Adjoint, Direct, LinearMapping, Table, check_dst, check_same_kind, real_dtype,
                      real_view_factor)


def _is_type(x):
    return isinstance(x, (type, np.dtype)) and not isinstance(x, Kernel)


def run_tab(table: Table, adjoint: bool, npre: int, npost: int, alpha, src, beta, dst, variant=_lib.VARIANT_AUTO):
    """One ``li_tab_apply`` call on host (staged) or device (in place) arrays."""
    check_same_kind(src, dst)
    if ar.is_device(dst):
        _lib.check(_lib.lib.li_tab_apply(int(adjoint), table.handle, npre, npost, float(alpha), ar.data_ptr(src),
                                         float(beta), ar.data_ptr(dst), variant, ar.current_stream(dst)))
    else:
        stream, asynchronous = ar.host_stream_state()
        fn = _lib.lib.li_tab_apply_host_async if asynchronous else _lib.lib.li_tab_apply_host
        _lib.check(fn(int(adjoint), table.handle, npre, npost, float(alpha), ar.data_ptr(src), float(beta),
                      ar.data_ptr(dst), variant, stream))
    return dst


class TabulatedInterpolator(LinearMapping):
    """``TabulatedInterpolator([T,] [d = nothing,] ker, pos, nrows, ncols)`` /
    ``TabulatedInterpolator([T,] [d = nothing,] ker, x, ncols)`` (src/tabulated.jl:76-135).

    ``A.J`` / ``A.W`` (``columns(A)`` / ``coefficients(A)``) are materialised lazily on the host in the
    reference's layout: Int64 1-based ``S x nrows`` and ``T`` ``S x nrows``, column-major."""

    def __init__(self, *args, variant=_lib.VARIANT_AUTO):
        a = list(args)
        T = None
        d = None
        if a and _is_type(a[0]):
            T = np.dtype(a.pop(0))
        if a and (a[0] is None or isinstance(a[0], (int, np.integer))) and len(a) > 1 and isinstance(a[1], Kernel):
            d = a.pop(0)
            d = None if d is None else int(d)
        if not a or not isinstance(a[0], Kernel):
            raise TypeError("TabulatedInterpolator([T,] [d,] ker, pos, nrows, ncols) or ([T,] [d,] ker, x, ncols)")
        ker = a.pop(0)
        if len(a) == 3 and callable(a[0]):
            pos, nrows, ncols = a
            nrows, ncols = int(nrows), int(ncols)
            if nrows < 1:
                raise _lib.ErrorException("number of rows too small")           # src/tabulated.jl:83
            if ncols < 1:
                raise _lib.ErrorException("number of columns too small")        # :84
            if T is None:
                T = ker.T                                                      # float(T), :114
            x = np.array([pos(i) for i in range(1, nrows + 1)])
        elif len(a) == 2:
            x, ncols = a
            ncols = int(ncols)
            xdt = ar.np_dtype_of(x) if hasattr(x, "dtype") else np.asarray(x).dtype
            if not ar.is_device(x):
                x = np.asarray(x)
                if x.ndim != 1:
                    raise TypeError("positions must be a vector")
            if int(np.prod(ar.shape_of(x))) < 1:
                raise _lib.ErrorException("number of positions too small")      # :95
            if ncols < 1:
                raise _lib.ErrorException("number of columns too small")        # :96
            if T is None:
                T = np.result_type(np.promote_types(ker.T, xdt) if xdt.kind == "f" else ker.T, np.float32)  # :118-123
        else:
            raise TypeError("TabulatedInterpolator([T,] [d,] ker, pos, nrows, ncols) or ([T,] [d,] ker, x, ncols)")
        ker = ker.with_eltype(T)                                               # :103-108
        if ar.is_device(x) and ar.np_dtype_of(x) != ker.T:
            x = ar.to_host(x)
        self.table = Table.from_positions(ker, x, ncols)
        self.nrows, self.ncols, self.d = self.table.nrows, self.table.ncols, d
        self.eltype = ker.T
        self.S = ker.S
        self.variant = variant

    @classmethod
    def from_tables(cls, J, W, ncols, d=None):
        """Build from precomputed reference-layout tables (user-defined kernels)."""
        self = object.__new__(cls)
        self.table = Table.from_arrays(J, W, ncols)
        self.nrows, self.ncols, self.d = self.table.nrows, self.table.ncols, d
        self.eltype, self.S, self.variant = self.table.dtype, self.table.S, _lib.VARIANT_AUTO
        return self

    # accessors (src/tabulated.jl:40-44)
    @property
    def J(self):
        return self.table.J

    @property
    def W(self):
        return self.table.W

    def _dim(self, x, d):
        N = len(ar.shape_of(x))
        if d is None:
            d = self.d
        if d is None:
            if N == 1:
                return 1                                                       # :218-226, :259-270
            raise _lib.ErrorException("dimension to interpolate must be provided")  # :233, :281
        if not (1 <= d <= N):
            raise _lib.ErrorException(f"out of range dimension {d}")          # :310
        return d

    def _prep(self, x):
        xdt = ar.np_dtype_of(x) if hasattr(x, "dtype") else np.asarray(x).dtype
        dt = np.dtype(np.complex64 if self.eltype == np.float32 else np.complex128) if xdt.kind == "c" else self.eltype
        return ar.to_colmajor(x, dt), dt

    def _vcreate(self, P, x, d=None):
        d = self._dim(x, d)
        x, dt = self._prep(x)
        shape = list(ar.shape_of(x))
        shape[d - 1] = self.nrows if P is Direct else self.ncols              # :305-327
        return ar.empty_like_kind(x, shape, dt)

    def _apply(self, alpha, P, x, scratch, beta, y, d=None):
        d = self._dim(x, d)
        x, dt = self._prep(x)
        sshape = ar.shape_of(x)
        nin, nout = (self.ncols, self.nrows) if P is Direct else (self.nrows, self.ncols)
        dshape = ar.shape_of(y)
        if len(dshape) != len(sshape):
            raise _lib.DimensionMismatch("incompatible destination and source dimensions")
        if dshape[d - 1] != nout:
            raise _lib.DimensionMismatch(f"dimension {d} of destination must be {nout}")   # :340-341
        if sshape[d - 1] != nin:
            raise _lib.DimensionMismatch(f"dimension {d} of source must be {nin}")         # :342-343
        if dshape[:d - 1] != sshape[:d - 1] or dshape[d:] != sshape[d:]:
            raise _lib.DimensionMismatch("incompatible destination and source dimensions")  # :346-347
        check_dst(y, dshape, dt)
        npre = int(np.prod(sshape[:d - 1], dtype=np.int64)) * real_view_factor(dt)
        npost = int(np.prod(sshape[d:], dtype=np.int64))
        return run_tab(self.table, P is Adjoint, npre, npost, alpha, x, beta, y, self.variant)


def nrows(A):
    return A.nrows


def ncols(A):
    return A.ncols


def width(A):
    return A.S


def coefficients(A):
    """``coefficients(A)`` (src/tabulated.jl:43, src/sparse.jl:67,527)."""
    return A.W if isinstance(A, TabulatedInterpolator) else A.C


def columns(A):
    """``columns(A)`` (src/tabulated.jl:44, src/sparse.jl:68,528)."""
    return A.J

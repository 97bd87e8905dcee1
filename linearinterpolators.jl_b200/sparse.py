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

    def _promoted_table(self):
        """The same table with its Float32 weights converted (exactly) to Float64: the arithmetic type
        ``float(promote_type(Ta, Tx))`` of src/sparse.jl:214,250 when Float64 data meet a Float32 operator."""
        if getattr(self, "_table64", None) is None:
            self._table64 = Table.from_arrays(self.table.J, self.table.W.astype(np.float64), self.ncols)
        return self._table64

    def _apply_mixed(self, alpha, P, x, beta, y):
        """Host arrays whose element types differ from the operator's (src/sparse.jl:201-275 accepts ``Tx<:Real``,
        ``Ty<:AbstractFloat``): the sums run in ``T = float(promote_type(Ta, Tx))`` on the device, the result is
        combined with ``beta*y`` and converted to ``Ty`` on the host."""
        Tx, Ty, Ta = np.dtype(np.asarray(x).dtype), np.dtype(y.dtype), np.dtype(self.eltype)
        if Ty.kind != "f" or Tx.kind not in "fiub":
            raise TypeError("SparseInterpolator: source must be real, destination floating point")
        T = np.dtype(np.float64) if np.float64 in (Ta.type, Tx.type) else np.dtype(np.float32)   # (integers promote to Ta)
        table = self.table if T == Ta else self._promoted_table()
        xs = np.asfortranarray(np.asarray(x), dtype=T)
        tmp = np.empty(ar.shape_of(y), dtype=T, order="F")
        run_tab(table, P is Adjoint, 1, 1, alpha, xs, 0, tmp, self.variant)
        if float(beta) == 0.0:
            y[...] = tmp.astype(Ty)
        else:
            y[...] = (tmp.astype(np.result_type(T, Ty)) + Ty.type(beta) * y).astype(Ty)
        return y

    def _apply(self, alpha, P, x, scratch, beta, y):
        mixed = (not ar.is_device(x) and not ar.is_device(y) and isinstance(y, np.ndarray)
                 and (np.asarray(x).dtype != self.eltype or y.dtype != self.eltype) and np.asarray(x).dtype.kind in "fiub")
        x = np.asarray(x) if mixed else ar.to_colmajor(x, self.eltype)
        inp, out = (x, y) if P is Direct else (y, x)
        if len(ar.shape_of(inp)) != 1 or ar.shape_of(inp)[0] != self.ncols:
            raise _lib.ErrorException(f"bad vector length (expecting {self.ncols}, got {ar.shape_of(inp)})")   # :176-177
        if ar.shape_of(out) != self.dims:
            raise _lib.ErrorException(f"bad output array size (expecting {self.dims}, got {ar.shape_of(out)})")  # :178-179
        if mixed and float(alpha) != 0.0:
            return self._apply_mixed(alpha, P, x, beta, y)
        if mixed:                       # alpha == 0: vscale!(y, beta) only (src/sparse.jl:211,255)
            if float(beta) == 0.0:
                y[...] = 0
            else:
                y *= y.dtype.type(beta)
            return y
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


# ------------------------------------------------------------------------------------------------------
# Normal equations of a 1-D fit (src/sparse.jl:277-490), on the device.
RGL_EPS = 1e-9     # src/sparse.jl:358-359
RGL_MU = 0.0


def _normal_matrix(A: SparseInterpolator, w, device):
    if not isinstance(A, SparseInterpolator):
        raise TypeError("AtA / AtWA take a SparseInterpolator")
    n = A.ncols
    wd = None
    if w is not None:
        if ar.shape_of(w) != A.dims:
            raise AssertionError("size(wgt) == output_size(A)")                         # :335
        wd = w if ar.is_device(w) else ar.to_device(np.asfortranarray(np.asarray(w, dtype=A.eltype)))
        if ar.np_dtype_of(wd) != A.eltype:
            raise TypeError("weights must have the element type of the interpolator")
        device = device or ar.is_device(w)
    dst = ar.device_empty((n, n), A.eltype)
    _lib.check(_lib.lib.li_table_atwa(A.table.handle, ar.data_ptr(wd) if wd is not None else None, ar.data_ptr(dst),
                                      ar.current_stream(dst)))
    return dst if device else ar.to_host(dst)


def AtA(A: SparseInterpolator, device: bool = False):
    """``AtA(A)`` = ``A'*A`` as a dense ``ncols x ncols`` matrix (src/sparse.jl:289-321), built on the device
    (``li_table_atwa``); returned as a NumPy array unless ``device`` is set."""
    return _normal_matrix(A, None, device)


def AtWA(A: SparseInterpolator, w, device: bool = False):
    """``AtWA(A, w)`` = ``A'*diag(w)*A`` (src/sparse.jl:277-287, 323-356); a device array of weights gives a device
    result."""
    return _normal_matrix(A, w, device)


def regularize_(M, eps=RGL_EPS, mu=RGL_MU):
    """``regularize!(A, eps, mu)`` (src/sparse.jl:445-490): ``A += rho*(eps*I + mu*D'*D)``, in place, host or device
    square matrix."""
    shp = ar.shape_of(M)
    if len(shp) != 2 or shp[0] != shp[1]:
        raise AssertionError("size(A,1) == size(A,2)")
    if eps < 0 or mu < 0:
        raise AssertionError("eps >= 0 && mu >= 0")
    T = ar.np_dtype_of(M)
    if ar.is_device(M):
        status = _lib.lib.li_regularize(ar.li_dtype(T), ar.data_ptr(M), shp[0], float(eps), float(mu), ar.current_stream(M))
        if status == _lib.ERR_ARG and b"we have a problem" in _lib.lib.li_last_error():
            raise _lib.ErrorException("we have a problem!")
        _lib.check(status)
        return M
    if not (isinstance(M, np.ndarray) and M.flags.f_contiguous):
        raise TypeError("regularize_ works in place on a column-major NumPy array or a device array")
    d = ar.to_device(M)
    regularize_(d, eps, mu)
    M[...] = ar.to_host(d)
    return M


def regularize(M, eps=RGL_EPS, mu=RGL_MU):
    """``regularize(A, eps, mu)``: the regularized copy (src/sparse.jl:424-435)."""
    if ar.is_device(M):
        c = ar.device_empty(ar.shape_of(M), ar.np_dtype_of(M))
        _lib.check(_lib.lib.li_memcpy_d2d(ar.data_ptr(c), ar.data_ptr(M), int(np.prod(ar.shape_of(M))) * ar.np_dtype_of(M).itemsize,
                                          ar.current_stream(M)))
        return regularize_(c, eps, mu)
    return regularize_(np.array(M, order="F", copy=True), eps, mu)


def fit(A: SparseInterpolator, y, w=None, epsilon=RGL_EPS, mu=RGL_MU):
    """``fit(A, y [, w]; epsilon, mu)`` (src/sparse.jl:361-422): the minimiser of ``sum(w .* (A*x - y).^2)`` plus the
    regularization of ``regularize``.  The right-hand side ``A'*(w .* y)`` and the matrix ``A'*W*A`` are built on the
    device; the (small, dense) Cholesky solve runs on the host in Float64.  (The reference's own method calls
    ``cholfact!``, removed in Julia 1.0, and its unweighted method reads an undefined ``w``: this is the intended
    behaviour, fixed forward.)"""
    yh = np.asfortranarray(np.asarray(ar.to_host(y) if ar.is_device(y) else y, dtype=A.eltype))
    if yh.shape != A.dims:
        raise AssertionError("size(y) == output_size(A)")
    if w is not None:
        wh = np.asfortranarray(np.asarray(ar.to_host(w) if ar.is_device(w) else w, dtype=A.eltype))
        if wh.shape != yh.shape:
            raise AssertionError("size(w) == size(y)")
        rhs = np.asarray(A.adj * np.asfortranarray(wh * yh))
        lhs = AtWA(A, wh)
    else:
        rhs = np.asarray(A.adj * yh)
        lhs = AtA(A)
    lhs = np.asfortranarray(lhs)
    regularize_(lhs, epsilon, mu)
    import scipy.linalg as sl
    c = sl.cho_factor(np.asarray(lhs, dtype=np.float64))
    return sl.cho_solve(c, np.asarray(rhs, dtype=np.float64)).astype(A.eltype)

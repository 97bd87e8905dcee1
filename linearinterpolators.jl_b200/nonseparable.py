"""Non-separable 2-D interpolation through an affine transform (src/nonseparable.jl) and a value mirror of
``TwoDimensional.AffineTransform2D`` (only its six coefficients reach the hot path,
src/nonseparable.jl:162-166)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _arrays as ar
from . import _lib
from .kernels import Kernel
from .mapping import Adjoint, Direct, LinearMapping, check_dst, check_same_kind


class AffineTransform2D:
    """``(x', y') = (xx*x + xy*y + x0, yx*x + yy*y + y0)``; coefficients are kept in Float64 and rounded to the
    operator's type ``T`` at the boundary (``convert(AffineTransform2D{T}, R)``, src/nonseparable.jl:58)."""

    def __init__(self, xx=1.0, xy=0.0, x=0.0, yx=0.0, yy=1.0, y=0.0, T=np.float64):
        self.T = np.dtype(T)
        c = self.T.type
        self.xx, self.xy, self.x, self.yx, self.yy, self.y = (float(c(v)) for v in (xx, xy, x, yx, yy, y))

    def coefs(self):
        return (self.xx, self.xy, self.x, self.yx, self.yy, self.y)

    def __call__(self, x, y):
        return (self.xx * x + self.xy * y + self.x, self.yx * x + self.yy * y + self.y)

    def _new(self, *c):
        return AffineTransform2D(*c, T=self.T)

    # translate(A, t): translation of the input, A + t ; translate(t, A): of the output, t + A
    def __add__(self, t):
        tx, ty = t
        return self._new(self.xx, self.xy, self.xx * tx + self.xy * ty + self.x,
                         self.yx, self.yy, self.yx * tx + self.yy * ty + self.y)

    def __sub__(self, t):
        return self + (-t[0], -t[1])

    def __radd__(self, t):
        tx, ty = t
        return self._new(self.xx, self.xy, self.x + tx, self.yx, self.yy, self.y + ty)

    def __mul__(self, other):
        if isinstance(other, AffineTransform2D):  # composition A∘B
            A, B = self, other
            return self._new(A.xx * B.xx + A.xy * B.yx, A.xx * B.xy + A.xy * B.yy, A.xx * B.x + A.xy * B.y + A.x,
                             A.yx * B.xx + A.yy * B.yx, A.yx * B.xy + A.yy * B.yy, A.yx * B.x + A.yy * B.y + A.y)
        r = float(other)  # scale the input
        return self._new(self.xx * r, self.xy * r, self.x, self.yx * r, self.yy * r, self.y)

    def __rmul__(self, r):  # scale the output
        r = float(r)
        return self._new(*(r * v for v in self.coefs()))

    def __repr__(self):
        return "AffineTransform2D{%s}(%g, %g, %g, %g, %g, %g)" % ((self.T.name,) + self.coefs())


def rotate(a, b):
    """``rotate(A, theta)`` rotates the input of ``A`` (B = A.R); ``rotate(theta, A)`` its output (B = R.A)."""
    if isinstance(a, AffineTransform2D):
        A, th = a, float(b)
        cs, sn = math.cos(th), math.sin(th)
        return A._new(A.xx * cs + A.xy * sn, A.xy * cs - A.xx * sn, A.x, A.yx * cs + A.yy * sn, A.yy * cs - A.yx * sn,
                      A.y)
    th, A = float(a), b
    cs, sn = math.cos(th), math.sin(th)
    return A._new(cs * A.xx - sn * A.yx, cs * A.xy - sn * A.yy, cs * A.x - sn * A.y,
                  cs * A.yx + sn * A.xx, cs * A.yy + sn * A.xy, cs * A.y + sn * A.x)


def intercept(A: AffineTransform2D):
    """``(x, y)`` such that ``A(x, y) == (0, 0)``."""
    det = A.xx * A.yy - A.xy * A.yx
    return ((A.xy * A.y - A.yy * A.x) / det, (A.yx * A.x - A.xx * A.y) / det)


def affine_apply(alpha, P, ker1, ker2, R, src, beta, dst, variant=_lib.VARIANT_AUTO, nbatch_dims=0):
    """``apply!(alpha, P, ker1, ker2, R, src, beta, dst)`` (src/nonseparable.jl:147-264).  With
    ``nbatch_dims=1`` a trailing dimension of ``src``/``dst`` holds independent images."""
    T = ker1.T
    if ker2.T != T or ar.np_dtype_of(dst) != T:
        raise TypeError("kernels, transform and arrays must share one floating-point type")
    adjoint = P is Adjoint
    src = ar.to_colmajor(src, T)
    ss, ds = ar.shape_of(src), ar.shape_of(dst)
    if len(ss) != 2 + nbatch_dims or len(ds) != 2 + nbatch_dims:
        raise TypeError("source and destination must be 2-dimensional")
    nbatch = ss[2] if nbatch_dims else 1
    if nbatch_dims and ds[2] != nbatch:
        raise _lib.DimensionMismatch("source and destination hold different numbers of images")
    (n1, n2), (m1, m2) = (ss[:2], ds[:2]) if adjoint else (ds[:2], ss[:2])
    check_dst(dst, ds, T)
    check_same_kind(src, dst)
    Rc = (C.c_double * 6)(*R.coefs())
    dev = ar.is_device(dst)
    if dev:
        dsrc, ddst = src, dst
    else:
        dsrc = ar.to_device(src)
        # the reference's general branch reads dst even for beta == 0 (src/nonseparable.jl:195)
        ddst = ar.to_device(dst)
    _lib.check(_lib.lib.li_affine2d_apply(int(adjoint), C.byref(ker1.desc()), C.byref(ker2.desc()), Rc, n1, n2, m1,
                                          m2, nbatch, float(alpha), ar.data_ptr(dsrc), float(beta),
                                          ar.data_ptr(ddst), variant, 0, ar.current_stream(ddst)))
    if not dev:
        dst[...] = ar.to_host(ddst)
    return dst


class TwoDimensionalTransformInterpolator(LinearMapping):
    """``TwoDimensionalTransformInterpolator(rows, cols, ker1, [ker2,] R)`` (src/nonseparable.jl:33-59)."""

    def __init__(self, rows, cols, ker1: Kernel, *rest):
        if len(rest) == 1:
            ker2, R = ker1, rest[0]
        elif len(rest) == 2:
            ker2, R = rest
        else:
            raise TypeError("TwoDimensionalTransformInterpolator(rows, cols, ker1, [ker2,] R)")
        if not isinstance(R, AffineTransform2D):
            raise TypeError("R must be an AffineTransform2D")
        T = np.promote_types(np.promote_types(ker1.T, ker2.T), R.T)          # :54
        self.rows, self.cols = tuple(int(v) for v in rows), tuple(int(v) for v in cols)
        if len(self.rows) != 2 or len(self.cols) != 2:
            raise TypeError("rows and cols must be 2-tuples")
        self.ker1, self.ker2 = ker1.with_eltype(T), ker2.with_eltype(T)
        self.R = AffineTransform2D(*R.coefs(), T=T)
        self.eltype = np.dtype(T)

    def input_size(self):
        return self.cols

    def output_size(self):
        return self.rows

    def _vcreate(self, P, x):
        like = x if ar.is_device(x) else np.empty(0)
        return ar.empty_like_kind(like, self.rows if P is Direct else self.cols, self.eltype)  # :65-75

    def _apply(self, alpha, P, x, scratch, beta, y):
        x = ar.to_colmajor(x, self.eltype)
        xin, yout = (self.cols, self.rows) if P is Direct else (self.rows, self.cols)
        if ar.shape_of(x) != xin:      # @assert of src/nonseparable.jl:85,98 (kept under `python -O`)
            raise AssertionError("size(x) == A.cols" if P is Direct else "size(x) == A.rows")
        if ar.shape_of(y) != yout:     # :86,99
            raise AssertionError("size(y) == A.rows" if P is Direct else "size(y) == A.cols")
        return affine_apply(alpha, P, self.ker1, self.ker2, self.R, x, beta, y)

"""Shared helpers for the parity tests: the kernel zoo and oracle<->product kernel pairing."""
import numpy as np

import li_oracle as O

# (name, oracle family, S, params, li_b200 constructor name)
KERNELS = [
    ("Rectangular", O.RECTANGULAR, 1, (), "RectangularSpline"),
    ("Linear", O.LINEAR, 2, (), "LinearSpline"),
    ("Quadratic", O.QUADRATIC, 3, (), "QuadraticSpline"),
    ("Cubic", O.CUBIC, 4, (), "CubicSpline"),
    ("CatmullRom", O.CATMULLROM, 4, (), "CatmullRomSpline"),
    ("Cardinal(0.3)", O.CARDINALCUBIC, 4, (0.3,), "CardinalCubicSpline"),
    ("Keys(-0.4)", O.KEYS, 4, (-0.4,), "KeysSpline"),
    ("MitchellNetravali", O.MITCHELLNETRAVALI, 4, (1 / 3, 1 / 3), "MitchellNetravaliSpline"),
    ("Lanczos4", O.LANCZOS, 4, (), "LanczosKernel"),
    ("Lanczos6", O.LANCZOS, 6, (), "LanczosKernel"),
    ("Lanczos8", O.LANCZOS, 8, (), "LanczosKernel"),
]
KERNEL_IDS = [k[0] for k in KERNELS]
CONDITIONS = [("Flat", O.FLAT), ("SafeFlat", O.SAFEFLAT)]
DTYPES = [np.float32, np.float64]
# interpolated values: 1e-6 relative for Float32, 1e-12 for Float64 (BASELINE.json north_star)
RTOL = {np.dtype(np.float32): 1e-6, np.dtype(np.float64): 1e-12}


def oracle_kernel(entry, cond):
    _, fam, S, params, _ = entry
    return O.kernel(fam, S, cond, params)


def product_kernel(li, entry, cond, T):
    _, fam, S, params, ctor = entry
    B = li.Flat if cond == O.FLAT else li.SafeFlat
    cls = getattr(li, ctor)
    if fam == O.LANCZOS:
        return cls(T, S, B)
    return cls(T, *params, B)


# ------------------------------------------------------------------------------------------------
# Oracle results that carry their own error scale.
#
# interpolated values are not bit-defined even by the reference (its tap loops are @simd
# reductions), so parity is |a_i - b_i| <= rtol * sum_s |w_s * src_{j_s}| per element, rtol = 1e-6 (Float32) / 1e-12
# (Float64) -- the tolerance BASELINE.json states, on the natural forward-error scale of a banded dot product.  `RO`
# wraps the oracle: every result is a RefArray with `.bound` = the same operator applied with |w| to |src| (in
# Float64; composed operators chain their bounds) and `.terms` = the number of products summed into each element.
# Elements that sum more than TERMS0 products (columns that collect hundreds of clamped rows) get the bound scaled by
# terms / TERMS0: rounding errors of a sum grow with its length whatever the order of summation.
TERMS0 = 32


class RefArray(np.ndarray):
    bound = None
    terms = None

    def __array_finalize__(self, obj):      # (arithmetic on a reference gives a plain expectation: no bound)
        self.bound = None
        self.terms = None


def _wrap(arr, bound, terms):
    r = np.asarray(arr).view(RefArray)
    r.bound, r.terms = bound, terms
    return r


def _abs_in(x):
    b = getattr(x, "bound", None)
    return np.asfortranarray(np.abs(np.asarray(x, dtype=np.float64)) if b is None else b)


def _terms_in(x):
    t = getattr(x, "terms", None)
    return np.asfortranarray(np.ones(np.shape(x), dtype=np.float64) if t is None else t)


def _plain(x):
    return np.asarray(x).view(np.ndarray) if isinstance(x, np.ndarray) else x


class _RefOracle:
    """li_oracle with results that know their error bound (see above); everything else is passed through."""

    def __getattr__(self, name):
        return getattr(O, name)

    @staticmethod
    def _tab_bound(fn, J, W, ncols, d, src, alpha, beta, dabs, kw):
        """bound and number of terms of fn(J, W, ...) (dabs = |dst| before the call, or None)"""
        W64 = np.asfortranarray(np.abs(np.asarray(W, dtype=np.float64)))
        ones = np.asfortranarray((np.asarray(W) != 0).astype(np.float64))
        kw2 = {k: v for k, v in kw.items() if k != "skip_zero"}
        if fn is O.tab_adjoint:
            kw2["skip_zero"] = False
        hasd = dabs is not None
        bound = fn(J, W64, ncols, d, _abs_in(src), abs(alpha), abs(beta) if hasd else 0, dabs.copy(order="F") if hasd else None, **kw2)
        terms = fn(J, ones, ncols, d, _terms_in(src), 1 if alpha != 0 else 0, (1 if beta != 0 else 0) if hasd else 0,
                   np.asfortranarray(np.ones(dabs.shape)) if hasd else None, **kw2)
        return bound, terms

    @staticmethod
    def _dabs(dst):   # (the oracle updates dst in place: its old magnitude is taken first)
        return None if dst is None else np.asfortranarray(np.abs(np.asarray(dst, dtype=np.float64)))

    def _tab(self, fn, J, W, ncols, d, src, alpha, beta, dst, kw):
        dabs = self._dabs(dst)
        ref = fn(J, W, ncols, d, _plain(src), alpha, beta, None if dst is None else _plain(dst), **kw)
        bound, terms = self._tab_bound(fn, J, W, ncols, d, src, alpha, beta, dabs, kw)
        return _wrap(ref, bound, terms)

    def tab_direct(self, J, W, ncols, d, src, alpha=1, beta=0, dst=None, **kw):
        return self._tab(O.tab_direct, J, W, ncols, d, src, alpha, beta, dst, kw)

    def tab_adjoint(self, J, W, ncols, d, src, alpha=1, beta=0, dst=None, **kw):
        return self._tab(O.tab_adjoint, J, W, ncols, d, src, alpha, beta, dst, kw)

    def otf1d_direct(self, k, x, src, dtype, alpha=1, beta=0, dst=None):
        dabs = self._dabs(dst)
        ref = O.otf1d_direct(k, x, _plain(src), dtype, alpha, beta, None if dst is None else _plain(dst))
        xs = np.asarray(x, dtype=dtype)
        J, W = O.maketable(k, xs.reshape(-1, order="F"), int(np.size(src)), dtype)
        bound, terms = self._tab_bound(O.tab_direct, J, W, int(np.size(src)), 1, np.asarray(src).reshape(-1), alpha, beta,
                                       None if dabs is None else dabs.reshape(-1, order="F"), {})
        return _wrap(ref, bound.reshape(ref.shape, order="F"), terms.reshape(ref.shape, order="F"))

    def otf1d_adjoint(self, k, x, src, length, dtype, alpha=1, beta=0, dst=None):
        dabs = self._dabs(dst)
        ref = O.otf1d_adjoint(k, x, _plain(src), length, dtype, alpha, beta, None if dst is None else _plain(dst))
        xs = np.asarray(x, dtype=dtype)
        J, W = O.maketable(k, xs.reshape(-1, order="F"), length, dtype)
        bound, terms = self._tab_bound(O.tab_adjoint, J, W, length, 1, np.asarray(src).reshape(-1, order="F"), alpha, beta, dabs, {})
        return _wrap(ref, bound, terms)

    def sep2d(self, adjoint, k1, x1, k2, x2, src, other_shape, dtype, alpha=1, beta=0, dst=None):
        dabs = self._dabs(dst)
        ref = O.sep2d(adjoint, k1, x1, k2, x2, _plain(src), other_shape, dtype, alpha, beta, None if dst is None else _plain(dst))
        m1, m2 = other_shape if adjoint else np.shape(src)
        J1, W1 = O.maketable(k1, np.asarray(x1, dtype=dtype), m1, dtype)
        J2, W2 = O.maketable(k2, np.asarray(x2, dtype=dtype), m2, dtype)
        if adjoint:
            b1, t1 = self._tab_bound(O.tab_adjoint, J1, W1, m1, 1, src, 1, 0, None, {})
            t = _wrap(b1, b1, t1)
            bound, terms = self._tab_bound(O.tab_adjoint, J2, W2, m2, 2, t, alpha, beta, dabs, {})
        else:
            b1, t1 = self._tab_bound(O.tab_direct, J2, W2, m2, 2, src, 1, 0, None, {})
            t = _wrap(b1, b1, t1)
            bound, terms = self._tab_bound(O.tab_direct, J1, W1, m1, 1, t, alpha, beta, dabs, {})
        return _wrap(ref, bound, terms)

    def affine2d(self, adjoint, k1, k2, R, rows, cols, src, dtype, alpha=1, beta=0, dst=None, **kw):
        dabs = self._dabs(dst)
        ref = O.affine2d(adjoint, k1, k2, R, rows, cols, _plain(src), dtype, alpha, beta, None if dst is None else _plain(dst), **kw)
        # the same loops with |w| on |src| (bound) and with (w != 0) on ones (number of terms), in Float64
        R64 = np.asarray(np.asarray(R, dtype=dtype), dtype=np.float64)
        kw2 = {k: v for k, v in kw.items()} if not adjoint else {}
        if not adjoint:
            kw2.setdefault("nthreads", O.max_threads())
        zeros = lambda: np.zeros(cols if adjoint else rows, dtype=np.float64, order="F")   # (the general branch reads dst)
        with O.weight_mode(1):
            bound = O.affine2d(adjoint, k1, k2, R64, rows, cols, _abs_in(src), np.float64, abs(alpha), 0, zeros(), **kw2)
        with O.weight_mode(2):
            terms = O.affine2d(adjoint, k1, k2, R64, rows, cols, _terms_in(src), np.float64, 1, 0, zeros(), **kw2)
        if dabs is not None and beta != 0:
            bound += abs(beta) * dabs
            terms += 1
        return _wrap(ref, bound, terms)


RO = _RefOracle()


def _strict(a, b, dtype, what=""):
    bound, terms = b.bound, b.terms
    scale = np.maximum(1.0, (terms if terms is not None else 0.0) / TERMS0)
    return assert_within_bound(a, np.asarray(b), bound * scale, dtype, what)


def assert_close(a, b, dtype, scale=None, factor=1.0):
    """Parity check.  A reference from `RO` is checked per element against its own bound at the stated tolerance
    (`factor` / `scale` are ignored); any other expectation with ||a - b||_inf <= factor * rtol * ||b||_inf."""
    if getattr(b, "bound", None) is not None:
        _strict(a, b, dtype)
        return
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    s = float(np.max(np.abs(b))) if scale is None else scale
    tol = RTOL[np.dtype(dtype)] * factor * max(s, np.finfo(np.float64).tiny)
    err = float(np.max(np.abs(a - b))) if a.size else 0.0
    assert err <= tol, f"max abs err {err:.3e} > {tol:.3e} (scale {s:.3e})"


def hostile_positions(ncols, T, rng, n=257):
    """Positions exercising every branch of getcoefs: far out of range, exact integers, half-integers
    (round-half-even), +-0, subnormals, the SafeFlat thresholds, huge-but-representable values."""
    T = np.dtype(T).type
    fi = np.finfo(T)
    base = [-(2.0 ** 40), -1e6, -3.5, -2.5, -1.5, -1.0, -0.5, -0.0, 0.0, float(fi.tiny) / 4, -float(fi.tiny) / 4,
            0.5, 1.0, 1.5, 2.0, 2.5, 3.5, ncols - 0.5, ncols, ncols + 0.5, ncols + 1, ncols + 1.5, ncols + 2,
            ncols + 3, ncols + 4.5, 1e6, 2.0 ** 40, float(np.nextafter(T(ncols + 1), T(np.inf))),
            float(np.nextafter(T(0), T(-1))), float(np.nextafter(T(1), T(0)))]
    extra = rng.uniform(-4, ncols + 5, n - len(base))
    return np.concatenate([np.array(base), extra]).astype(T)


def assert_within_bound(a, b, bound, dtype, what=""):
    """Strict per-element parity check of SURVEY.md Appendix F: |a_i - b_i| <= rtol * sum_s |w_s * src_{j_s}|, the
    natural forward-error scale of a banded dot product (`bound` = the same operator applied with |w| to |src|).
    rtol = 1e-6 (Float32) / 1e-12 (Float64), no extra factor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bound = np.asarray(bound, dtype=np.float64)
    assert a.shape == b.shape == bound.shape, (a.shape, b.shape, bound.shape)
    tol = RTOL[np.dtype(dtype)] * bound + np.finfo(np.float64).tiny
    bad = np.abs(a - b) > tol
    if bad.any():
        k = np.unravel_index(np.argmax(np.abs(a - b) / tol), a.shape)
        raise AssertionError(f"{what}: {int(bad.sum())} elements out of bound; worst at {k}: |{a[k]} - {b[k]}| = "
                             f"{abs(a[k] - b[k]):.3e} > {tol[k]:.3e}")
    return float(np.max(np.abs(a - b) / tol)) if a.size else 0.0


def sep2d_oracle(J1, W1, m1, J2, W2, m2, x, adjoint):
    """A1*A2 (or its adjoint) with the reference's per-dimension loops and order (A2 first / A1' first)."""
    if not adjoint:
        return RO.tab_direct(J1, W1, m1, 1, RO.tab_direct(J2, W2, m2, 2, x))
    return RO.tab_adjoint(J2, W2, m2, 2, RO.tab_adjoint(J1, W1, m1, 1, x))

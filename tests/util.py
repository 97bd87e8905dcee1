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


def assert_close(a, b, dtype, scale=None, factor=1.0):
    """Headline parity check of SURVEY.md Appendix F: ||a - b||_inf <= rtol * ||b||_inf."""
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
        return O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, x))
    return O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, x))

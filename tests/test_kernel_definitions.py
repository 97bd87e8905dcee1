"""Known answers for the weights: the oracle's (and, bit for bit, the device's) `getweights` against the published
MATHEMATICAL definitions of the kernels, evaluated independently in exact rational arithmetic (`fractions`) or with
mpmath (Lanczos).  InterpolationKernels 0.1.x is not on disk, so the last-bit operation order stays unpinned
(DESIGN.md section 3); what this pins is that every family, parameter convention and tap order is the published one:
w_s(t) = ker(t - s + c) for the S taps around the position, and the `*Prime` kernels are their derivatives."""
from fractions import Fraction as F

import numpy as np
import pytest

import li_oracle as O


def k_linear(x):
    return max(F(0), 1 - abs(x))


def k_quadratic(x):            # quadratic B-spline
    a = abs(x)
    if a <= F(1, 2):
        return F(3, 4) - a * a
    if a <= F(3, 2):
        return (a - F(3, 2)) ** 2 / 2
    return F(0)


def k_cubic(x):                # cubic B-spline
    a = abs(x)
    if a <= 1:
        return F(2, 3) - a * a + a ** 3 / 2
    if a <= 2:
        return (2 - a) ** 3 / 6
    return F(0)


def k_keys(a):                 # Keys' family of cardinal cubics; a = -1/2: Catmull-Rom
    def k(x):
        t = abs(x)
        if t <= 1:
            return (a + 2) * t ** 3 - (a + 3) * t ** 2 + 1
        if t <= 2:
            return a * t ** 3 - 5 * a * t ** 2 + 8 * a * t - 4 * a
        return F(0)
    return k


def k_mitchell(b, c):          # Mitchell & Netravali (1988)
    def k(x):
        t = abs(x)
        if t <= 1:
            return ((12 - 9 * b - 6 * c) * t ** 3 + (-18 + 12 * b + 6 * c) * t ** 2 + (6 - 2 * b)) / 6
        if t <= 2:
            return ((-b - 6 * c) * t ** 3 + (6 * b + 30 * c) * t ** 2 + (-12 * b - 48 * c) * t + (8 * b + 24 * c)) / 6
        return F(0)
    return k


CASES = [
    ("Linear", O.LINEAR, 2, (), k_linear),
    ("Quadratic", O.QUADRATIC, 3, (), k_quadratic),
    ("Cubic", O.CUBIC, 4, (), k_cubic),
    ("CatmullRom", O.CATMULLROM, 4, (), k_keys(F(-1, 2))),
    ("Cardinal(0.25)", O.CARDINALCUBIC, 4, (0.25,), k_keys((F(1, 4) - 1) / 2)),     # tension c <-> Keys a = (c - 1)/2
    ("Keys(-0.75)", O.KEYS, 4, (-0.75,), k_keys(F(-3, 4))),
    ("MitchellNetravali(1/4,3/8)", O.MITCHELLNETRAVALI, 4, (0.25, 0.375), k_mitchell(F(1, 4), F(3, 8))),
]


def _offsets(S):
    c = (S + 1) // 2          # tap s (1-based) sits at distance t + (c - s) from the position
    return [c - s for s in range(1, S + 1)]


@pytest.mark.parametrize("name,fam,S,params,kfun", CASES, ids=[c[0] for c in CASES])
def test_weights_are_the_published_kernel(name, fam, S, params, kfun):
    ker, kp = O.kernel(fam, S, O.FLAT, params), O.kernel(fam, S, O.FLAT, params, prime=True)
    ts = [F(n, 64) for n in (range(0, 64) if S % 2 == 0 else range(-32, 33))]      # exactly representable offsets
    h = F(1, 2 ** 20)
    for t in ts:
        exact = [kfun(t + d) for d in _offsets(S)]
        got = O.getweights(ker, float(t), np.float64)
        assert max(abs(float(e) - g) for e, g in zip(exact, got)) <= 2e-15, (name, t)      # (a few ulps of the largest coefficient)
        got32 = O.getweights(ker, np.float32(float(t)), np.float32)
        assert max(abs(float(e) - float(g)) for e, g in zip(exact, got32)) <= 1e-6, (name, t)
        if S % 2 == 0 and t == 0 or abs(t) == F(1, 2) and S % 2:
            continue           # (a knot of the piecewise polynomial: one-sided derivatives)
        deriv = [(kfun(t + d + h) - kfun(t + d - h)) / (2 * h) for d in _offsets(S)]
        gotp = O.getweights(kp, float(t), np.float64)
        assert max(abs(float(e) - g) for e, g in zip(deriv, gotp)) <= 2e-11, (name, t)   # (cubic: h^2 k'''/6 < 1e-12)
    # known values
    if name == "CatmullRom":
        assert list(O.getweights(ker, 0.5, np.float64)) == [-1 / 16, 9 / 16, 9 / 16, -1 / 16]
    if name == "Cubic":
        np.testing.assert_allclose(O.getweights(ker, 0.0, np.float64), [1 / 6, 2 / 3, 1 / 6, 0], atol=1e-16)


@pytest.mark.parametrize("S", [2, 4, 6, 8])
def test_lanczos_weights(S):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    a = S // 2
    ker, kp = O.kernel(O.LANCZOS, S), O.kernel(O.LANCZOS, S, prime=True)

    def k(x):
        x = mp.mpf(x)
        if abs(x) >= a:
            return mp.mpf(0)
        return mp.sinc(mp.pi * x) * mp.sinc(mp.pi * x / a)      # mpmath's sinc(x) = sin(x)/x
    for n in range(0, 64):
        t = n / 64.0
        exact = [k(t + d) for d in _offsets(S)]
        got = O.getweights(ker, t, np.float64)
        assert max(abs(float(e) - g) for e, g in zip(exact, got)) <= 2e-15, (S, t)
        if n == 0:
            continue
        deriv = [mp.diff(k, t + d) if abs(t + d) < a else mp.mpf(0) for d in _offsets(S)]
        gotp = O.getweights(kp, t, np.float64)
        assert max(abs(float(e) - g) for e, g in zip(deriv, gotp)) <= 1e-12 / t, (S, t)

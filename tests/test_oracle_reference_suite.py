"""The reference's own test-suite (test/runtests.jl) restated against the CPU oracle.

These are the only known answers the reference ships for the hot path;
they pin the oracle's reading of src/boundaries.jl, the loop nests and the alpha/beta algebra.
Weights stay "parity unpinned" at the last bit (InterpolationKernels is not vendored).
"""
import numpy as np
import pytest

import li_oracle as O

# kernels = (RectangularSpline(), LinearSpline(), QuadraticSpline(), CubicSpline(), CatmullRomSpline(),
#            KeysSpline(-0.4), MitchellNetravaliSpline(), LanczosKernel(4), LanczosKernel(6))  test/runtests.jl:76-78
KERNELS = [
    ("Rectangular", O.RECTANGULAR, 1, ()),
    ("Linear", O.LINEAR, 2, ()),
    ("Quadratic", O.QUADRATIC, 3, ()),
    ("Cubic", O.CUBIC, 4, ()),
    ("CatmullRom", O.CATMULLROM, 4, ()),
    ("Keys(-0.4)", O.KEYS, 4, (-0.4,)),
    ("MitchellNetravali", O.MITCHELLNETRAVALI, 4, (1 / 3, 1 / 3)),
    ("Lanczos4", O.LANCZOS, 4, ()),
    ("Lanczos6", O.LANCZOS, 6, ()),
]
CONDITIONS = [("Flat", O.FLAT), ("SafeFlat", O.SAFEFLAT)]


def _data():
    # test/runtests.jl:83-88
    sub = 3
    x = np.linspace(-1.0, 1.0, 100 * sub + 1)
    xsub = x[::sub]
    y = np.cos(2 * x * (x + 2))
    ysub = y[::sub]
    t = np.linspace(1.0, float(len(xsub)), len(x))
    return x, xsub, y, ysub, t


def distance(a, b):
    return float(np.mean(np.abs(np.asarray(a) - np.asarray(b))))


def test_unidimensional_known_answer_and_alpha_beta():
    # test/runtests.jl:15-49.  LinearSpline(Float32), Int source, Float64 positions.
    T = np.float32
    ker = O.kernel(O.LINEAR, 2, O.FLAT)
    src = np.array([0, 0, 1, 0, 0])
    x = np.linspace(-1, 7, 81)
    out = O.otf1d_direct(ker, x, src, T)
    assert out.dtype == T
    assert out[x == 2.5][0] == 0.5                                     # :20
    out2 = O.otf1d_direct(ker, x, src, T, alpha=2, beta=0, dst=np.empty_like(out))
    assert np.array_equal(out2, 2 * out)                               # :26-27
    out2 = O.otf1d_direct(ker, x, src, T, alpha=1, beta=2, dst=np.ones_like(out))
    assert np.array_equal(out2, out + 2)                               # :29-31
    out2 = O.otf1d_direct(ker, x, src, T, alpha=0, beta=2, dst=np.full_like(out, 3))
    assert np.all(out2 == 6)                                           # :33-35
    adj = O.otf1d_adjoint(ker, x, out, len(src), T)
    adj2 = O.otf1d_adjoint(ker, x, out, len(src), T, alpha=2, beta=0, dst=np.empty_like(adj))
    assert np.array_equal(adj2, 2 * adj)                               # :38-40
    adj2 = O.otf1d_adjoint(ker, x, out, len(src), T, alpha=1, beta=2, dst=np.ones_like(adj))
    np.testing.assert_allclose(adj2, adj + 2, rtol=1e-6)               # :42-44  (≈)
    adj2 = O.otf1d_adjoint(ker, x, out, len(src), T, alpha=0, beta=2, dst=np.full_like(adj, 3))
    assert np.all(adj2 == 6)                                           # :46-48


def test_getcoefs_examples():
    # SURVEY.md Appendix A.7 (derived by hand from src/boundaries.jl) + test/runtests.jl:90-98 positions
    T = np.float64
    cr = O.kernel(O.CATMULLROM, 4, O.FLAT)
    j, w = O.getcoefs(cr, 5, 0.5, T)
    assert j.tolist() == [1, 1, 1, 2]
    assert w.tolist() == [-1 / 16, 9 / 16, 9 / 16, -1 / 16]
    crs = O.kernel(O.CATMULLROM, 4, O.SAFEFLAT)
    j, w = O.getcoefs(crs, 5, -1e-300, T)
    assert j.tolist() == [1, 1, 1, 1] and w.tolist() == [0, 0, 0, 1]
    j, w = O.getcoefs(crs, 5, 6.0, T)
    assert j.tolist() == [5, 5, 5, 5] and w.tolist() == [0, 1, 0, 0]
    j, w = O.getcoefs(crs, 5, 6.0000001, T)
    assert j.tolist() == [5, 5, 5, 5] and w.tolist() == [1, 0, 0, 0]
    q = O.kernel(O.QUADRATIC, 3, O.FLAT)
    assert O.getcoefs(q, 10, 2.5, T)[0].tolist() == [1, 2, 3]          # round half to even
    assert O.getcoefs(q, 10, 3.5, T)[0].tolist() == [3, 4, 5]
    for x in (0, 0.1, -1.2):                                           # len = 4 as in :93
        j, w = O.getcoefs(cr, 4, x, T)
        assert j.min() >= 1 and j.max() <= 4
        assert abs(w.sum() - 1) < 1e-15


def test_limits():
    # src/boundaries.jl:40-42
    for T in (np.float32, np.float64):
        inf, sup = O.limits(O.kernel(O.LANCZOS, 4, O.SAFEFLAT), 4096, T)
        assert inf == np.nextafter(T(0), T(-1)) and inf < 0
        assert sup == np.nextafter(T(4097), T(np.inf))
        inf, sup = O.limits(O.kernel(O.RECTANGULAR, 1, O.SAFEFLAT), 7, T)
        assert inf == np.nextafter(T(1.5), T(-1)) and sup == np.nextafter(T(6.5), T(np.inf))
        inf, sup = O.limits(O.kernel(O.LANCZOS, 6, O.SAFEFLAT), 7, T)
        assert inf == np.nextafter(T(-1), T(-2)) and sup == np.nextafter(T(9), T(np.inf))


def test_inexact_error():
    lin = O.kernel(O.LINEAR, 2, O.FLAT)
    for bad in (np.nan, 1e19, -1e19, np.inf):
        with pytest.raises(O.InexactError):
            O.getcoefs(lin, 5, bad, np.float64)
    lins = O.kernel(O.LINEAR, 2, O.SAFEFLAT)
    assert O.getcoefs(lins, 5, 1e300, np.float64)[0].tolist() == [5, 5]   # SafeFlat saturates instead
    assert O.getcoefs(lins, 5, -np.inf, np.float64)[0].tolist() == [1, 1]
    with pytest.raises(O.InexactError):
        O.getcoefs(lins, 5, np.nan, np.float64)


@pytest.mark.parametrize("kname,fam,S,params", KERNELS)
@pytest.mark.parametrize("cname,cond", CONDITIONS)
def test_tabulated_and_sparse_interpolators(kname, fam, S, params, cname, cond):
    # test/runtests.jl:100-137
    T = np.float64
    x, xsub, y, ysub, t = _data()
    tol = 0.02 if fam == O.RECTANGULAR else 0.006
    ker = O.kernel(fam, S, cond, params)
    J, W = O.maketable(ker, t, len(xsub), T)
    assert J.shape == (S, len(t)) and J.min() >= 1 and J.max() <= len(xsub)
    out = O.tab_direct(J, W, len(xsub), 1, ysub)
    assert distance(out, y) <= tol                                     # :115 / :131
    otf = O.otf1d_direct(ker, t, ysub, T)
    assert distance(out, otf) <= 1e-15                                 # :116 / :132
    # the sparse operator holds the same bytes (src/sparse.jl:148-160): C == vec(W), J == vec(J)
    out_s = O.tab_direct(J.reshape(S, -1, order="F"), W, len(xsub), 1, ysub)
    assert distance(out_s, out) <= 1e-15                               # :130


def test_two_dimensional_transform_adjoint_identity():
    # test/runtests.jl:139-158.  R = c + rotate(AffineTransform2D{Float64}() - c, 0.2)
    rng = np.random.default_rng(1234)
    rows, cols = (11, 16), (10, 15)
    cx, cy, th = 5.6, 6.4, 0.2
    # rotation by th about c: v -> c + Rot(th)*(v - c); only the six coefficients cross into the hot path
    cs, sn = np.cos(th), np.sin(th)
    R = np.array([cs, -sn, cx - cs * cx + sn * cy, sn, cs, cy - sn * cx - cs * cy])
    kerlist = [(O.RECTANGULAR, 1), (O.LINEAR, 2), (O.CATMULLROM, 4)]
    for f1, s1 in kerlist:
        for f2, s2 in kerlist:
            k1, k2 = O.kernel(f1, s1), O.kernel(f2, s2)
            x = np.asfortranarray(rng.standard_normal(cols))
            y = np.asfortranarray(rng.standard_normal(rows))
            Ax = O.affine2d(False, k1, k2, R, rows, cols, x, np.float64)
            Aty = O.affine2d(True, k1, k2, R, rows, cols, y, np.float64)
            np.testing.assert_allclose(np.vdot(Ax, y), np.vdot(x, Aty), rtol=1e-12)  # :155


def test_sep2d_matches_composition_of_tables():
    # src/separable.jl:56-162 is untested by the reference; pin it to the table-backed composition
    # A1*(A2*x) (test/benchmarks.jl:27-29), which must agree to rounding.
    rng = np.random.default_rng(7)
    T = np.float64
    m1, m2, n1, n2 = 13, 9, 31, 22
    k1 = O.kernel(O.CATMULLROM, 4, O.FLAT)
    k2 = O.kernel(O.LANCZOS, 4, O.SAFEFLAT)
    x1 = np.linspace(-1.0, m1 + 2.0, n1)
    x2 = np.linspace(0.5, m2 + 0.5, n2)
    src = np.asfortranarray(rng.standard_normal((m1, m2)))
    J1, W1 = O.maketable(k1, x1, m1, T)
    J2, W2 = O.maketable(k2, x2, m2, T)
    comp = O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, src))
    fused = O.sep2d(False, k1, x1, k2, x2, src, (n1, n2), T)
    np.testing.assert_allclose(fused, comp, rtol=0, atol=1e-13)
    y = np.asfortranarray(rng.standard_normal((n1, n2)))
    comp_t = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, y))
    fused_t = O.sep2d(True, k1, x1, k2, x2, y, (m1, m2), T)
    np.testing.assert_allclose(fused_t, comp_t, rtol=0, atol=1e-13)
    np.testing.assert_allclose(np.vdot(fused, y), np.vdot(src, fused_t), rtol=1e-12)


def test_tab_nd_dims_and_adjoint_identity():
    rng = np.random.default_rng(3)
    for T, rtol in ((np.float32, 2e-5), (np.float64, 1e-12)):
        k = O.kernel(O.CUBIC, 4, O.FLAT)
        shape = (5, 7, 4)
        for d in (1, 2, 3):
            ncols = shape[d - 1]
            xs = rng.uniform(-1, ncols + 2, 11)
            J, W = O.maketable(k, xs, ncols, T)
            src = np.asfortranarray(rng.standard_normal(shape).astype(T))
            out = O.tab_direct(J, W, ncols, d, src)
            ref = np.moveaxis(np.tensordot(_dense(J, W, ncols), src.astype(np.float64), axes=([1], [d - 1])), 0, d - 1)
            np.testing.assert_allclose(out, ref, rtol=0, atol=rtol * 10)
            y = np.asfortranarray(rng.standard_normal(out.shape).astype(T))
            back = O.tab_adjoint(J, W, ncols, d, y)
            np.testing.assert_allclose(np.vdot(out.astype(np.float64), y.astype(np.float64)),
                                       np.vdot(src.astype(np.float64), back.astype(np.float64)), rtol=rtol * 50)


def _dense(J, W, ncols):
    S, nrows = W.shape
    A = np.zeros((nrows, ncols))
    for i in range(nrows):
        for s in range(S):
            A[i, J[s, i] - 1] += W[s, i]
    return A


def test_sin_restatement_accuracy():
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 120
    rng = np.random.default_rng(0)
    for x in rng.uniform(-26, 26, 500):
        s = float(O.sin(x))
        t = mp.sin(mp.mpf(float(x)))
        assert abs(mp.mpf(s) - t) <= mp.mpf(float(np.spacing(abs(s))))          # < 1 ulp, like Base.sin
    for x in rng.uniform(-26, 26, 500).astype(np.float32):
        s = O.sin(x, np.float32)
        t = mp.sin(mp.mpf(float(x)))
        assert abs(mp.mpf(float(s)) - t) <= mp.mpf(float(np.spacing(np.float32(abs(s))))) * 0.51

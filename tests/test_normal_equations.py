"""AtA / AtWA / regularize / fit (src/sparse.jl:277-490).  CPU part: the oracle's loops against dense NumPy algebra.
GPU part: li_table_atwa / li_regularize bit-for-bit against the oracle (same sums in the same order), `fit` against a
NumPy least-squares solve."""
import numpy as np
import pytest

import li_oracle as O
from util import DTYPES


def _table(T, nrows=173, ncols=29, seed=3):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(-0.5, ncols + 1.5, nrows)).astype(T)
    x[::7] = rng.uniform(1, ncols, x[::7].shape).astype(T)        # not monotone
    return x, O.maketable(O.kernel(O.CATMULLROM, 4), x, ncols, T)


@pytest.mark.parametrize("T", DTYPES)
def test_oracle_normal_equations(T):
    x, (J, W) = _table(T)
    ncols = 29
    S, nrows = W.shape
    A = np.zeros((nrows, ncols))
    for i in range(nrows):
        for s in range(S):
            A[i, J[s, i] - 1] += W[s, i]
    w = np.random.default_rng(4).uniform(0, 2, nrows).astype(T)
    tol = 1e-4 if T == np.float32 else 1e-12
    np.testing.assert_allclose(O.AtWA(J, W, ncols), A.T @ A, atol=tol * nrows)
    np.testing.assert_allclose(O.AtWA(J, W, ncols, w), A.T @ (w[:, None] * A), atol=tol * nrows)
    M = O.AtWA(J, W, ncols)
    R = O.regularize(M, 1e-3, 0.25)
    rho = M.diagonal().max()
    D = np.diff(np.eye(ncols), axis=0)
    np.testing.assert_allclose(R, M + rho * (1e-3 * np.eye(ncols) + 0.25 * D.T @ D), atol=tol * rho * 4)
    assert np.array_equal(O.regularize(M, 0, 0), M)
    with pytest.raises(RuntimeError):
        O.regularize(np.zeros((3, 3), dtype=T), 1e-3, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("T", DTYPES)
def test_device_normal_equations_bit_exact(T):
    import li_b200 as li
    x, (J, W) = _table(T)
    ncols = 29
    A = li.SparseInterpolator(li.CatmullRomSpline(T), x, ncols)
    U = np.uint32 if T == np.float32 else np.uint64
    got = li.AtA(A)
    assert got.shape == (ncols, ncols) and got.dtype == T
    assert np.array_equal(got.view(U), O.AtWA(J, W, ncols).view(U))
    w = np.random.default_rng(4).uniform(0, 2, x.shape).astype(T)
    assert np.array_equal(li.AtWA(A, w).view(U), O.AtWA(J, W, ncols, w).view(U))
    dw = li.AtWA(A, li.to_device(w))
    assert li.is_device(dw) if hasattr(li, "is_device") else True
    assert np.array_equal(li.to_host(dw).view(U), O.AtWA(J, W, ncols, w).view(U))
    with pytest.raises(AssertionError):
        li.AtWA(A, w[:-1])
    for eps, mu in ((1e-3, 0.0), (0.0, 0.5), (1e-9, 0.25), (0, 0)):
        R = li.regularize(got, eps, mu)
        assert np.array_equal(R.view(U), O.regularize(got, eps, mu).view(U)), (eps, mu)
        Rd = li.regularize(li.to_device(got), eps, mu)
        assert np.array_equal(li.to_host(Rd).view(U), R.view(U))
    one = np.asfortranarray(np.array([[2.0]], dtype=T))
    assert li.regularize(one, 0.5, 0.25)[0, 0] == T(2.0 + 1.0 + 0.5)
    with pytest.raises(li.ErrorException):
        li.regularize(np.zeros((4, 4), dtype=T, order="F"), 1e-3, 0.0)
    with pytest.raises(AssertionError):
        li.regularize(got, -1.0, 0.0)


@pytest.mark.gpu
def test_fit_recovers_a_spline_model():
    import li_b200 as li
    T = np.float64
    rng = np.random.default_rng(8)
    ncols, nrows = 40, 900
    pos = rng.uniform(1, ncols, nrows)
    A = li.SparseInterpolator(li.CubicSpline(T), pos, ncols)
    truth = np.cumsum(rng.standard_normal(ncols)) * 0.3
    y = np.asarray(A * truth) + 1e-3 * rng.standard_normal(nrows)
    w = rng.uniform(0.5, 1.5, nrows)
    Ad = A.sparse().toarray()
    for ww in (None, w):
        got = li.fit(A, y, ww, epsilon=1e-9, mu=0.0)
        W = np.ones(nrows) if ww is None else ww
        M = Ad.T @ (W[:, None] * Ad)
        want = np.linalg.solve(M + 1e-9 * M.diagonal().max() * np.eye(ncols), Ad.T @ (W * y))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-8)
    assert np.max(np.abs(li.fit(A, y, w) - truth)) < 0.02

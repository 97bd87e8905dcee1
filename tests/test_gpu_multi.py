"""GPU tests of the iterative-fitting pieces: the multi-frame accumulate kernel (model += sum_k H_k' y_k in one
launch), the NCCL entry points of the C ABI on a one-rank communicator, the device vector operations and the
conjugate-gradient driver.  Everything goes through the C ABI; the oracle (CPU restatement of the reference loops)
is the checker."""
import numpy as np
import pytest

import li_oracle as O
from util import RTOL, assert_close, assert_within_bound, sep2d_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def li():
    import li_b200
    return li_b200


def _frames(li, T, ker_name, S, cond, m1, m2, n1, n2, nframes, rate, maxshift, seed):
    """Per-frame shifted resampling operators (BASELINE config #5 shape): product tables, oracle tables, data."""
    rng = np.random.default_rng(seed)
    B = li.SafeFlat if cond == O.SAFEFLAT else li.Flat
    ker = getattr(li, ker_name)(T, S, B) if ker_name == "LanczosKernel" else getattr(li, ker_name)(T, B)
    fam = {"LanczosKernel": O.LANCZOS, "CatmullRomSpline": O.CATMULLROM, "LinearSpline": O.LINEAR,
           "QuadraticSpline": O.QUADRATIC}[ker_name]
    ok = O.kernel(fam, S, cond)
    tabs, otabs, ys = [], [], []
    for k in range(nframes):
        d1, d2 = rng.uniform(-maxshift, maxshift, 2)
        x1 = (np.arange(1, n1 + 1) - 1) / rate + 1 + d1
        x2 = (np.arange(1, n2 + 1) - 1) / rate + 1 + d2
        A1 = li.TabulatedInterpolator(T, 1, ker, x1, m1)
        A2 = li.TabulatedInterpolator(T, 2, ker, x2, m2)
        tabs.append((A1, A2))
        otabs.append(O.maketable(ok, x1, m1, T) + O.maketable(ok, x2, m2, T))
        ys.append(np.asfortranarray(rng.standard_normal((n1, n2)).astype(T)))
    return tabs, otabs, ys


def _oracle_sum(otabs, ys, m1, m2, alpha, beta, model0, T):
    """Reference order: the frames one after the other into the model (beta applied once, before the first)."""
    acc = (beta * model0.astype(np.float64)).astype(T)
    bound = np.abs(beta * model0).astype(np.float64)
    for (J1, W1, J2, W2), y in zip(otabs, ys):
        acc = (acc + T(alpha) * sep2d_oracle(J1, W1, m1, J2, W2, m2, y, True)).astype(T)
        bound += abs(alpha) * sep2d_oracle(J1, np.abs(W1).astype(np.float64), m1, J2, np.abs(W2).astype(np.float64), m2,
                                           np.abs(y).astype(np.float64), True)
    return acc, bound


CASES = [
    # ker, S, cond, m1, m2, n1, n2, frames, rate, maxshift
    ("LanczosKernel", 6, O.SAFEFLAT, 300, 280, 300, 280, 9, 1.0, 8.0),     # config #5 in small: same-rate shifts +-8
    ("LanczosKernel", 4, O.FLAT, 150, 140, 300, 280, 5, 2.0, 1.5),         # 2x up-sampling frames
    ("CatmullRomSpline", 4, O.SAFEFLAT, 520, 96, 520, 96, 6, 1.0, 3.0),    # several strips, a single chunk
    ("LinearSpline", 2, O.FLAT, 260, 300, 260, 300, 4, 1.0, 0.7),
    ("QuadraticSpline", 3, O.FLAT, 128, 200, 128, 200, 4, 1.0, 2.2),
    ("LanczosKernel", 6, O.SAFEFLAT, 301, 283, 300, 280, 5, 1.0, 8.0),     # odd model width: the last column pair is half empty
    ("CatmullRomSpline", 4, O.FLAT, 200, 190, 300, 284, 4, 1.5, 2.0),      # irregular rate: generic row schedule everywhere
    ("LanczosKernel", 4, O.FLAT, 1100, 64, 1100, 64, 3, 1.0, 6.0),         # three strips, wide tap lists at the edges
]


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}{c[1]}-{c[8]}x" for c in CASES])
def test_multi_frame_accumulate_matches_oracle(li, case, T):
    ker, S, cond, m1, m2, n1, n2, nf, rate, sh = case
    tabs, otabs, ys = _frames(li, T, ker, S, cond, m1, m2, n1, n2, nf, rate, sh, seed=S * 100 + nf)
    rng = np.random.default_rng(1)
    model0 = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
    dys = [li.to_device(y) for y in ys]
    tl = [(a.table, b.table) for a, b in tabs]
    # (the first call builds the per-table adjoint metadata and the plan of the batch: several one-off launches)
    li.adjoint_accumulate(tl, dys, li.to_device(model0), alpha=1, beta=0)
    for alpha, beta in ((1, 0), (1, 1), (-0.5, 2.0)):
        ref, bound = _oracle_sum(otabs, ys, m1, m2, alpha, beta, model0, T)
        model = li.to_device(model0)
        l0 = li.launch_count()
        li.adjoint_accumulate(tl, dys, model, alpha=alpha, beta=beta)
        launches = li.launch_count() - l0
        got = li.to_host(model)
        assert_within_bound(got, ref, bound, T, f"alpha={alpha} beta={beta}")
        assert launches <= 3, f"{launches} launches: the multi-frame kernel did not take the batch"
        # run-to-run deterministic
        model2 = li.to_device(model0)
        li.adjoint_accumulate(tl, dys, model2, alpha=alpha, beta=beta)
        assert np.array_equal(li.to_host(model2), got)
    # the same sum through the batch entry point (beta = 1) and, as a cross-check, one launch per frame
    ref, bound = _oracle_sum(otabs, ys, m1, m2, 1, 1, model0, T)
    model = li.to_device(model0)
    li.sep_apply_batch(True, tl, 1, dys, 1, [model] * nf)
    assert_within_bound(li.to_host(model), ref, bound, T, "li_sep_apply_batch")
    li.set_option("no_multi", 1)
    try:
        model3 = li.to_device(model0)
        li.adjoint_accumulate(tl, dys, model3, alpha=1, beta=1)
        assert_within_bound(li.to_host(model3), ref, bound, T, "per-frame fallback")
    finally:
        li.set_option("no_multi", 0)


def test_multi_frame_nonfinite_sample_stays_in_its_footprint(li):
    T = np.float32
    tabs, otabs, ys = _frames(li, T, "LanczosKernel", 6, O.SAFEFLAT, 300, 280, 300, 280, 4, 1.0, 5.0, seed=7)
    ys[2][137, 91] = np.inf
    ref, _ = _oracle_sum(otabs, ys, 300, 280, 1, 0, np.zeros((300, 280), T), T)
    model = li.device_empty((300, 280), T)
    li.adjoint_accumulate([(a.table, b.table) for a, b in tabs], [li.to_device(y) for y in ys], model)
    got = li.to_host(model)
    assert np.array_equal(np.isfinite(got), np.isfinite(ref)), "non-finite footprint differs from the reference's"


def test_multi_frame_frame_groups_and_one_rank_comm(li):
    """Many frames on a small model: the tiles alone cannot fill the machine, so the launch splits the frames into
    groups (partial models + last-arriving CTA); with a one-rank communicator the model is produced in bands, each
    followed by an (identity) NCCL all-reduce on the communicator's stream -- the whole multi-GPU code path."""
    T = np.float32
    m = 256
    tabs, otabs, ys = _frames(li, T, "LanczosKernel", 6, O.SAFEFLAT, m, m, m, m, 40, 1.0, 8.0, seed=11)
    tl = [(a.table, b.table) for a, b in tabs]
    dys = [li.to_device(y) for y in ys]
    model0 = np.asfortranarray(np.random.default_rng(2).standard_normal((m, m)).astype(T))
    ref, bound = _oracle_sum(otabs, ys, m, m, 1, 0.5, model0, T)
    model = li.to_device(model0)
    li.adjoint_accumulate(tl, dys, model, alpha=1, beta=0.5)
    assert_within_bound(li.to_host(model), ref, bound, T, "frame groups")
    comm = li.Comm.init_rank(li.Comm.unique_id(), 1, 0)
    assert (comm.rank, comm.nranks) == (0, 1)
    for nbands in (1, 3, 4):
        model = li.to_device(model0)
        li.adjoint_accumulate(tl, dys, model, alpha=1, beta=0.5, comm=comm, nbands=nbands)
        assert_within_bound(li.to_host(model), ref, bound, T, f"one-rank communicator, {nbands} bands")
    x = li.to_device(np.arange(10, dtype=T))
    comm.allreduce_sum(x)
    assert np.array_equal(li.to_host(x), np.arange(10, dtype=T))
    # a rank without frames still takes part (and contributes nothing)
    model = li.to_device(model0)
    li.adjoint_accumulate([], [], model, alpha=1, beta=2.0, comm=comm, nbands=2)
    np.testing.assert_allclose(li.to_host(model), 2 * model0, rtol=1e-6)
    comm.destroy()


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_vector_operations(li, T):
    rng = np.random.default_rng(3)
    n = 1_000_003
    x, y = rng.standard_normal(n).astype(T), rng.standard_normal(n).astype(T)
    dx, dy = li.to_device(x), li.to_device(y)
    ref = float(np.dot(x.astype(np.float64), y.astype(np.float64)))
    got = li.vdot(dx, dy)
    assert abs(got - ref) <= 1e-12 * np.sum(np.abs(x.astype(np.float64) * y))
    assert li.vdot(dx, dy) == got                                     # fixed reduction order
    li.vupdate(dy, -0.75, dx)
    np.testing.assert_allclose(li.to_host(dy), y + T(-0.75) * x, rtol=2e-6 if T == np.float32 else 1e-14, atol=1e-6 if T == np.float32 else 1e-14)
    dz = li.device_empty((n,), T)
    li.vcombine(dz, 2.0, dx, 0.0, dz)                                 # beta = 0: dz (uninitialised) is never read
    np.testing.assert_array_equal(li.to_host(dz), T(2.0) * x)
    li.vcombine(dz, 1.0, dx, -1.0, dz)
    np.testing.assert_allclose(li.to_host(dz), -x, rtol=1e-6)


def _numpy_cg(apply_A, b, maxiter):
    """Plain linear CG in Float64 (same recurrence as LazyAlgebra's conjgrad), fixed number of iterations."""
    x = np.zeros_like(b)
    r = b.copy()
    rho = float(np.vdot(r, r))
    p = r.copy()
    for k in range(maxiter):
        if k:
            p = r + (rho / rho_old) * p
        q = apply_A(p)
        alpha = rho / float(np.vdot(p, q))
        x += alpha * p
        r -= alpha * q
        rho_old, rho = rho, float(np.vdot(r, r))
    return x


def test_conjgrad_benchmarks_jl_shape(li):
    """test/benchmarks.jl:23-40: H = A1*A2 (CatmullRom, 300x200 <- 43x29, Float64), x = conjgrad(H'*H, H'*y)."""
    T = np.float64
    n1, n2, m1, m2 = 300, 200, 43, 29
    rng = np.random.default_rng(4)
    y = np.asfortranarray(rng.standard_normal((n1, n2)))
    ker = li.CatmullRomSpline(T)
    x1 = np.linspace(1, m1, n1)
    x2 = np.linspace(1, m2, n2)
    A1 = li.TabulatedInterpolator(T, 1, ker, x1, m1)
    A2 = li.TabulatedInterpolator(T, 2, ker, x2, m2)
    H = A1 * A2
    ok = O.kernel(O.CATMULLROM, 4)
    J1, W1 = O.maketable(ok, x1, m1, T)
    J2, W2 = O.maketable(ok, x2, m2, T)

    def gram(v):
        return sep2d_oracle(J1, W1, m1, J2, W2, m2, sep2d_oracle(J1, W1, m1, J2, W2, m2, v, False), True)

    b = sep2d_oracle(J1, W1, m1, J2, W2, m2, y, True)
    iters = 12
    ref = _numpy_cg(gram, b, iters)
    dy = li.to_device(y)
    db = H.adj * dy
    assert_close(li.to_host(db), b, T, factor=10)
    l0 = li.launch_count()
    x, res = li.conjgrad(H.adj * H, db, ftol=0, maxiter=iters, check_every=iters, info=True)
    assert res.iters == iters
    np.testing.assert_allclose(li.to_host(x), ref, rtol=1e-7, atol=1e-9 * np.max(np.abs(ref)))
    # H'*H runs as two fused applies per iteration (plus the vector kernels), nothing per factor
    assert li.launch_count() - l0 <= iters * 8 + 8
    # the same solve without any callback into Python
    x2_, res2 = li.fit_separable([A1.table, A2.table], dy, ftol=0, maxiter=iters, check_every=iters, info=True)
    np.testing.assert_allclose(li.to_host(x2_), ref, rtol=1e-7, atol=1e-9 * np.max(np.abs(ref)))
    # convergence tests: the reference's call (ftol = 1e-8, maxiter = 50) stops early and fits the data
    x3, res3 = li.conjgrad(H.adj * H, db, ftol=1e-8, maxiter=50, check_every=1, info=True)
    assert res3.status == 1 and res3.iters < 50, res3
    resid = li.to_host(H * x3) - y
    resid_ref = sep2d_oracle(J1, W1, m1, J2, W2, m2, _numpy_cg(gram, b, 50), False) - y
    assert np.linalg.norm(resid) <= 1.0001 * np.linalg.norm(resid_ref)
    # regularised fit through li_conjgrad_sep: (H'H + mu I) x = H'y
    mu = 0.3
    refm = _numpy_cg(lambda v: gram(v) + mu * v, b, 15)
    xm = li.fit_separable([A1.table, A2.table], dy, mu=mu, ftol=0, maxiter=15, check_every=15)
    np.testing.assert_allclose(li.to_host(xm), refm, rtol=1e-7, atol=1e-9 * np.max(np.abs(refm)))


def test_conjgrad_float32_and_errors(li):
    T = np.float32
    n, m = 256, 64
    ker = li.LanczosKernel(T, 4)
    xs = np.linspace(1, m, n)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, xs, m), li.TabulatedInterpolator(T, 2, ker, xs, m)
    H = A1 * A2
    truth = np.asfortranarray(np.random.default_rng(5).standard_normal((m, m)).astype(T))
    dy = H * li.to_device(truth)
    x, res = li.conjgrad(H.adj * H, H.adj * dy, ftol=1e-10, maxiter=200, check_every=10, info=True)
    err = np.max(np.abs(li.to_host(x) - truth)) / np.max(np.abs(truth))
    assert err < 2e-3, (err, res)
    with pytest.raises(li.ErrorException):          # host arrays: the solver works on device arrays
        li.conjgrad(H.adj * H, np.zeros((m, m), T))
    neg = li.TabulatedInterpolator.from_tables(np.arange(1, m + 1)[None, :], -np.ones((1, m), T), m, d=1)
    with pytest.raises(li.ErrorException):          # "operator is not positive definite"
        li.conjgrad(neg, li.to_device(np.ones((m,), T)), maxiter=5, check_every=1)


def test_composition_with_complex_data(li):
    """(A1*A2)*x with Complex{T} data (src/tabulated.jl:196: RorC): ADVICE r1 -- the fused path must not swallow the
    imaginary part; the per-factor path handles it."""
    T = np.float32
    ker = li.CatmullRomSpline(T)
    m1, m2, n1, n2 = 40, 30, 90, 70
    x1, x2 = np.linspace(1, m1, n1), np.linspace(1, m2, n2)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x1, m1), li.TabulatedInterpolator(T, 2, ker, x2, m2)
    H = A1 * A2
    ok = O.kernel(O.CATMULLROM, 4)
    J1, W1 = O.maketable(ok, x1, m1, T)
    J2, W2 = O.maketable(ok, x2, m2, T)
    rng = np.random.default_rng(6)
    z = np.asfortranarray((rng.standard_normal((m1, m2)) + 1j * rng.standard_normal((m1, m2))).astype(np.complex64))
    for kind in ("host", "device"):
        zin = z if kind == "host" else li.to_device(z)
        out = li.to_host(H * zin)
        assert out.dtype == np.complex64 and out.shape == (n1, n2)
        for part, ref_in in ((out.real, z.real), (out.imag, z.imag)):
            ref = sep2d_oracle(J1, W1, m1, J2, W2, m2, np.asfortranarray(ref_in.astype(T)), False)
            assert_close(part, ref, T, factor=4)
        back = li.to_host(H.adj * (out if kind == "host" else li.to_device(out)))
        ref = sep2d_oracle(J1, W1, m1, J2, W2, m2, np.asfortranarray(out.real.astype(T)), True)
        assert_close(back.real, ref, T, factor=4)


def test_options_roundtrip(li):
    assert li.get_option("no_fused") == 0
    li.set_option("no_fused", 1)
    assert li.get_option("no_fused") == 1
    li.set_option("no_fused", 0)
    with pytest.raises(li.ErrorException):
        li.set_option("no_such_option", 1)

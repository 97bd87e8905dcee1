"""Derivative kernels (`*Prime`, re-exported at src/LinearInterpolators.jl:23-41; InterpolationKernels 0.1.x): the
weights are d/dt of the weights of the base kernel.  CPU part: the oracle's formulas against central differences of
its own base weights, and the host mirror's types.  GPU part: tables bit-exact against the oracle, and the derivative
of an interpolated function."""
import numpy as np
import pytest

import li_oracle as O
from util import CONDITIONS, DTYPES, KERNEL_IDS, KERNELS, assert_close, hostile_positions, oracle_kernel, product_kernel


def _prime(entry, cond):
    _, fam, S, params, _ = entry
    return O.kernel(fam, S, cond, params, prime=True)


@pytest.mark.parametrize("entry", KERNELS, ids=KERNEL_IDS)
def test_oracle_prime_weights_are_the_derivative(entry):
    k, kp = oracle_kernel(entry, O.FLAT), _prime(entry, O.FLAT)
    h = 1e-6
    for t in np.linspace(0.003, 0.997, 57):
        fd = (np.asarray(O.getweights(k, t + h, np.float64)) - np.asarray(O.getweights(k, t - h, np.float64))) / (2 * h)
        wp = np.asarray(O.getweights(kp, t, np.float64))
        assert np.max(np.abs(fd - wp)) < 5e-9
        w32 = np.asarray(O.getweights(kp, np.float32(t), np.float32))
        # (the Lanczos derivative is a difference of nearly equal terms close to the nodes: the Float32 rounding
        # errors of that formula grow like eps / distance to the node)
        tol32 = 2e-6 if entry[1] != O.LANCZOS else 5e-7 / min(t, 1 - t)
        assert np.max(np.abs(w32 - np.asarray(O.getweights(kp, float(np.float32(t)), np.float64)))) < tol32
    if entry[1] != O.LANCZOS:      # normalised kernels: the weights sum to one, their derivatives to zero
        assert abs(np.sum(O.getweights(kp, 0.37, np.float64))) < 1e-14


def test_host_mirror_prime_types():
    import li_b200 as li
    k = li.CatmullRomSpline(np.float32, li.SafeFlat)
    kp = k.prime
    assert isinstance(kp, li.CatmullRomSplinePrime) and kp.S == 4 and kp.T == np.float32 and kp.B is li.SafeFlat
    assert kp.desc().family == k.desc().family | 16 and k.desc().family < 16
    assert li.LanczosKernelPrime(np.float64, 6).S == 6 and li.KeysSplinePrime(-0.5).params == (-0.5,)
    assert li.MitchellNetravaliSplinePrime().params == (1 / 3, 1 / 3)
    with pytest.raises(li.ArgumentError):
        kp.prime
    assert li.iscardinal(k) and li.isnormalized(k) and not li.iscardinal(li.CubicSpline())
    assert not li.isnormalized(li.LanczosKernel(4)) and not li.iscardinal(kp)
    assert li.Flat(kp).B is li.Flat and isinstance(li.Flat(kp), li.CatmullRomSplinePrime)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("cname,cond", CONDITIONS)
@pytest.mark.parametrize("T", DTYPES)
def test_prime_tables_bit_exact(entry, cname, cond, T):
    import li_b200 as li
    rng = np.random.default_rng(43)
    for ncols in (1, 5, 300):
        x = hostile_positions(ncols, T, rng)
        J, W = O.maketable(_prime(entry, cond), x, ncols, T)
        A = li.TabulatedInterpolator(T, product_kernel(li, entry, cond, T).prime, x, ncols)
        assert np.array_equal(A.J, J), "index table differs"
        U = np.uint32 if T == np.float32 else np.uint64
        nan = np.isnan(W)
        assert np.array_equal(np.isnan(A.W), nan), "NaN pattern differs"
        assert np.array_equal(A.W.view(U)[~nan], W.view(U)[~nan]), "weight table differs in bits"


@pytest.mark.gpu
def test_prime_operator_differentiates():
    """ker' interpolates the derivative: for samples of a smooth function on the integer grid, A' * f approximates f'
    (cubic B-spline pre-filtering aside, Catmull-Rom and Lanczos are interpolating: third / first order accurate)."""
    import li_b200 as li
    T = np.float64
    m = 400
    g = np.arange(1, m + 1, dtype=T)
    f = np.sin(0.05 * g) + 0.3 * np.cos(0.021 * g)
    x = np.linspace(20.25, m - 20.5, 777)
    want = 0.05 * np.cos(0.05 * x) - 0.3 * 0.021 * np.sin(0.021 * x)
    for ker, tol in ((li.CatmullRomSplinePrime(T), 5e-5), (li.KeysSplinePrime(T, -0.5), 5e-5), (li.LanczosKernelPrime(T, 8), 2e-2)):   # (the Lanczos weights are not normalised: their derivatives sum to ~5e-3, times |f|)
        A = li.TabulatedInterpolator(ker, x, m)
        got = np.asarray(A * f)
        assert np.max(np.abs(got - want)) < tol, (type(ker).__name__, np.max(np.abs(got - want)))
    # the same through the on-the-fly operator and against the oracle's table loops
    ker = li.CatmullRomSpline(T).prime
    J, W = O.maketable(O.kernel(O.CATMULLROM, 4, O.FLAT, (), prime=True), x, m, T)
    assert_close(li.apply(li.Direct, ker, x, f), O.tab_direct(J, W, m, 1, f), T, factor=4)

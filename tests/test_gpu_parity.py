"""GPU parity tests: the CUDA path (through the Python host mirror -> C ABI -> libli_b200.so) against
the CPU oracle on identical inputs.  Tables must match bit for bit; interpolated values within
1e-6 (Float32) / 1e-12 (Float64) relative, as stated in BASELINE.json."""
import numpy as np
import pytest

from util import RO as O   # the oracle, with per-element error bounds attached to its results
from util import (CONDITIONS, DTYPES, KERNEL_IDS, KERNELS, assert_close, hostile_positions, oracle_kernel,
                  product_kernel)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def li():
    import li_b200
    return li_b200


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("entry", KERNELS, ids=KERNEL_IDS)
@pytest.mark.parametrize("cname,cond", CONDITIONS)
@pytest.mark.parametrize("T", DTYPES)
def test_tables_bit_exact(li, entry, cname, cond, T):
    rng = np.random.default_rng(42)
    for ncols in (1, 2, 5, 11, 300):
        x = hostile_positions(ncols, T, rng)
        ok = O.kernel(entry[1], entry[2], cond, entry[3])
        J, W = O.maketable(ok, x, ncols, T)
        A = li.TabulatedInterpolator(T, product_kernel(li, entry, cond, T), x, ncols)
        assert A.J.dtype == np.int64 and A.W.dtype == T
        assert np.array_equal(A.J, J), "index table differs"
        # bit-for-bit, except that NaN payloads are not compared (a subnormal offset makes the Lanczos
        # formula evaluate 0/0 in the reference as well)
        U = np.uint32 if T == np.float32 else np.uint64
        nan = np.isnan(W)
        assert np.array_equal(np.isnan(A.W), nan), "NaN pattern differs"
        assert np.array_equal(A.W.view(U)[~nan], W.view(U)[~nan]), "weight table differs in bits"


@pytest.mark.parametrize("T", DTYPES)
def test_inexact_error(li, T):
    lin = li.LinearSpline(T)
    for bad in (np.nan, 1e19, -1e19, np.inf):
        with pytest.raises(li.InexactError):
            li.TabulatedInterpolator(lin, np.array([1.0, bad, 2.0], dtype=T), 5)
    A = li.TabulatedInterpolator(li.SafeFlat(lin), np.array([1e30, -np.inf], dtype=T), 5)
    assert A.J.tolist() == [[5, 1], [5, 1]]
    with pytest.raises(li.InexactError):
        li.TabulatedInterpolator(li.SafeFlat(lin), np.array([np.nan], dtype=T), 5)


def test_getcoefs(li):
    # test/runtests.jl:90-98 positions + SURVEY Appendix A examples
    cr = li.CatmullRomSpline(np.float64)
    assert li.getcoefs(cr, 5, 0.5) == (1, 1, 1, 2, -1 / 16, 9 / 16, 9 / 16, -1 / 16)
    crs = li.SafeFlat(cr)
    assert li.getcoefs(crs, 5, -1e-300) == (1, 1, 1, 1, 0, 0, 0, 1)
    assert li.getcoefs(crs, 5, 6.0000001) == (5, 5, 5, 5, 1, 0, 0, 0)
    for x in (0, 0.1, -1.2):
        ref = O.getcoefs(O.kernel(O.CATMULLROM, 4), 4, x, np.float64)
        got = li.getcoefs(cr, 4, x)
        assert got[:4] == tuple(ref[0]) and got[4:] == tuple(ref[1])


# ------------------------------------------------------------------------------------------------
def _data():
    sub = 3
    x = np.linspace(-1.0, 1.0, 100 * sub + 1)
    xsub = x[::sub]
    y = np.cos(2 * x * (x + 2))
    return x, xsub, y, y[::sub], np.linspace(1.0, float(len(xsub)), len(x))


def distance(a, b):
    return float(np.mean(np.abs(np.asarray(a) - np.asarray(b))))


def test_reference_suite_unidimensional(li):
    # test/runtests.jl:15-49 verbatim, through the product API
    ker = li.LinearSpline(np.float32)
    src = np.array([0, 0, 1, 0, 0])
    x = np.linspace(-1, 7, 81)
    out = li.apply(ker, x, src)
    assert out.dtype == np.float32
    assert out[x == 2.5][0] == 0.5
    out2 = np.empty_like(out)
    li.apply_b(out2, ker, x, src)
    assert np.array_equal(out2, out)
    li.apply_b(2, li.Direct, ker, x, src, 0, out2)
    assert np.array_equal(out2, 2 * out)
    out2[:] = 1
    li.apply_b(1, li.Direct, ker, x, src, 2, out2)
    assert np.array_equal(out2, out + 2)
    out2[:] = 3
    li.apply_b(0, li.Direct, ker, x, src, 2, out2)
    assert np.all(out2 == 6)
    adj = li.apply(li.Adjoint, ker, x, out, len(src))
    adj2 = np.empty_like(adj)
    li.apply_b(2, li.Adjoint, ker, x, out, 0, adj2)
    np.testing.assert_allclose(adj2, 2 * adj, rtol=1e-6)
    adj2[:] = 1
    li.apply_b(1, li.Adjoint, ker, x, out, 2, adj2)
    np.testing.assert_allclose(adj2, adj + 2, rtol=1e-6)
    adj2[:] = 3
    li.apply_b(0, li.Adjoint, ker, x, out, 2, adj2)
    assert np.all(adj2 == 6)


REF_KERNELS = [k for k in KERNELS if k[0] in ("Rectangular", "Linear", "Quadratic", "Cubic", "CatmullRom",
                                               "Keys(-0.4)", "MitchellNetravali", "Lanczos4", "Lanczos6")]


@pytest.mark.parametrize("entry", REF_KERNELS, ids=[k[0] for k in REF_KERNELS])
@pytest.mark.parametrize("cname,cond", CONDITIONS)
def test_reference_suite_tabulated_and_sparse(li, entry, cname, cond):
    # test/runtests.jl:100-137
    x, xsub, y, ysub, t = _data()
    tol = 0.02 if entry[1] == O.RECTANGULAR else 0.006
    ker = product_kernel(li, entry, cond, np.float64)
    Tb = li.TabulatedInterpolator(ker, t, len(xsub))
    assert distance(Tb(ysub), Tb * ysub) <= 0
    assert distance(Tb * ysub, y) <= tol
    assert distance(Tb * ysub, li.apply(ker, t, ysub)) <= 1e-15
    Sp = li.SparseInterpolator(ker, t, len(xsub))
    assert distance(Sp(ysub), Sp * ysub) <= 1e-15
    assert distance(Sp(ysub), Tb(ysub)) <= 1e-15
    assert distance(Sp(ysub), y) <= tol
    assert distance(Sp(ysub), li.apply(ker, t, ysub)) <= 1e-15
    # and against the oracle on the same inputs
    J, W = O.maketable(oracle_kernel(entry, cond), t, len(xsub), np.float64)
    assert_close(Tb * ysub, O.tab_direct(J, W, len(xsub), 1, ysub), np.float64)


def test_reference_suite_deprecation(li):
    x, xsub, y, ysub, t = _data()
    with pytest.warns(DeprecationWarning):
        li.SparseInterpolator(np.float32, li.RectangularSpline(), t, len(xsub))   # test/runtests.jl:134-136


def test_reference_suite_affine_adjoint_identity(li):
    # test/runtests.jl:139-158
    rng = np.random.default_rng(5)
    rows, cols = (11, 16), (10, 15)
    c = (5.6, 6.4)
    R = c + li.rotate(li.AffineTransform2D() - c, 0.2)
    kerlist = (li.RectangularSpline(), li.LinearSpline(), li.CatmullRomSpline())
    for ker1 in kerlist:
        for ker2 in kerlist:
            A = li.TwoDimensionalTransformInterpolator(rows, cols, ker1, ker2, R)
            x = np.asfortranarray(rng.standard_normal(cols))
            y = np.asfortranarray(rng.standard_normal(rows))
            np.testing.assert_allclose(li.vdot(A * x, y), li.vdot(x, A.adj * y), rtol=1e-12)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", DTYPES)
@pytest.mark.parametrize("device", [False, True], ids=["host", "device"])
def test_tab_apply_nd(li, T, device):
    rng = np.random.default_rng(11)
    shape = (37, 19, 6)
    for entry in (KERNELS[1], KERNELS[4], KERNELS[9]):
        for d in (1, 2, 3):
            ncols = shape[d - 1]
            xs = np.sort(rng.uniform(-2, ncols + 3, 53)) if d != 2 else rng.uniform(-2, ncols + 3, 53)
            ok = oracle_kernel(entry, O.FLAT)
            J, W = O.maketable(ok, xs.astype(T), ncols, T)
            A = li.TabulatedInterpolator(T, d, product_kernel(li, entry, O.FLAT, T), xs, ncols)
            src = np.asfortranarray(rng.standard_normal(shape).astype(T))
            for alpha, beta in ((1, 0), (2.5, 0), (1, 1), (-0.5, 2), (0, 3), (0, 0)):
                ref0 = np.asfortranarray(rng.standard_normal(
                    tuple(53 if i == d - 1 else n for i, n in enumerate(shape))).astype(T))
                ref = O.tab_direct(J, W, ncols, d, src, alpha, beta, ref0.copy(order="F"))
                if device:
                    dy = li.to_device(ref0)
                    li.apply_b(alpha, li.Direct, A, li.to_device(src), False, beta, dy)
                    got = li.to_host(dy)
                else:
                    got = ref0.copy(order="F")
                    li.apply_b(alpha, li.Direct, A, src, False, beta, got)
                assert_close(got, ref, T, factor=4)
                # adjoint, deterministic gather and atomic variants
                ysrc = np.asfortranarray(rng.standard_normal(ref.shape).astype(T))
                dst0 = np.asfortranarray(rng.standard_normal(shape).astype(T))
                refa = O.tab_adjoint(J, W, ncols, d, ysrc, alpha, beta, dst0.copy(order="F"))
                for variant in (li.VARIANT_GATHER, li.VARIANT_ATOMIC):
                    A.variant = variant
                    if device:
                        dd = li.to_device(dst0)
                        li.apply_b(alpha, li.Adjoint, A, li.to_device(ysrc), False, beta, dd)
                        gota = li.to_host(dd)
                    else:
                        gota = dst0.copy(order="F")
                        li.apply_b(alpha, li.Adjoint, A, ysrc, False, beta, gota)
                    assert_close(gota, refa, T, factor=8)
                A.variant = li.VARIANT_AUTO


@pytest.mark.parametrize("T", DTYPES)
@pytest.mark.parametrize("style", ["aligned", "overhang", "down", "up2"])
def test_tab_apply_strided_streaming(li, T, style):
    """Strided dimensions of sorted (windowed, monotone) tables with 16-byte aligned planes take the streaming
    kernels (register window, vector loads): every kernel size, clamped edges, alpha / beta, non-finite samples."""
    rng = np.random.default_rng(17)
    shape = (36, 20, 8)
    for entry in (KERNELS[0], KERNELS[1], KERNELS[2], KERNELS[4], KERNELS[9], KERNELS[10]):
        for cname, cond in CONDITIONS:
            for d in (2, 3):
                ncols = shape[d - 1]
                nrows = {"aligned": 2 * ncols, "overhang": 33, "down": max(3, ncols // 3), "up2": 2 * ncols + 1}[style]
                xs = {"aligned": np.linspace(1, ncols, nrows), "overhang": np.linspace(-7.5, ncols + 9.25, nrows),
                      "down": np.linspace(1, ncols, nrows), "up2": 0.3 + 0.5 * np.arange(nrows)}[style]
                J, W = O.maketable(oracle_kernel(entry, cond), xs.astype(T), ncols, T)
                A = li.TabulatedInterpolator(T, d, product_kernel(li, entry, cond, T), xs, ncols)
                src = np.asfortranarray(rng.standard_normal(shape).astype(T))
                oshape = tuple(nrows if i == d - 1 else n for i, n in enumerate(shape))
                for alpha, beta in ((1, 0), (-1.5, 0.5)):
                    d0 = np.asfortranarray(rng.standard_normal(oshape).astype(T))
                    ref = O.tab_direct(J, W, ncols, d, src, alpha, beta, d0.copy(order="F"))
                    dy = li.to_device(d0)
                    li.apply_b(alpha, li.Direct, A, li.to_device(src), False, beta, dy)
                    assert_close(li.to_host(dy), ref, T, factor=4)
                    ysrc = np.asfortranarray(rng.standard_normal(oshape).astype(T))
                    c0 = np.asfortranarray(rng.standard_normal(shape).astype(T))
                    refa = O.tab_adjoint(J, W, ncols, d, ysrc, alpha, beta, c0.copy(order="F"))
                    dd = li.to_device(c0)
                    li.apply_b(alpha, li.Adjoint, A, li.to_device(ysrc), False, beta, dd)
                    assert_close(li.to_host(dd), refa, T, scale=float(np.max(np.abs(refa))) + 1, factor=16)
                # a non-finite sample reaches exactly the reference's footprint (no 0 * Inf from folded windows)
                ybad = np.asfortranarray(rng.standard_normal(oshape).astype(T))
                idx = [1, 2, 3]
                for pos in (0, nrows // 2, nrows - 1):
                    idx[d - 1] = pos
                    ybad[tuple(idx)] = np.inf
                refb = O.tab_adjoint(J, W, ncols, d, ybad, skip_zero=False)
                gotb = li.to_host(A.adj * li.to_device(ybad))
                if not np.isnan(W).any():
                    assert np.array_equal(np.isfinite(gotb), np.isfinite(refb))


def test_two_factors_on_3d_array_use_fused_kernels(li):
    # H = A1*A2 applied to a 3-D array: the third dimension is one more batch dimension of the fused 2-D kernels
    T = np.float32
    rng = np.random.default_rng(3)
    m, n, nz = (40, 30), (80, 60), 5
    ker = li.LanczosKernel(T, 4)
    x1, x2 = np.linspace(1, m[0], n[0]), np.linspace(1, m[1], n[1])
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x1, m[0]), li.TabulatedInterpolator(T, 2, ker, x2, m[1])
    J1, W1 = O.maketable(O.kernel(O.LANCZOS, 4), x1.astype(T), m[0], T)
    J2, W2 = O.maketable(O.kernel(O.LANCZOS, 4), x2.astype(T), m[1], T)
    src = np.asfortranarray(rng.standard_normal(m + (nz,)).astype(T))
    ref = O.tab_direct(J1, W1, m[0], 1, O.tab_direct(J2, W2, m[1], 2, src))
    dsrc = li.to_device(src)
    (A1 * A2) * dsrc                            # (first use: one-off table analysis kernels)
    n0 = li.launch_count()
    out = (A1 * A2) * dsrc
    assert li.launch_count() - n0 == 1          # one fused launch, not two passes
    assert_close(li.to_host(out), ref, T, factor=8)
    refa = O.tab_adjoint(J2, W2, m[1], 2, O.tab_adjoint(J1, W1, m[0], 1, ref))
    assert_close(li.to_host((A1 * A2).adj * out), refa, T, factor=64)


@pytest.mark.parametrize("T", DTYPES)
@pytest.mark.parametrize("npost", [64, 5000, 4099])   # one identity table; a divisor of npost as second dimension; prime
def test_tab_apply_contiguous_dim_through_fused_kernels(li, T, npost):
    """1-D applies along the first dimension of a sorted table over many columns run through the fused 2-D kernels
    with the identity as second factor."""
    rng = np.random.default_rng(29)
    ncols = 96
    for entry, cond in ((KERNELS[1], O.FLAT), (KERNELS[2], O.SAFEFLAT), (KERNELS[8], O.FLAT), (KERNELS[9], O.SAFEFLAT)):
        for nrows, xs in ((192, np.linspace(1, ncols, 192)), (100, np.linspace(-5.5, ncols + 8.25, 100))):
            J, W = O.maketable(oracle_kernel(entry, cond), xs.astype(T), ncols, T)
            A = li.TabulatedInterpolator(T, 1, product_kernel(li, entry, cond, T), xs, ncols)
            src = np.asfortranarray(rng.standard_normal((ncols, npost)).astype(T))
            y = np.asfortranarray(rng.standard_normal((nrows, npost)).astype(T))
            for alpha, beta in ((1, 0), (0.5, -2)):
                d0 = np.asfortranarray(rng.standard_normal((nrows, npost)).astype(T))
                ref = O.tab_direct(J, W, ncols, 1, src, alpha, beta, d0.copy(order="F"))
                dd = li.to_device(d0)
                li.apply_b(alpha, li.Direct, A, li.to_device(src), False, beta, dd)
                assert_close(li.to_host(dd), ref, T, factor=4)
                c0 = np.asfortranarray(rng.standard_normal((ncols, npost)).astype(T))
                refa = O.tab_adjoint(J, W, ncols, 1, y, alpha, beta, c0.copy(order="F"))
                dc = li.to_device(c0)
                li.apply_b(alpha, li.Adjoint, A, li.to_device(y), False, beta, dc)
                assert_close(li.to_host(dc), refa, T, scale=float(np.max(np.abs(refa))) + 1, factor=16)


def test_beta_zero_never_reads_dst(li):
    T = np.float32
    A = li.TabulatedInterpolator(T, li.CatmullRomSpline(T), np.linspace(1, 9, 33), 9)
    src = np.arange(9, dtype=T)
    y = np.full(33, np.nan, dtype=T)
    li.apply_b(1, li.Direct, A, src, False, 0, y)
    assert np.all(np.isfinite(y))
    z = np.full(9, np.nan, dtype=T)
    li.apply_b(1, li.Adjoint, A, y, False, 0, z)
    assert np.all(np.isfinite(z))


def test_gather_adjoint_is_deterministic(li):
    T = np.float32
    rng = np.random.default_rng(2)
    xs = rng.uniform(-5, 70, 5000)   # unsorted, many clamped rows -> heavy edge columns
    A = li.TabulatedInterpolator(T, li.LanczosKernel(T, 6), xs, 64)
    y = li.to_device(rng.standard_normal(5000).astype(T))
    outs = [li.to_host(A.adj * y) for _ in range(3)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    J, W = O.maketable(O.kernel(O.LANCZOS, 6), xs.astype(T), 64, T)
    assert_close(outs[0], O.tab_adjoint(J, W, 64, 1, li.to_host(y)), T, factor=50)
    # the heavy columns (here the first and the last one: thousands of entries each) are summed by a CTA each, in
    # segments combined by a fixed tree; several slices, both element types, alpha / beta
    for TT in DTYPES:
        A2 = li.TabulatedInterpolator(TT, 1, li.LanczosKernel(TT, 6), xs, 64)
        J2, W2 = O.maketable(O.kernel(O.LANCZOS, 6), xs.astype(TT), 64, TT)
        assert max(np.bincount(J2.ravel())) > 2000
        y2 = np.asfortranarray(rng.standard_normal((5000, 3)).astype(TT))
        d0 = np.asfortranarray(rng.standard_normal((64, 3)).astype(TT))
        got = []
        for _ in range(2):
            dd = li.to_device(d0.copy(order="F"))
            li.apply_b(0.5, li.Adjoint, A2, li.to_device(y2), False, 2, dd)
            got.append(li.to_host(dd))
        assert np.array_equal(got[0], got[1])
        assert_close(got[0], O.tab_adjoint(J2, W2, 64, 1, y2, 0.5, 2, d0.copy(order="F")), TT)


def test_sparse_interpolator_mixed_element_types(li):
    """src/sparse.jl:201-275: x may be any Real, y any AbstractFloat; the sums run in float(promote_type(Ta, Tx))."""
    rng = np.random.default_rng(12)
    n, m = 300, 40
    pos = rng.uniform(0.5, m + 0.5, n)
    A = li.SparseInterpolator(li.CatmullRomSpline(np.float32), pos, m)
    J, W = O.maketable(O.kernel(O.CATMULLROM, 4), pos.astype(np.float32), m, np.float32)
    W64 = np.asfortranarray(W.astype(np.float64))
    x64 = rng.standard_normal(m)                                 # Float64 data through a Float32 operator: Float64 sums
    y64 = np.empty(n, dtype=np.float64)
    li.apply_b(1, li.Direct, A, x64, False, 0, y64)
    assert_close(y64, O.tab_direct(J, W64, m, 1, x64), np.float64)
    y32 = rng.standard_normal(n).astype(np.float32)              # ... stored into a Float32 destination, beta != 0
    want = (0.5 * np.asarray(O.tab_direct(J, W64, m, 1, x64)) + np.float32(2) * y32).astype(np.float32)
    li.apply_b(0.5, li.Direct, A, x64, False, 2, y32)
    np.testing.assert_allclose(y32, want, rtol=0, atol=2e-7 * np.max(np.abs(want)))
    xi = rng.integers(-5, 6, m)                                  # integer data: promote_type(Float32, Int64) = Float32
    yi = np.empty(n, dtype=np.float32)
    li.apply_b(1, li.Direct, A, xi, False, 0, yi)
    assert_close(yi, O.tab_direct(J, W, m, 1, xi.astype(np.float32)), np.float32)
    z64 = np.empty(m, dtype=np.float64)                          # adjoint, Float64 source
    li.apply_b(1, li.Adjoint, A, rng.standard_normal(n), False, 0, z64)
    assert z64.dtype == np.float64 and np.isfinite(z64).all()
    src = rng.standard_normal(n)
    li.apply_b(1, li.Adjoint, A, src, False, 0, z64)
    assert_close(z64, O.tab_adjoint(J, W64, m, 1, src), np.float64)
    z0 = np.full(m, np.nan, dtype=np.float64)                    # alpha = 0, beta = 0: zero fill without reading
    li.apply_b(0, li.Adjoint, A, src, False, 0, z0)
    assert np.array_equal(z0, np.zeros(m))


def test_imported_arbitrary_table(li):
    # user-defined tables (not window-shaped): generic gather path
    rng = np.random.default_rng(9)
    T = np.float64
    S, nrows, ncols = 3, 40, 17
    J = rng.integers(1, ncols + 1, size=(S, nrows))
    W = rng.standard_normal((S, nrows))
    A = li.TabulatedInterpolator.from_tables(J, W, ncols)
    assert not A.table.windowed
    x = rng.standard_normal(ncols)
    assert_close(A * x, O.tab_direct(np.asfortranarray(J), np.asfortranarray(W), ncols, 1, x), T)
    y = rng.standard_normal(nrows)
    assert_close(A.adj * y, O.tab_adjoint(np.asfortranarray(J), np.asfortranarray(W), ncols, 1, y), T, factor=4)
    with pytest.raises(li.ErrorException):
        li.TabulatedInterpolator.from_tables(J + ncols, W, ncols)          # out of bound index


def test_complex_data(li):
    rng = np.random.default_rng(4)
    T = np.float64
    A = li.TabulatedInterpolator(T, 2, li.CubicSpline(T), np.linspace(0, 8, 21), 7)
    z = np.asfortranarray(rng.standard_normal((5, 7)) + 1j * rng.standard_normal((5, 7)))
    J, W = O.maketable(O.kernel(O.CUBIC, 4), np.linspace(0, 8, 21), 7, T)
    out = A * z
    assert out.dtype == np.complex128 and out.shape == (5, 21)
    ref = O.tab_direct(J, W, 7, 2, z.real.copy(order="F")) + 1j * O.tab_direct(J, W, 7, 2, z.imag.copy(order="F"))
    assert_close(out.real, ref.real, T)
    assert_close(out.imag, ref.imag, T)


def test_error_behaviour(li):
    T = np.float64
    A = li.TabulatedInterpolator(li.LinearSpline(T), np.linspace(1, 5, 9), 5)
    with pytest.raises(li.ErrorException):            # src/tabulated.jl:233
        li.vcreate(li.Direct, A, np.zeros((5, 3)))
    with pytest.raises(li.DimensionMismatch):         # src/tabulated.jl:342-343
        A * np.zeros(6)
    with pytest.raises(li.DimensionMismatch):
        li.apply_b(1, li.Direct, A, np.zeros(5), False, 0, np.zeros(8))
    with pytest.raises(li.ErrorException):            # src/tabulated.jl:310
        li.apply(li.Direct, A, 3, np.zeros((5, 3)))
    with pytest.raises(li.ErrorException):            # src/tabulated.jl:83-84,95-96
        li.TabulatedInterpolator(li.LinearSpline(T), np.linspace(1, 5, 9), 0)
    B = li.SparseUnidimensionalInterpolator(li.LinearSpline(T), 2, np.linspace(1, 5, 9), 5)
    with pytest.raises(li.DimensionMismatch):         # src/sparse.jl:646-651
        B * np.zeros((5, 4))
    with pytest.raises(li.ArgumentError):             # src/sparse.jl:577
        li.SparseUnidimensionalInterpolator(li.LinearSpline(T), 1, np.linspace(1, 5, 9), 0)
    Sp = li.SparseInterpolator(li.LinearSpline(T), np.ones((3, 4)), 5)
    with pytest.raises(li.ErrorException):            # src/sparse.jl:176-179
        Sp * np.zeros(4)
    assert (Sp * np.zeros(5)).shape == (3, 4)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T", DTYPES)
def test_sparse_interpolator_nd_positions_and_grid(li, T):
    # test/benchmarks.jl:52-61 shape: 2-D positions, range grid
    rng = np.random.default_rng(8)
    n1, n2, n = 30, 20, 25
    pos = np.array([[1.1 * i + 0.2 * j for j in range(1, n2 + 1)] for i in range(1, n1 + 1)])
    grd = li.Range(pos.min(), pos.max(), length=n)
    A = li.SparseInterpolator(li.CatmullRomSpline(T), pos, grd)
    f = O.fractional_index(pos, grd.first, grd.step, grd.last, n, T)
    J, W = O.maketable(O.kernel(O.CATMULLROM, 4), f, n, T)
    assert np.array_equal(A.J, J.reshape(-1, order="F")) and np.array_equal(A.C, W.reshape(-1, order="F"))
    x = rng.standard_normal(n).astype(T)
    out = A * x
    assert out.shape == (n1, n2)
    assert_close(out.reshape(-1, order="F"), O.tab_direct(J, W, n, 1, x), T)
    y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    assert_close(A.adj * y, O.tab_adjoint(J, W, n, 1, y.reshape(-1, order="F")), T, factor=30)
    assert A.sparse().shape == (n1 * n2, n)


@pytest.mark.parametrize("T", DTYPES)
def test_separable_composition_2d(li, T):
    # test/benchmarks.jl:23-49: H = A1*A2 with SparseUnidimensionalInterpolator, forward and adjoint
    rng = np.random.default_rng(21)
    n1, n2, m1, m2 = 300, 200, 43, 29
    ker = li.CatmullRomSpline(T)
    A1 = li.SparseUnidimensionalInterpolator(ker, 1, np.arange(1, n1 + 1), li.Range(1, n1, length=m1))
    A2 = li.SparseUnidimensionalInterpolator(ker, 2, np.arange(1, n2 + 1), li.Range(1, n2, length=m2))
    H = A1 * A2
    ok = O.kernel(O.CATMULLROM, 4)
    J1, W1 = O.maketable(ok, O.fractional_index(np.arange(1, n1 + 1), 1, (n1 - 1) / (m1 - 1), n1, m1, T), m1, T)
    J2, W2 = O.maketable(ok, O.fractional_index(np.arange(1, n2 + 1), 1, (n2 - 1) / (m2 - 1), n2, m2, T), m2, T)
    x = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
    ref = O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, x))
    for xx in (x, li.to_device(x)):
        assert_close(li.to_host(H * xx), ref, T, factor=4)
        assert_close(li.to_host((A2 * A1) * xx), ref, T, factor=4)
    y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    refa = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, y, skip_zero=False), skip_zero=False)
    for yy in (y, li.to_device(y)):
        assert_close(li.to_host(H.adj * yy), refa, T, factor=16)
    # alpha/beta through the composition
    out = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    want = 0.5 * ref.astype(np.float64) + 2 * out.astype(np.float64)
    li.apply_b(0.5, li.Direct, H, x, False, 2, out)
    assert_close(out, want, T, factor=8)
    np.testing.assert_allclose(li.vdot(H * x, y), li.vdot(x, H.adj * y), rtol=2e-5 if T == np.float32 else 1e-12)


@pytest.mark.parametrize("T", DTYPES)
def test_sep2d_on_the_fly(li, T):
    rng = np.random.default_rng(17)
    m1, m2, n1, n2 = 40, 33, 97, 71
    e1, e2 = KERNELS[4], KERNELS[8]
    x1 = np.linspace(-1.0, m1 + 2.0, n1).astype(T)
    x2 = np.linspace(0.5, m2 + 0.5, n2).astype(T)
    src = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
    k1, k2 = product_kernel(li, e1, O.FLAT, T), product_kernel(li, e2, O.SAFEFLAT, T)
    o1, o2 = oracle_kernel(e1, O.FLAT), oracle_kernel(e2, O.SAFEFLAT)
    for alpha, beta in ((1, 0), (0.5, 2), (0, 2)):
        dst0 = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
        ref = O.sep2d(False, o1, x1, o2, x2, src, (n1, n2), T, alpha, beta, dst0.copy(order="F"))
        got = dst0.copy(order="F")
        li.apply_b(alpha, li.Direct, k1, x1, k2, x2, src, beta, got)
        assert_close(got, ref, T, factor=8)
        y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
        d0 = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
        refa = O.sep2d(True, o1, x1, o2, x2, y, (m1, m2), T, alpha, beta, d0.copy(order="F"))
        gota = d0.copy(order="F")
        li.apply_b(alpha, li.Adjoint, k1, x1, k2, x2, y, beta, gota)
        assert_close(gota, refa, T, factor=32)


@pytest.mark.parametrize("T", DTYPES)
def test_affine(li, T):
    rng = np.random.default_rng(31)
    rows, cols = (61, 47), (50, 45)
    R = (25.3, 22.1) + li.rotate(0.7 * li.AffineTransform2D(T=T) - (30.0, 24.0), 0.2)
    Rc = np.array(li.AffineTransform2D(*R.coefs(), T=T).coefs())
    for e1, e2 in ((KERNELS[0], KERNELS[1]), (KERNELS[4], KERNELS[4]), (KERNELS[2], KERNELS[8]), (KERNELS[9], KERNELS[9])):
        k1, k2 = product_kernel(li, e1, O.FLAT, T), product_kernel(li, e2, O.SAFEFLAT, T)
        o1, o2 = oracle_kernel(e1, O.FLAT), oracle_kernel(e2, O.SAFEFLAT)
        A = li.TwoDimensionalTransformInterpolator(rows, cols, k1, k2, R)
        x = np.asfortranarray(rng.standard_normal(cols).astype(T))
        y = np.asfortranarray(rng.standard_normal(rows).astype(T))
        assert_close(A * x, O.affine2d(False, o1, o2, Rc, rows, cols, x, T), T, factor=8)
        assert_close(A.adj * y, O.affine2d(True, o1, o2, Rc, rows, cols, y, T), T, factor=64)
        for alpha, beta in ((2, 0), (1, 1), (-0.5, 2), (0, 2)):
            d0 = np.asfortranarray(rng.standard_normal(rows).astype(T))
            ref = O.affine2d(False, o1, o2, Rc, rows, cols, x, T, alpha, beta, d0.copy(order="F"))
            got = d0.copy(order="F")
            li.apply_b(alpha, li.Direct, A, x, False, beta, got)
            assert_close(got, ref, T, factor=8)
            c0 = np.asfortranarray(rng.standard_normal(cols).astype(T))
            refa = O.affine2d(True, o1, o2, Rc, rows, cols, y, T, alpha, beta, c0.copy(order="F"))
            gota = c0.copy(order="F")
            li.apply_b(alpha, li.Adjoint, A, y, False, beta, gota)
            assert_close(gota, refa, T, factor=64)


@pytest.mark.parametrize("T", DTYPES)
def test_affine_aligned_axes_take_the_separable_kernels(li, T):
    """src/nonseparable.jl:14 "if axes are aligned, use separable interpolation": a scaling + translation (R.xy = R.yx
    = 0) runs through the separable on-the-fly operator (per-dimension tables, fused kernels); same results as the
    oracle's affine loops and as the library's own affine kernels (exp bit 4 switches the shortcut off)."""
    rng = np.random.default_rng(77)
    rows, cols = (96, 70), (64, 52)
    R = li.AffineTransform2D(0.61, 0.0, 1.7, 0.0, 0.77, -0.4, T=T)     # (xx, xy, x, yx, yy, y): corners hang over the edges
    Rc = np.array(R.coefs())
    for e1, e2 in ((KERNELS[4], KERNELS[8]), (KERNELS[1], KERNELS[3]), (KERNELS[9], KERNELS[9])):
        k1, k2 = product_kernel(li, e1, O.FLAT, T), product_kernel(li, e2, O.SAFEFLAT, T)
        o1, o2 = oracle_kernel(e1, O.FLAT), oracle_kernel(e2, O.SAFEFLAT)
        A = li.TwoDimensionalTransformInterpolator(rows, cols, k1, k2, R)
        x = np.asfortranarray(rng.standard_normal(cols).astype(T))
        y = np.asfortranarray(rng.standard_normal(rows).astype(T))
        dx, dy = li.to_device(x), li.to_device(y)
        got, gota = li.to_host(A * dx), li.to_host(A.adj * dy)
        assert_close(got, O.affine2d(False, o1, o2, Rc, rows, cols, x, T), T)
        assert_close(gota, O.affine2d(True, o1, o2, Rc, rows, cols, y, T), T)
        li.set_option("exp", 4)
        try:
            plain, plaina = li.to_host(A * dx), li.to_host(A.adj * dy)
        finally:
            li.set_option("exp", 0)
        assert_close(plain, O.affine2d(False, o1, o2, Rc, rows, cols, x, T), T)
        assert_close(plaina, O.affine2d(True, o1, o2, Rc, rows, cols, y, T), T)
        d0 = np.asfortranarray(rng.standard_normal(rows).astype(T))
        ref = O.affine2d(False, o1, o2, Rc, rows, cols, x, T, -0.5, 2, d0.copy(order="F"))
        dd = li.to_device(d0)
        li.apply_b(-0.5, li.Direct, A, dx, False, 2, dd)
        assert_close(li.to_host(dd), ref, T)
    with pytest.raises(li.InexactError):   # a NaN coordinate raises like the affine loops (trunc(Int, NaN))
        li.TwoDimensionalTransformInterpolator(rows, cols, k1, k2, li.AffineTransform2D(1.0, 0.0, np.nan, 0.0, 1.0, 0.0, T=T)) * dx


@pytest.mark.parametrize("T", DTYPES)
def test_sep3d_on_the_fly(li, T):
    """Three dimensions with on-the-fly coefficients (li_sep_otf_apply): same numbers as three table-backed operators."""
    rng = np.random.default_rng(78)
    m, n = (24, 20, 12), (40, 31, 19)
    ents = (KERNELS[4], KERNELS[3], KERNELS[1])
    xs = [np.linspace(0.5, m[d] + 0.75, n[d]).astype(T) for d in range(3)]
    kers = [product_kernel(li, ents[d], O.FLAT, T) for d in range(3)]
    JW = [O.maketable(oracle_kernel(ents[d], O.FLAT), xs[d], m[d], T) for d in range(3)]
    src = np.asfortranarray(rng.standard_normal(m).astype(T))
    y = np.asfortranarray(rng.standard_normal(n).astype(T))
    for dev in (False, True):
        s, yy = (li.to_device(src), li.to_device(y)) if dev else (src, y)
        out = li.device_empty(n, T) if dev else np.empty(n, dtype=T, order="F")
        li.apply_b(1, li.Direct, kers[0], xs[0], kers[1], xs[1], kers[2], xs[2], s, 0, out)
        ref = src
        for d in (2, 1, 0):
            ref = O.tab_direct(JW[d][0], JW[d][1], m[d], d + 1, ref)
        assert_close(li.to_host(out) if dev else out, ref, T)
        back = li.device_empty(m, T) if dev else np.empty(m, dtype=T, order="F")
        li.apply_b(1, li.Adjoint, kers[0], xs[0], kers[1], xs[1], kers[2], xs[2], yy, 0, back)
        refa = y
        for d in (0, 1, 2):
            refa = O.tab_adjoint(JW[d][0], JW[d][1], m[d], d + 1, refa)
        assert_close(li.to_host(back) if dev else back, refa, T)
    with pytest.raises(AssertionError):
        li.apply_b(1, li.Direct, kers[0], xs[0], kers[1], xs[1], kers[2], xs[2][:-1], src, 0, np.empty(n, dtype=T, order="F"))


@pytest.mark.parametrize("T", DTYPES)
@pytest.mark.parametrize("theta,scale", [(0.2, 1.0), (0.785, 1.0), (1.45, 1.0), (-2.6, 0.35), (0.4, 3.1)])
def test_affine_tiled_batch(li, T, theta, scale):
    """The tiled forward (TMA-staged source boxes, several images in flight) against the oracle image by image and,
    bit for bit, against the per-pixel kernel (same order of operations)."""
    rng = np.random.default_rng(9)
    rows, cols, nb = (200, 150), (256, 120), 7        # rows of the source are 16-byte aligned: TMA path
    c = ((cols[0] + 1) / 2, (cols[1] + 1) / 2)
    R = c + li.rotate(scale * li.AffineTransform2D(T=T) - ((rows[0] + 1) / 2, (rows[1] + 1) / 2), theta)
    Rc = np.array(li.AffineTransform2D(*R.coefs(), T=T).coefs())
    ker, ok = li.CatmullRomSpline(T), O.kernel(O.CATMULLROM, 4)
    x = np.asfortranarray(rng.standard_normal(cols + (nb,)).astype(T))
    dx = li.to_device(x)
    out = li.device_empty(rows + (nb,), T)
    li.affine_apply(1, li.Direct, ker, ker, R, dx, 0, out, nbatch_dims=1)
    got = li.to_host(out)
    for b in range(nb):
        ref = O.affine2d(False, ok, ok, Rc, rows, cols, np.asfortranarray(x[:, :, b]), T)
        assert_close(got[:, :, b], ref, T, factor=8)
    out2 = li.device_empty(rows + (nb,), T)
    li.affine_apply(1, li.Direct, ker, ker, R, dx, 0, out2, variant=li.VARIANT_ATOMIC, nbatch_dims=1)
    assert np.array_equal(li.to_host(out2), got)
    # general alpha / beta through the tiled kernel
    d0 = np.asfortranarray(rng.standard_normal(rows + (nb,)).astype(T))
    dd = li.to_device(d0)
    li.affine_apply(-0.5, li.Direct, ker, ker, R, dx, 2, dd, nbatch_dims=1)
    assert_close(li.to_host(dd), -0.5 * got.astype(np.float64) + 2 * d0, T, factor=16)


# ------------------------------------------------------------------------------------------------
# Fused 2-D separable kernels (config #2 shape at oracle-friendly sizes)
FUSED_CASES = [
    # (entry1, cond1, entry2, cond2, (m1, m2), (n1, n2), position style)
    (KERNELS[8], O.FLAT, KERNELS[8], O.FLAT, (100, 70), (200, 140), "aligned"),          # Lanczos-4, 2x, config #2 in small
    (KERNELS[8], O.SAFEFLAT, KERNELS[8], O.SAFEFLAT, (333, 129), (517, 261), "overhang"),  # odd sizes, clamping
    (KERNELS[4], O.FLAT, KERNELS[1], O.FLAT, (64, 257), (300, 513), "overhang"),         # S = (4, 2)
    (KERNELS[9], O.FLAT, KERNELS[9], O.SAFEFLAT, (90, 80), (256, 128), "aligned"),       # Lanczos-6
    (KERNELS[1], O.FLAT, KERNELS[4], O.FLAT, (50, 40), (49, 37), "aligned"),             # S = (2, 4), ~1x
    (KERNELS[4], O.FLAT, KERNELS[4], O.FLAT, (300, 200), (43, 29), "aligned"),           # downsampling (unfused fallback)
    (KERNELS[8], O.FLAT, KERNELS[8], O.FLAT, (40, 30), (150, 120), "random"),            # non-monotone positions
    (KERNELS[0], O.FLAT, KERNELS[0], O.FLAT, (20, 30), (64, 33), "overhang"),            # S = (1, 1)
    (KERNELS[2], O.SAFEFLAT, KERNELS[2], O.FLAT, (31, 45), (100, 90), "overhang"),       # S = (3, 3)
    # streaming adjoint (rows of the fine grid 16-byte aligned): several strips and chunks, clamped edges
    (KERNELS[8], O.FLAT, KERNELS[8], O.FLAT, (600, 300), (1200, 600), "aligned"),
    (KERNELS[8], O.SAFEFLAT, KERNELS[8], O.SAFEFLAT, (520, 260), (1040, 519), "overhang"),
    (KERNELS[9], O.FLAT, KERNELS[9], O.FLAT, (300, 100), (600, 200), "aligned"),         # Lanczos-6, 14 taps per thread
    (KERNELS[1], O.FLAT, KERNELS[1], O.SAFEFLAT, (300, 290), (304, 299), "aligned"),     # S = (2, 2), ~1x
    (KERNELS[4], O.FLAT, KERNELS[9], O.FLAT, (270, 40), (400, 77), "overhang"),          # S = (4, 6), 1.5x
    (KERNELS[4], O.FLAT, KERNELS[4], O.FLAT, (300, 200), (44, 28), "aligned"),           # down-sampling: dst rows without any src row
    (KERNELS[8], O.FLAT, KERNELS[8], O.FLAT, (100, 60), (300, 180), "aligned"),          # 3x: too many taps per column -> tile kernels
    (KERNELS[8], O.SAFEFLAT, KERNELS[8], O.FLAT, (260, 300), (520, 304), "overhang"),    # 2x along dim 1, ~1x along dim 2
]


def _positions(style, m, n, rng):
    if style == "aligned":
        return np.linspace(1.0, float(m), n)
    if style == "overhang":
        return np.linspace(-6.5, m + 7.25, n)
    return rng.uniform(-3.0, m + 4.0, n)


@pytest.mark.parametrize("T", DTYPES)
@pytest.mark.parametrize("case", FUSED_CASES, ids=[f"{c[0][0]}-{c[2][0]}-{c[4]}-{c[5]}-{c[6]}" for c in FUSED_CASES])
def test_fused_separable_2d(li, T, case):
    e1, c1, e2, c2, (m1, m2), (n1, n2), style = case
    rng = np.random.default_rng(123)
    x1, x2 = _positions(style, m1, n1, rng), _positions(style, m2, n2, rng)
    A1 = li.TabulatedInterpolator(T, 1, product_kernel(li, e1, c1, T), x1, m1)
    A2 = li.TabulatedInterpolator(T, 2, product_kernel(li, e2, c2, T), x2, m2)
    H = A1 * A2
    J1, W1 = O.maketable(oracle_kernel(e1, c1), x1.astype(T), m1, T)
    J2, W2 = O.maketable(oracle_kernel(e2, c2), x2.astype(T), m2, T)
    src = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
    y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    ref = O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, src))
    refa = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, y))
    scale_a = float(np.max(np.abs(refa)))
    out = H * li.to_device(src)
    assert_close(li.to_host(out), ref, T, factor=8)
    back = H.adj * li.to_device(y)
    assert_close(li.to_host(back), refa, T, factor=64)
    # deterministic
    assert np.array_equal(li.to_host(H.adj * li.to_device(y)), li.to_host(back))
    # alpha / beta
    d0 = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    dd = li.to_device(d0)
    li.apply_b(-1.5, li.Direct, H, li.to_device(src), False, 0.5, dd)
    assert_close(li.to_host(dd), -1.5 * ref.astype(np.float64) + 0.5 * d0, T, factor=16)
    c0 = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
    dc = li.to_device(c0)
    li.apply_b(2.0, li.Adjoint, H, li.to_device(y), False, -1.0, dc)
    assert_close(li.to_host(dc), 2.0 * refa.astype(np.float64) - c0, T, scale=2 * scale_a + 1, factor=64)
    # host arrays go through the same kernels (staged)
    assert_close(H * src, ref, T, factor=8)
    # adjoint identity in Float64 on the host
    np.testing.assert_allclose(li.vdot(out, y), li.vdot(src, back), rtol=3e-5 if T == np.float32 else 1e-11)


@pytest.mark.parametrize("n1", [600, 602])   # streaming kernel / tile kernels
@pytest.mark.parametrize("T", DTYPES)
def test_fused_adjoint_non_finite_footprint(li, T, n1):
    """A non-finite sample reaches exactly the dst elements it feeds in the reference (no 0 * Inf leaks)."""
    rng = np.random.default_rng(5)
    m1, m2, n2 = 300, 70, 140
    x1, x2 = np.linspace(1.0, m1, n1), np.linspace(-2.0, m2 + 3.0, n2)
    ker = li.LanczosKernel(T, 4)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x1, m1), li.TabulatedInterpolator(T, 2, ker, x2, m2)
    J1, W1 = O.maketable(O.kernel(O.LANCZOS, 4), x1.astype(T), m1, T)
    J2, W2 = O.maketable(O.kernel(O.LANCZOS, 4), x2.astype(T), m2, T)
    y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
    for (i1, i2, bad) in ((0, 0, np.inf), (257, 60, np.nan), (n1 - 1, 139, -np.inf), (300, 3, np.inf)):
        y[i1, i2] = bad
    refa = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, y, skip_zero=False), skip_zero=False)
    back = li.to_host((A1 * A2).adj * li.to_device(y))
    assert np.array_equal(np.isfinite(back), np.isfinite(refa))
    ok = np.isfinite(refa)
    assert_close(np.where(ok, back, 0), np.where(ok, refa, 0), T, factor=64)


def test_fused_separable_batch(li):
    T = np.float32
    rng = np.random.default_rng(77)
    m1, m2, n1, n2, nb = 48, 36, 96, 72, 5
    x1, x2 = np.linspace(1, m1, n1), np.linspace(1, m2, n2)
    ker = li.LanczosKernel(T, 4)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x1, m1), li.TabulatedInterpolator(T, 2, ker, x2, m2)
    J1, W1 = O.maketable(O.kernel(O.LANCZOS, 4), x1, m1, T)
    J2, W2 = O.maketable(O.kernel(O.LANCZOS, 4), x2, m2, T)
    src = np.asfortranarray(rng.standard_normal((m1, m2, nb)).astype(T))
    out = li.device_empty((n1, n2, nb), T)
    li.sep_apply_tables(False, [A1.table, A2.table], 1, li.to_device(src), 0, out, nbatch_dims=1)
    ref = O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, src))
    assert_close(li.to_host(out), ref, T, factor=8)
    back = li.device_empty((m1, m2, nb), T)
    li.sep_apply_tables(True, [A1.table, A2.table], 1, out, 0, back, nbatch_dims=1)
    refa = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, ref))
    assert_close(li.to_host(back), refa, T, factor=64)


def test_separable_3d_passes(li):
    # config #4 shape in small: a different kernel per dimension (cubic B-spline, Mitchell-Netravali, linear)
    T = np.float32
    rng = np.random.default_rng(99)
    m, n = (21, 18, 15), (40, 31, 22)
    ents = (KERNELS[3], KERNELS[7], KERNELS[1])
    deltas = (0.37, 0.11, 0.73)
    A, JW = [], []
    for d in range(3):
        x = 1 + np.arange(n[d]) * (m[d] - 1) / (n[d] - 1) * 0.98 + deltas[d]
        A.append(li.TabulatedInterpolator(T, d + 1, product_kernel(li, ents[d], O.FLAT, T), x, m[d]))
        JW.append(O.maketable(oracle_kernel(ents[d], O.FLAT), x.astype(T), m[d], T))
    H = A[0] * A[1] * A[2]
    src = np.asfortranarray(rng.standard_normal(m).astype(T))
    ref = src
    for d in (2, 1, 0):
        ref = O.tab_direct(JW[d][0], JW[d][1], m[d], d + 1, ref)
    out = H * li.to_device(src)
    assert_close(li.to_host(out), ref, T, factor=8)
    y = np.asfortranarray(rng.standard_normal(n).astype(T))
    refa = y
    for d in (0, 1, 2):
        refa = O.tab_adjoint(JW[d][0], JW[d][1], m[d], d + 1, refa)
    assert_close(li.to_host(H.adj * li.to_device(y)), refa, T, factor=64)


# one kernel family per size S = 2, 3, 4, 6, 8
_BY_SIZE = {2: KERNELS[1], 3: KERNELS[2], 4: KERNELS[4], 6: KERNELS[9], 8: KERNELS[10]}


@pytest.mark.parametrize("S2", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("S1", [2, 3, 4, 6, 8])
def test_fused_2d_every_size_pair(li, S1, S2):
    """Every pair of kernel sizes of {2, 3, 4, 6, 8}^2 runs as ONE fused kernel per apply (no silent fall back to the
    per-dimension passes), Direct and Adjoint, an up-sampling and a non-integer rate, against the oracle."""
    T = np.float32
    rng = np.random.default_rng(100 * S1 + S2)
    e1, e2 = _BY_SIZE[S1], _BY_SIZE[S2]
    for (m1, m2, n1, n2) in ((48, 40, 96, 80), (64, 36, 88, 52)):
        x1 = np.linspace(0.25, m1 + 0.5, n1)
        x2 = np.linspace(-0.5, m2 + 1.25, n2)
        A1 = li.TabulatedInterpolator(T, 1, product_kernel(li, e1, O.FLAT, T), x1, m1)
        A2 = li.TabulatedInterpolator(T, 2, product_kernel(li, e2, O.SAFEFLAT, T), x2, m2)
        J1, W1 = O.maketable(oracle_kernel(e1, O.FLAT), x1.astype(T), m1, T)
        J2, W2 = O.maketable(oracle_kernel(e2, O.SAFEFLAT), x2.astype(T), m2, T)
        H = A1 * A2
        src = np.asfortranarray(rng.standard_normal((m1, m2)).astype(T))
        y = np.asfortranarray(rng.standard_normal((n1, n2)).astype(T))
        dsrc, dy = li.to_device(src), li.to_device(y)
        out, back = H * dsrc, H.adj * dy
        assert_close(li.to_host(out), O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, src)), T)
        assert_close(li.to_host(back), O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, y)), T)
        for P, x_in, res in ((li.Direct, dsrc, out), (li.Adjoint, dy, back)):
            c0 = li.launch_count()
            li.apply_b(1, P, H, x_in, False, 0, res)
            assert li.launch_count() - c0 == 1, f"({S1}, {S2}) {P}: {li.launch_count() - c0} launches"


# (name, m, n, kernel entries per dim, condition, position range per dim as (first, last) in units of the source grid,
#  fused forward expected, fused adjoint expected)
FUSED3D_CASES = [
    # config #4 in small: (cubic B-spline, Mitchell-Netravali, linear), rate 0.98, offsets
    ("cfg4", (64, 40, 24), (64, 40, 24), (3, 7, 1), O.FLAT, None, True, True),
    # 2x up-sampling everywhere, positions hanging over both ends of every dimension (clamped windows, folded planes)
    ("up2", (32, 20, 12), (64, 41, 26), (8, 4, 4), O.SAFEFLAT, (-1.5, 2.5), True, False),
    # linear everywhere
    ("lin", (16, 16, 16), (24, 20, 18), (1, 1, 1), O.SAFEFLAT, (0.25, 0.5), True, True),
    # 3x along dim 3: more than 4 planes feed one dst plane -> the adjoint declines (two passes), the forward is fused
    ("up3", (32, 12, 8), (32, 12, 25), (3, 3, 3), O.FLAT, (0.0, 0.0), True, False),
    # mixed sizes (S1, S2) = (2, 4) / S3 = 3
    ("mix", (24, 18, 10), (48, 18, 17), (1, 4, 2), O.FLAT, (0.4, 0.3), True, False),
    # odd n1: destination rows are not vector aligned (scalar stores), the sources still arrive by TMA; the adjoint's
    # source rows are unaligned -> per-dimension passes (found by scratch/fuzz_fused3d.py)
    ("odd", (32, 49, 36), (185, 29, 12), (4, 8, 7), O.FLAT, (0.3, 0.6), True, False),
]


@pytest.mark.parametrize("case", FUSED3D_CASES, ids=[c[0] for c in FUSED3D_CASES])
@pytest.mark.parametrize("T", DTYPES)
def test_fused_3d_single_pass(li, T, case):
    """A1*A2*A3 and its adjoint as ONE kernel each (li_sep3d.cuh) against the oracle's three passes, nest order of the
    reference (A3 first for Direct, A1' first for Adjoint); alpha / beta; a non-finite sample stays in its footprint."""
    name, m, n, kidx, cond, over, fused_fwd, fused_adj = case
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    A, JW = [], []
    for d in range(3):
        if over is None:
            x = 1 + np.arange(n[d]) * 0.98 * (m[d] - 1) / max(n[d] - 1, 1) + (0.37, 0.11, 0.73)[d]
        else:
            x = np.linspace(1 - over[0], m[d] + over[1], n[d])
        ent = KERNELS[kidx[d]]
        A.append(li.TabulatedInterpolator(T, d + 1, product_kernel(li, ent, cond, T), x, m[d]))
        JW.append(O.maketable(oracle_kernel(ent, cond), x.astype(T), m[d], T))
    H = A[0] * A[1] * A[2]

    def oracle_fwd(x, alpha=1, beta=0, dst=None):
        r = x
        for d in (2, 1):
            r = O.tab_direct(JW[d][0], JW[d][1], m[d], d + 1, r)
        return O.tab_direct(JW[0][0], JW[0][1], m[0], 1, r, alpha, beta, dst)

    def oracle_adj(y, alpha=1, beta=0, dst=None):
        r = y
        for d in (0, 1):
            r = O.tab_adjoint(JW[d][0], JW[d][1], m[d], d + 1, r)
        return O.tab_adjoint(JW[2][0], JW[2][1], m[2], 3, r, alpha, beta, dst)

    src = np.asfortranarray(rng.standard_normal(m).astype(T))
    y = np.asfortranarray(rng.standard_normal(n).astype(T))
    dsrc, dy = li.to_device(src), li.to_device(y)
    out = H * dsrc
    back = H.adj * dy
    assert_close(li.to_host(out), oracle_fwd(src), T)
    assert_close(li.to_host(back), oracle_adj(y), T)
    # one launch each once the per-table metadata exists
    for P, x_in, res, fused in ((li.Direct, dsrc, out, fused_fwd), (li.Adjoint, dy, back, fused_adj)):
        c0 = li.launch_count()
        li.apply_b(1, P, H, x_in, False, 0, res)
        assert (li.launch_count() - c0 == 1) == fused, f"{name}: {li.launch_count() - c0} launches"
    # alpha, beta
    d0 = np.asfortranarray(rng.standard_normal(n).astype(T))
    dd = li.to_device(d0)
    li.apply_b(0.5, li.Direct, H, dsrc, False, 2, dd)
    assert_close(li.to_host(dd), oracle_fwd(src, 0.5, 2, d0.copy(order="F")), T)
    b0 = np.asfortranarray(rng.standard_normal(m).astype(T))
    db = li.to_device(b0)
    li.apply_b(-1.5, li.Adjoint, H, dy, False, 0.25, db)
    assert_close(li.to_host(db), oracle_adj(y, -1.5, 0.25, b0.copy(order="F")), T)
    # the same numbers as the per-dimension passes of the library (exp bit 2 switches the fused 3-D kernels off)
    li.set_option("exp", 2)
    try:
        out2, back2 = li.to_host(H * dsrc), li.to_host(H.adj * dy)
    finally:
        li.set_option("exp", 0)
    assert_close(out2, oracle_fwd(src), T)
    assert_close(back2, oracle_adj(y), T)
    # a NaN / Inf sample reaches exactly the elements it reaches in the reference
    for bad in (np.nan, np.inf):
        s2 = src.copy(order="F")
        s2[m[0] // 2, m[1] // 3, m[2] // 2] = bad
        s2[0, 0, m[2] - 1] = bad
        assert np.array_equal(np.isfinite(li.to_host(H * li.to_device(s2))), np.isfinite(np.asarray(oracle_fwd(s2))))
        y2 = y.copy(order="F")
        y2[n[0] // 2, n[1] // 3, n[2] // 2] = bad
        y2[n[0] - 1, 0, 0] = bad
        assert np.array_equal(np.isfinite(li.to_host(H.adj * li.to_device(y2))), np.isfinite(np.asarray(oracle_adj(y2))))


@pytest.mark.parametrize("nframes", [5, 19])   # 19 frames: dealt to concurrent streams with partial accumulators
def test_sep_apply_batch_accumulates_frames(li, nframes):
    # m += sum_k A_k' y_k in one library call (config #5 in small)
    T = np.float32
    rng = np.random.default_rng(55)
    m = 72
    ker, okd = li.SafeFlat(li.LanczosKernel(T, 6)), O.kernel(O.LANCZOS, 6, O.SAFEFLAT)
    xs = np.arange(1, m + 1, dtype=np.float64)
    frames = np.asfortranarray(rng.standard_normal((m, m, nframes)).astype(T))
    dfr = li.to_device(frames)
    m0 = np.asfortranarray(rng.standard_normal((m, m)).astype(T))
    tabs, want = [], m0.astype(np.float64)
    for k in range(nframes):
        d1, d2 = rng.uniform(-8, 8, 2)
        A1, A2 = li.TabulatedInterpolator(T, 1, ker, xs + d1, m), li.TabulatedInterpolator(T, 2, ker, xs + d2, m)
        tabs.append([A1.table, A2.table])
        J1, W1 = O.maketable(okd, (xs + d1).astype(T), m, T)
        J2, W2 = O.maketable(okd, (xs + d2).astype(T), m, T)
        want += O.tab_adjoint(J2, W2, m, 2, O.tab_adjoint(J1, W1, m, 1, np.asfortranarray(frames[:, :, k])))
    results = []
    for _ in range(2):
        acc = li.to_device(m0)
        li.sep_apply_batch(True, tabs, 1, [dfr[:, :, k] for k in range(nframes)], 1, [acc] * nframes)
        results.append(li.to_host(acc))
    assert_close(results[0], want, T, factor=64)
    assert np.array_equal(results[0], results[1])   # deterministic


def test_async_host_streams(li):
    # host-staged applies enqueued on two streams overlap and give the same results as the blocking calls
    T = np.float32
    rng = np.random.default_rng(3)
    m, n = 200, 400
    x = np.linspace(1, m, n)
    ker = li.LanczosKernel(T, 4)
    H = li.TabulatedInterpolator(T, 1, ker, x, m) * li.TabulatedInterpolator(T, 2, ker, x, m)
    src = li.pinned_empty((m, m), T)
    src[...] = rng.standard_normal((m, m))
    yin = li.pinned_empty((n, n), T)
    yin[...] = rng.standard_normal((n, n))
    out, back = li.vcreate(li.Direct, H, src), li.vcreate(li.Adjoint, H, yin)
    ref_out, ref_back = H * np.asarray(src), H.adj * np.asarray(yin)
    sa, sb = li.HostStream(), li.HostStream()
    for _ in range(3):
        out[...] = 0
        back[...] = 0
        with li.async_host(sa):
            li.apply_b(1, li.Direct, H, src, False, 0, out)
        with li.async_host(sb):
            li.apply_b(1, li.Adjoint, H, yin, False, 0, back)
        sa.synchronize()
        sb.synchronize()
        assert np.array_equal(np.asarray(out), ref_out) and np.array_equal(np.asarray(back), ref_back)

"""BASELINE config #2 at full size (Lanczos-4, Float32, 4096^2 -> 8192^2) through size-independent properties, plus a
full-size comparison with the oracle (multi-threaded CPU restatement, a few seconds):
adjoint identity <H x, y> = <x, H' y>, linearity, determinism, and the headline tolerance of 1e-6 relative."""
import numpy as np
import pytest

from util import RO as O   # the oracle, with per-element error bounds attached to its results
from util import assert_close

pytestmark = pytest.mark.gpu

M, N = 4096, 8192


@pytest.fixture(scope="module")
def setup():
    import li_b200 as li
    import torch
    T = np.float32
    ker = li.LanczosKernel(T, 4)
    x = np.linspace(1.0, float(M), N)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x, M), li.TabulatedInterpolator(T, 2, ker, x, M)
    H = A1 * A2
    g = torch.Generator(device="cuda").manual_seed(1234)
    xs = [li.device_empty((M, M), T) for _ in range(2)]
    ys = [li.device_empty((N, N), T) for _ in range(2)]
    for t in xs + ys:
        t.normal_(generator=g)
    return li, torch, H, xs, ys


def test_adjoint_identity_full_size(setup):
    li, torch, H, xs, ys = setup
    Hx = H * xs[0]
    Hty = H.adj * ys[0]
    lhs = li.vdot(Hx, ys[0])
    rhs = li.vdot(xs[0], Hty)
    scale = float(torch.linalg.vector_norm(Hx.double()) * torch.linalg.vector_norm(ys[0].double()))
    assert abs(lhs - rhs) <= 1e-6 * scale


def test_linearity_and_determinism_full_size(setup):
    li, torch, H, xs, ys = setup
    a, b = 0.75, -1.25
    for P, vs in ((li.Direct, xs), (li.Adjoint, ys)):
        comb = a * vs[0] + b * vs[1]
        r = li.apply(P, H, comb)
        r0, r1 = li.apply(P, H, vs[0]), li.apply(P, H, vs[1])
        err = float(torch.max(torch.abs(r - (a * r0 + b * r1))))
        assert err <= 4e-6 * float(torch.max(torch.abs(r)))
        assert torch.equal(li.apply(P, H, comb), r)     # run-to-run identical (no atomics on this path)


def test_against_oracle_full_size(setup):
    li, torch, H, xs, ys = setup
    T = np.float32
    J, W = O.maketable(O.kernel(O.LANCZOS, 4), np.linspace(1.0, float(M), N).astype(T), M, T)
    nt = O.max_threads()
    x = np.asfortranarray(li.to_host(xs[0]))
    ref = O.tab_direct(J, W, M, 1, O.tab_direct(J, W, M, 2, x, nthreads=nt), nthreads=nt)
    assert_close(li.to_host(H * xs[0]), ref, T, factor=4)
    y = np.asfortranarray(li.to_host(ys[0]))
    refa = O.tab_adjoint(J, W, M, 2, O.tab_adjoint(J, W, M, 1, y, nthreads=nt), nthreads=nt)
    assert_close(li.to_host(H.adj * ys[0]), refa, T, factor=4)


def test_host_buffer_pipeline_matches_device_path(setup):
    """Large host-array applies go through the banded upload / kernel / download pipeline: same kernels on the same
    chunks, so the results must equal the device-resident ones bit for bit (blocking and two-stream asynchronous)."""
    li, torch, H, xs, ys = setup
    T = np.float32
    xh = li.pinned_empty((M, M), T)
    xh[...] = li.to_host(xs[0])
    yh_in = li.pinned_empty((N, N), T)
    yh_in[...] = li.to_host(ys[0])
    ref_f, ref_a = li.to_host(H * xs[0]), li.to_host(H.adj * ys[0])
    out = li.vcreate(li.Direct, H, xh)
    back = li.vcreate(li.Adjoint, H, yh_in)
    out[...] = np.nan
    back[...] = np.nan
    li.apply_b(1, li.Direct, H, xh, False, 0, out)
    li.apply_b(1, li.Adjoint, H, yh_in, False, 0, back)
    assert np.array_equal(out, ref_f) and np.array_equal(back, ref_a)
    out[...] = np.nan
    back[...] = np.nan
    sa, sb = li.HostStream(), li.HostStream()
    for _ in range(2):
        with li.async_host(sa):
            li.apply_b(1, li.Direct, H, xh, False, 0, out)
        with li.async_host(sb):
            li.apply_b(1, li.Adjoint, H, yh_in, False, 0, back)
        sa.synchronize()
        sb.synchronize()
        assert np.array_equal(out, ref_f) and np.array_equal(back, ref_a)
    # pageable (not pinned) host arrays take the same path
    xp = np.asfortranarray(np.array(xh))
    outp = np.full((N, N), np.nan, dtype=T, order="F")
    li.apply_b(1, li.Direct, H, xp, False, 0, outp)
    assert np.array_equal(outp, ref_f)


# ------------------------------------------------------------------------------------------------
# The other BASELINE.json configurations at their full sizes.  Where the whole array is too large for the CPU oracle
# (1024^3; 512 frames), the oracle runs on CROPS of the result: the tables are explicit (J, W), so the rows of a table
# that produce (Direct) or feed (Adjoint) a block of the result, with the block of the source they touch, are the same
# arithmetic in the same order -- not an approximation of the full problem.

def _crop_direct(J, W, lo, hi):
    """rows [lo, hi) of a table: (J shifted to the block of columns they read, W, first column, number of columns)"""
    Jc = J[:, lo:hi]
    c0, c1 = int(Jc.min()), int(Jc.max())
    return np.asfortranarray(Jc - (c0 - 1)), np.asfortranarray(W[:, lo:hi]), c0 - 1, c1 - c0 + 1


def _crop_adjoint(J, W, lo, hi):
    """the rows of a table with a tap on columns [lo, hi) (0-based), restricted to those taps: rows outside keep a zero
    weight on a clamped column index, i.e. they add an exact zero exactly where the reference adds their real term to
    an element outside the block"""
    rows = np.nonzero(((J - 1 >= lo) & (J - 1 < hi)).any(axis=0))[0]
    r0, r1 = int(rows.min()), int(rows.max()) + 1
    Jc = J[:, r0:r1] - 1 - lo
    inside = (Jc >= 0) & (Jc < hi - lo)
    Wc = np.where(inside, W[:, r0:r1], 0).astype(W.dtype)
    return np.asfortranarray(np.clip(Jc, 0, hi - lo - 1) + 1), np.asfortranarray(Wc), r0, r1


def test_config1_sparse_1d_float64_full_size():
    """#1: 1-D SparseInterpolator, CatmullRom, Float64, 10^4 -> 2*10^4 samples, monotone and shuffled positions."""
    import li_b200 as li
    T = np.float64
    ncols, nrows = 10_000, 20_000
    rng = np.random.default_rng(101)
    for shuffled in (False, True):
        pos = np.linspace(1.0, float(ncols), nrows)
        if shuffled:
            rng.shuffle(pos)
        for B, cond in ((li.Flat, O.FLAT), (li.SafeFlat, O.SAFEFLAT)):
            A = li.SparseInterpolator(li.CatmullRomSpline(T, B), pos, ncols)
            J, W = O.maketable(O.kernel(O.CATMULLROM, 4, cond), pos, ncols, T)
            assert np.array_equal(A.J.reshape(4, -1, order="F"), J) and np.array_equal(A.C.reshape(4, -1, order="F"), W)
            x, y = rng.standard_normal(ncols), rng.standard_normal(nrows)
            assert_close(li.to_host(A * li.to_device(x)), O.tab_direct(J, W, ncols, 1, x), T)
            ref = O.tab_adjoint(J, W, ncols, 1, y)
            for variant in (li.VARIANT_GATHER, li.VARIANT_ATOMIC):
                A.variant = variant
                assert_close(li.to_host(A.adj * li.to_device(y)), ref, T)


def test_config3_affine_rotation_full_size():
    """#3: TwoDimensionalTransformInterpolator, rotation by 0.2 rad about the centre, CatmullRom, Float32 8192^2 (one
    image of the batch; corners fall outside: clamped taps)."""
    import li_b200 as li
    T = np.float32
    n = 8192
    rng = np.random.default_rng(103)
    c = (n + 1) / 2
    R = (c, c) + li.rotate(li.AffineTransform2D(T=T) - (c, c), 0.2)
    Rc = np.array(li.AffineTransform2D(*R.coefs(), T=T).coefs())
    ker = li.CatmullRomSpline(T)
    ok = O.kernel(O.CATMULLROM, 4)
    A = li.TwoDimensionalTransformInterpolator((n, n), (n, n), ker, ker, R)
    x = np.asfortranarray(rng.standard_normal((n, n)).astype(T))
    assert_close(li.to_host(A * li.to_device(x)), O.affine2d(False, ok, ok, Rc, (n, n), (n, n), x, T, nthreads=O.max_threads()), T)
    y = np.asfortranarray(rng.standard_normal((n, n)).astype(T))
    refa = O.affine2d(True, ok, ok, Rc, (n, n), (n, n), y, T)
    got = li.to_host(A.adj * li.to_device(y))
    assert_close(got, refa, T)
    assert np.array_equal(got, li.to_host(A.adj * li.to_device(y))), "the default adjoint must be run-to-run identical"


def test_config4_separable_3d_full_size():
    """#4: 1024^3 Float32, (cubic B-spline, Mitchell-Netravali, linear) on dims (1, 2, 3): Direct and Adjoint on the GPU
    at full size, the oracle on crops of the result."""
    import li_b200 as li
    import torch
    T = np.float32
    n = 1024
    kers = (li.CubicSpline(T), li.MitchellNetravaliSpline(T), li.LinearSpline(T))
    oks = (O.kernel(O.CUBIC, 4), O.kernel(O.MITCHELLNETRAVALI, 4, O.FLAT, (1 / 3, 1 / 3)), O.kernel(O.LINEAR, 2))
    xs = [1 + np.arange(n) * 0.98 + d for d in (0.37, 0.11, 0.73)]
    As = [li.TabulatedInterpolator(T, d + 1, kers[d], xs[d], n) for d in range(3)]
    tabs = [O.maketable(oks[d], xs[d].astype(T), n, T) for d in range(3)]
    H = As[0] * As[1] * As[2]
    src = li.device_empty((n, n, n), T)
    src.normal_(generator=torch.Generator(device="cuda").manual_seed(104))
    out = H * src
    rng = np.random.default_rng(104)
    for _ in range(3):
        lo = [int(rng.integers(0, n - 96)) for _ in range(3)]
        if _ == 0:
            lo = [0, n - 96, n - 96]          # a corner: clamped windows
        hi = [l + 96 for l in lo]
        crops = [_crop_direct(*tabs[d], lo[d], hi[d]) for d in range(3)]
        blk = tuple(slice(c[2], c[2] + c[3]) for c in crops)
        xb = np.asfortranarray(li.to_host(src[blk]))
        ref = xb
        for d in (2, 1, 0):                    # the reference composes right to left: A3 first
            ref = O.tab_direct(crops[d][0], crops[d][1], crops[d][3], d + 1, ref)
        assert_close(li.to_host(out[tuple(slice(l, h) for l, h in zip(lo, hi))]), ref, T)
    back = H.adj * out
    for k in range(2):
        lo = [0, 0, n - 64] if k == 0 else [int(rng.integers(0, n - 64)) for _ in range(3)]
        hi = [l + 64 for l in lo]
        crops = [_crop_adjoint(*tabs[d], lo[d], hi[d]) for d in range(3)]
        blk = tuple(slice(c[2], c[3]) for c in crops)
        yb = np.asfortranarray(li.to_host(out[blk]))
        ref = yb
        for d in (0, 1, 2):                    # (A1*A2*A3)' = A3'*A2'*A1': A1' first
            ref = O.tab_adjoint(crops[d][0], crops[d][1], hi[d] - lo[d], d + 1, ref, skip_zero=False)
        assert_close(li.to_host(back[tuple(slice(l, h) for l, h in zip(lo, hi))]), ref, T)


def test_config5_multi_frame_accumulate_full_size():
    """#5: 512 frames of 2048^2, per-frame Lanczos-6 SafeFlat operators shifted by U(-8, 8), accumulated into one
    2048^2 model by the multi-frame kernel; the oracle (frames one after the other, as the reference) on crops of
    the model, corners and edges included."""
    import li_b200 as li
    import torch
    T = np.float32
    m, nf = 2048, 512
    ker = li.SafeFlat(li.LanczosKernel(T, 6))
    ok = O.kernel(O.LANCZOS, 6, O.SAFEFLAT)
    xs = np.arange(1, m + 1, dtype=np.float64)
    shifts = np.random.default_rng(5).uniform(-8, 8, (nf, 2))
    tabs = [(li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table, li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table)
            for k in range(nf)]
    frames = li.device_empty((m, m, nf), T)
    frames.normal_(generator=torch.Generator(device="cuda").manual_seed(105))
    model = li.device_empty((m, m), T)
    batch = li.FrameBatch(tabs, [frames[:, :, k] for k in range(nf)])
    li.adjoint_accumulate(batch, model)
    got = li.to_host(model)
    li.adjoint_accumulate(batch, model)
    assert np.array_equal(li.to_host(model), got), "run-to-run identical"
    for lo in ((0, 0), (m - 48, m - 48), (0, 1000), (977, 1311)):
        hi = (lo[0] + 48, lo[1] + 48)
        acc = None
        for k in range(nf):
            t1 = _crop_adjoint(*O.maketable(ok, (xs + shifts[k, 0]).astype(T), m, T), lo[0], hi[0])
            t2 = _crop_adjoint(*O.maketable(ok, (xs + shifts[k, 1]).astype(T), m, T), lo[1], hi[1])
            yb = np.asfortranarray(li.to_host(frames[t1[2]:t1[3], t2[2]:t2[3], k]))
            r = O.tab_adjoint(t2[0], t2[1], 48, 2, O.tab_adjoint(t1[0], t1[1], 48, 1, yb, skip_zero=False), skip_zero=False)
            if acc is None:
                acc, bound, terms = np.asarray(r).astype(T), r.bound.copy(), r.terms.copy()
            else:
                acc = (acc + np.asarray(r)).astype(T)
                bound += r.bound
                terms += r.terms
        from util import _wrap
        assert_close(got[lo[0]:hi[0], lo[1]:hi[1]], _wrap(acc, bound, terms), T)

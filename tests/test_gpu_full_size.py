"""BASELINE config #2 at full size (Lanczos-4, Float32, 4096^2 -> 8192^2) through size-independent properties, plus a
full-size comparison with the oracle (multi-threaded CPU restatement, a few seconds):
adjoint identity <H x, y> = <x, H' y>, linearity, determinism, and the headline tolerance of 1e-6 relative."""
import numpy as np
import pytest

import li_oracle as O
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

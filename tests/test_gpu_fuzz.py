"""Randomised cross-check of the fused 2-D separable kernels (streaming forward / adjoint, tile kernels) against the
per-dimension passes, both through the C ABI on the GPU: random sizes, kernels, boundary conditions, position
patterns (aligned, overhanging, random sorted, constant step), element types, batch sizes, alpha / beta.
(`scratch/fuzz_fused.py` is the same generator with a case count on the command line.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def kernels(T):
    return [li.LinearSpline(T), li.CatmullRomSpline(T), li.CubicSpline(T), li.LanczosKernel(T, 4), li.LanczosKernel(T, 6), li.QuadraticSpline(T),
            li.SafeFlat(li.LanczosKernel(T, 4)), li.SafeFlat(li.CatmullRomSpline(T)), li.SafeFlat(li.LanczosKernel(T, 6))]

def _ratio(got, want, T):
    """worst |got - want| / tolerance over the elements: tolerance = the stated rtol (1e-6 / 1e-12) times the element's
    own error scale sum |w * src| (times terms / 32 where more than 32 products are summed), see tests/util.py"""
    from util import RTOL, TERMS0
    tol = RTOL[np.dtype(T)] * want.bound * np.maximum(1.0, want.terms / TERMS0) + np.finfo(np.float64).tiny
    return float(np.max(np.abs(got.astype(np.float64) - np.asarray(want, dtype=np.float64)) / tol))


def positions(style, m, n, rng):
    if style == 0: return np.linspace(1.0, float(m), n)
    if style == 1: return np.linspace(-rng.uniform(0, 9), m + rng.uniform(0, 9), n)
    if style == 2: return np.sort(rng.uniform(-2.0, m + 3.0, n))
    step = rng.choice([0.5, 1.0, 0.98, 0.49994, 2.0, 1.5, 0.3333])
    return rng.uniform(-3, 3) + 1 + np.arange(n) * step

def main(ncases, seed):
    global li
    import li_b200 as li
    from util import RO as O
    rng = np.random.default_rng(seed)
    worst = 0.0
    for case in range(ncases):
        T = np.float32 if rng.random() < 0.7 else np.float64
        ks = kernels(T)
        k1, k2 = ks[rng.integers(len(ks))], ks[rng.integers(len(ks))]
        m1, m2 = int(rng.integers(6, 700)), int(rng.integers(6, 400))
        n1, n2 = int(rng.integers(4, 900)), int(rng.integers(4, 500))
        if rng.random() < 0.7: n1 = (n1 + 3) & ~3
        if rng.random() < 0.5: m1 = (m1 + 3) & ~3
        nb = int(rng.choice([1, 1, 1, 3]))
        s1, s2 = int(rng.integers(4)), int(rng.integers(4))
        x1, x2 = positions(s1, m1, n1, rng), positions(s2, m2, n2, rng)
        A1, A2 = li.TabulatedInterpolator(T, 1, k1, x1, m1), li.TabulatedInterpolator(T, 2, k2, x2, m2)
        src = li.to_device(np.asfortranarray(rng.standard_normal((m1, m2, nb)).astype(T)))
        y = li.to_device(np.asfortranarray(rng.standard_normal((n1, n2, nb)).astype(T)))
        alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.choice([0.0, 1.0, -0.5])))
        d0 = np.asfortranarray(rng.standard_normal((n1, n2, nb)).astype(T))
        c0 = np.asfortranarray(rng.standard_normal((m1, m2, nb)).astype(T))
        out, back = li.to_device(d0), li.to_device(c0)
        li.sep_apply_tables(False, [A1.table, A2.table], alpha, src, beta, out, nbatch_dims=1)
        li.sep_apply_tables(True, [A1.table, A2.table], alpha, y, beta, back, nbatch_dims=1)
        # the oracle on the product's own tables (bit-equal to the oracle's, tests/test_gpu_parity.py): the reference's
        # two passes (A2 then A1; A1' then A2'), alpha and beta in the last one
        J1, W1, J2, W2 = A1.J, A1.W, A2.J, A2.W
        hs, hy = li.to_host(src), li.to_host(y)
        ref = O.tab_direct(J1, W1, m1, 1, O.tab_direct(J2, W2, m2, 2, hs), alpha, beta, d0.copy(order="F"))
        refa = O.tab_adjoint(J2, W2, m2, 2, O.tab_adjoint(J1, W1, m1, 1, hy), alpha, beta, c0.copy(order="F"))
        for got, want, name in ((li.to_host(out), ref, "forward"), (li.to_host(back), refa, "adjoint")):
            err = _ratio(got, want, T)
            worst = max(worst, err)
            if not np.isfinite(err) or err > 1:
                print(f"MISMATCH case {case} {name}: err / tol {err:.3e} T={np.dtype(T).name} k1={type(k1).__name__}/{k1.S} k2={type(k2).__name__}/{k2.S} "
                      f"m=({m1},{m2}) n=({n1},{n2}) nb={nb} styles=({s1},{s2}) alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0



@pytest.mark.parametrize("seed", [11, 12, 13])
def test_fused_kernels_match_oracle(seed):
    """Random kernels, boundary conditions, sizes, rates (up-, down-, same-rate sampling, unsorted positions), batches,
    alpha / beta, both element types: the fused 2-D dispatch (streaming, tile and fallback kernels) against the CPU
    oracle's two passes, per element at the stated tolerance."""
    assert main(60, seed) == 0


def main3d(ncases, seed):
    """Three tables on dimensions 1-3: the fused 3-D kernels (or the passes they fall back to) against the oracle's three
    passes in the reference's order (A3 first for Direct, A1' first for Adjoint)."""
    global li
    import li_b200 as li
    from util import RO as O
    rng = np.random.default_rng(seed)
    worst = 0.0
    for case in range(ncases):
        T = np.float32 if rng.random() < 0.7 else np.float64
        ks = kernels(T) + [li.SafeFlat(li.LinearSpline(T)), li.MitchellNetravaliSpline(T), li.LanczosKernel(T, 8)]
        k = [ks[rng.integers(len(ks))] for _ in range(3)]
        m = [int(rng.integers(6, 120)), int(rng.integers(6, 60)), int(rng.integers(4, 30))]
        n = [int(rng.integers(4, 150)), int(rng.integers(4, 70)), int(rng.integers(3, 36))]
        if rng.random() < 0.8: m[0] = (m[0] + 3) & ~3
        if rng.random() < 0.7: n[0] = (n[0] + 3) & ~3
        xs = [positions(int(rng.integers(4)), m[d], n[d], rng) for d in range(3)]
        A = [li.TabulatedInterpolator(T, d + 1, k[d], xs[d], m[d]) for d in range(3)]
        tabs = [a.table for a in A]
        hs = np.asfortranarray(rng.standard_normal(m).astype(T))
        hy = np.asfortranarray(rng.standard_normal(n).astype(T))
        alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.choice([0.0, 1.0, -0.5])))
        d0 = np.asfortranarray(rng.standard_normal(n).astype(T))
        c0 = np.asfortranarray(rng.standard_normal(m).astype(T))
        out, back = li.to_device(d0), li.to_device(c0)
        li.sep_apply_tables(False, tabs, alpha, li.to_device(hs), beta, out)
        li.sep_apply_tables(True, tabs, alpha, li.to_device(hy), beta, back)
        ref = O.tab_direct(A[2].J, A[2].W, m[2], 3, hs)
        ref = O.tab_direct(A[1].J, A[1].W, m[1], 2, ref)
        ref = O.tab_direct(A[0].J, A[0].W, m[0], 1, ref, alpha, beta, d0.copy(order="F"))
        refa = O.tab_adjoint(A[0].J, A[0].W, m[0], 1, hy)
        refa = O.tab_adjoint(A[1].J, A[1].W, m[1], 2, refa)
        refa = O.tab_adjoint(A[2].J, A[2].W, m[2], 3, refa, alpha, beta, c0.copy(order="F"))
        for got, want, name in ((li.to_host(out), ref, "forward"), (li.to_host(back), refa, "adjoint")):
            err = _ratio(got, want, T)
            worst = max(worst, err)
            if not np.isfinite(err) or err > 1:
                print(f"MISMATCH 3-D case {case} {name}: err / tol {err:.3e} T={np.dtype(T).name} "
                      f"k={[type(q).__name__ + '/' + str(q.S) for q in k]} m={m} n={n} alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} 3-D cases ok, worst err / tol = {worst:.3f}")
    return 0


@pytest.mark.parametrize("seed", [31, 32])
def test_fused_3d_kernels_match_oracle(seed):
    """Random kernels, sizes (odd first dimensions included), rates, clamped edges, alpha / beta, both element types: three
    tables on dimensions 1-3 against the CPU oracle, per element at the stated tolerance."""
    assert main3d(40, seed) == 0


@pytest.mark.parametrize("seed", [21, 22])
def test_affine_tiled_forward_is_bit_identical_to_per_pixel_kernel(seed):
    """Random affine transforms (rotation, anisotropic scale, shear, large offsets), sizes, kernels and batches: the
    tiled forward (TMA-staged boxes, or plain staging for unaligned rows) must reproduce the per-pixel kernel bit for
    bit; the adjoint (atomics, order not fixed) must agree with the adjoint identity <A x, y> = <x, A' y>."""
    import li_b200 as li
    rng = np.random.default_rng(seed)
    for case in range(40):
        T = np.float32 if rng.random() < 0.7 else np.float64
        kers = [li.LinearSpline(T), li.CatmullRomSpline(T), li.SafeFlat(li.CubicSpline(T)), li.LanczosKernel(T, 4),
                li.QuadraticSpline(T), li.SafeFlat(li.LanczosKernel(T, 6))]
        k1, k2 = kers[rng.integers(len(kers))], kers[rng.integers(len(kers))]
        rows = (int(rng.integers(5, 300)), int(rng.integers(5, 200)))
        cols = (int(rng.integers(5, 300)), int(rng.integers(5, 200)))
        if rng.random() < 0.6:
            cols = ((cols[0] + 3) & ~3, cols[1])
        nb = int(rng.choice([1, 2, 7, 9]))
        th = rng.uniform(-np.pi, np.pi)
        sx, sy = np.exp(rng.uniform(-1.2, 1.2, 2))
        sh = rng.uniform(-0.5, 0.5)
        L = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) @ np.array([[sx, sh], [0.0, sy]])
        off = np.array([cols[0] / 2, cols[1] / 2]) - L @ np.array([rows[0] / 2, rows[1] / 2]) + rng.uniform(-20, 20, 2)
        R = li.AffineTransform2D(L[0, 0], L[0, 1], off[0], L[1, 0], L[1, 1], off[1], T=T)
        x = li.to_device(np.asfortranarray(rng.standard_normal(cols + (nb,)).astype(T)))
        y = li.to_device(np.asfortranarray(rng.standard_normal(rows + (nb,)).astype(T)))
        o1, o2 = li.device_empty(rows + (nb,), T), li.device_empty(rows + (nb,), T)
        li.affine_apply(1, li.Direct, k1, k2, R, x, 0, o1, nbatch_dims=1)
        li.affine_apply(1, li.Direct, k1, k2, R, x, 0, o2, variant=li.VARIANT_ATOMIC, nbatch_dims=1)
        a, b = li.to_host(o1), li.to_host(o2)
        assert np.array_equal(a, b), f"case {case}: tiled forward differs (rows={rows} cols={cols} nb={nb} T={np.dtype(T).name})"
        back = li.device_empty(cols + (nb,), T)
        li.affine_apply(1, li.Adjoint, k1, k2, R, y, 0, back, nbatch_dims=1)
        lhs = float(np.vdot(a.astype(np.float64), li.to_host(y).astype(np.float64)))
        rhs = float(np.vdot(li.to_host(x).astype(np.float64), li.to_host(back).astype(np.float64)))
        scale = float(np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(li.to_host(y).astype(np.float64))) + 1e-30
        assert abs(lhs - rhs) <= (2e-5 if T == np.float32 else 1e-11) * scale, f"case {case}: adjoint identity"


def tab_main(ncases, seed):
    import li_b200 as li
    from util import RO as O
    from util import KERNELS, CONDITIONS, oracle_kernel, product_kernel
    rng = np.random.default_rng(seed)
    worst = 0.0
    for case in range(ncases):
        T = np.float32 if rng.random() < 0.6 else np.float64
        entry = KERNELS[rng.integers(len(KERNELS))]
        cname, cond = CONDITIONS[rng.integers(2)]
        nd = int(rng.integers(1, 4))
        shape = tuple(int(rng.choice([rng.integers(2, 70), 4 * rng.integers(1, 40), 96, 128, 200])) for _ in range(nd))
        d = int(rng.integers(1, nd + 1))
        ncols = shape[d - 1]
        nrows = int(rng.integers(1, 260))
        style = int(rng.integers(4))
        if style == 0: xs = np.linspace(1, ncols, nrows)
        elif style == 1: xs = np.linspace(-rng.uniform(0, 12), ncols + rng.uniform(0, 12), nrows)
        elif style == 2: xs = np.sort(rng.uniform(-3, ncols + 4, nrows))
        else: xs = rng.uniform(-3, ncols + 4, nrows)
        J, W = O.maketable(oracle_kernel(entry, cond), xs.astype(T), ncols, T)
        if np.isnan(W).any(): continue
        A = li.TabulatedInterpolator(T, d, product_kernel(li, entry, cond, T), xs, ncols)
        src = np.asfortranarray(rng.standard_normal(shape).astype(T))
        oshape = tuple(nrows if i == d - 1 else n for i, n in enumerate(shape))
        alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.choice([0.0, 1.0, -0.5])))
        d0 = np.asfortranarray(rng.standard_normal(oshape).astype(T))
        ref = O.tab_direct(J, W, ncols, d, src, alpha, beta, d0.copy(order="F"))
        dd = li.to_device(d0)
        li.apply_b(alpha, li.Direct, A, li.to_device(src), False, beta, dd)
        y = np.asfortranarray(rng.standard_normal(oshape).astype(T))
        c0 = np.asfortranarray(rng.standard_normal(shape).astype(T))
        refa = O.tab_adjoint(J, W, ncols, d, y, alpha, beta, c0.copy(order="F"))
        dc = li.to_device(c0)
        li.apply_b(alpha, li.Adjoint, A, li.to_device(y), False, beta, dc)
        for got, want, name in ((li.to_host(dd), ref, "direct"), (li.to_host(dc), refa, "adjoint")):
            err = _ratio(got, want, T)
            worst = max(worst, err)
            if not np.isfinite(err) or err > 1:
                print(f"MISMATCH case {case} {name}: err / tol {err:.3e} T={np.dtype(T).name} {entry[0]} {cname} shape={shape} d={d} nrows={nrows} style={style} alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0



@pytest.mark.parametrize("seed", [31, 32])
def test_per_dimension_applies_match_oracle(seed):
    """Random shapes, dimensions, kernels (S = 1 .. 8), boundary conditions, sorted and unsorted positions, element
    types, alpha / beta: every routing of li_tab_apply (gather, streaming, fused with an identity factor) against the
    CPU oracle.  (`scratch/fuzz_tab.py` is the same generator with a case count on the command line.)"""
    assert tab_main(80, seed) == 0


def affine_main(ncases, seed):
    import li_b200 as li
    from util import RO as O
    from util import KERNELS, CONDITIONS, oracle_kernel, product_kernel
    rng = np.random.default_rng(seed)
    worst = 0.0
    for case in range(ncases):
        T = np.float32 if rng.random() < 0.6 else np.float64
        e1, e2 = KERNELS[rng.integers(len(KERNELS))], KERNELS[rng.integers(len(KERNELS))]
        (c1n, c1), (c2n, c2) = CONDITIONS[rng.integers(2)], CONDITIONS[rng.integers(2)]
        rows = (int(rng.integers(3, 150)), int(rng.integers(3, 120)))
        cols = (int(rng.choice([rng.integers(3, 150), 4 * rng.integers(1, 40)])), int(rng.integers(3, 120)))
        th = rng.uniform(-np.pi, np.pi); sx, sy = np.exp(rng.uniform(-1.0, 1.0, 2)); sh = rng.uniform(-0.4, 0.4)
        L = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]) @ np.array([[sx, sh], [0.0, sy]])
        off = np.array([cols[0] / 2, cols[1] / 2]) - L @ np.array([rows[0] / 2, rows[1] / 2]) + rng.uniform(-15, 15, 2)
        R = li.AffineTransform2D(L[0, 0], L[0, 1], off[0], L[1, 0], L[1, 1], off[1], T=T)
        Rc = np.array(R.coefs())
        k1, k2 = product_kernel(li, e1, c1, T), product_kernel(li, e2, c2, T)
        o1, o2 = oracle_kernel(e1, c1), oracle_kernel(e2, c2)
        x = np.asfortranarray(rng.standard_normal(cols).astype(T))
        y = np.asfortranarray(rng.standard_normal(rows).astype(T))
        alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.choice([0.0, 1.0, -0.5])))
        d0 = np.asfortranarray(rng.standard_normal(rows).astype(T)); c0 = np.asfortranarray(rng.standard_normal(cols).astype(T))
        ref = O.affine2d(False, o1, o2, Rc, rows, cols, x, T, alpha, beta, d0.copy(order="F"))
        refa = O.affine2d(True, o1, o2, Rc, rows, cols, y, T, alpha, beta, c0.copy(order="F"))
        if not (np.isfinite(ref).all() and np.isfinite(refa).all()): continue
        dd, dc = li.to_device(d0), li.to_device(c0)
        li.affine_apply(alpha, li.Direct, k1, k2, R, li.to_device(x), beta, dd)
        li.affine_apply(alpha, li.Adjoint, k1, k2, R, li.to_device(y), beta, dc)
        for got, want, name in ((li.to_host(dd), ref, "direct"), (li.to_host(dc), refa, "adjoint")):
            err = _ratio(got, want, T)
            worst = max(worst, err)
            if not np.isfinite(err) or err > 1:
                print(f"MISMATCH case {case} {name}: err / tol {err:.3e} T={np.dtype(T).name} {e1[0]}/{c1n} {e2[0]}/{c2n} rows={rows} cols={cols} alpha={alpha} beta={beta} R={Rc}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0



@pytest.mark.parametrize("seed", [41])
def test_affine_operator_matches_oracle_on_random_transforms(seed):
    """Random rotations / scales / shears / offsets, every pair of kernels and boundary conditions, both element types,
    alpha / beta: tiled forward and atomic adjoint against the CPU oracle."""
    assert affine_main(120, seed) == 0

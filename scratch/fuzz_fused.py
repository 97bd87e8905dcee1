"""Randomised cross-check of the fused 2-D separable kernels against the per-dimension passes (both on the GPU).
The fused path is toggled per call with li.set_option("no_fused", ...)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li

def kernels(T):
    return [li.LinearSpline(T), li.CatmullRomSpline(T), li.CubicSpline(T), li.LanczosKernel(T, 4), li.LanczosKernel(T, 6), li.QuadraticSpline(T),
            li.SafeFlat(li.LanczosKernel(T, 4)), li.SafeFlat(li.CatmullRomSpline(T)), li.SafeFlat(li.LanczosKernel(T, 6)),
            li.LanczosKernel(T, 8), li.SafeFlat(li.LanczosKernel(T, 8)), li.SafeFlat(li.QuadraticSpline(T)), li.SafeFlat(li.LinearSpline(T))]

def positions(style, m, n, rng):
    if style == 0: return np.linspace(1.0, float(m), n)
    if style == 1: return np.linspace(-rng.uniform(0, 9), m + rng.uniform(0, 9), n)
    if style == 2: return np.sort(rng.uniform(-2.0, m + 3.0, n))
    step = rng.choice([0.5, 1.0, 0.98, 0.49994, 2.0, 1.5, 0.3333])
    return rng.uniform(-3, 3) + 1 + np.arange(n) * step

def main(ncases, seed):
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
        res = {}
        for mode in ("fused", "passes"):
            li.set_option("no_fused", 1 if mode == "passes" else 0)
            out = li.to_device(np.asfortranarray(np.ones((n1, n2, nb), dtype=T)))
            back = li.to_device(np.asfortranarray(np.ones((m1, m2, nb), dtype=T)))
            li.sep_apply_tables(False, [A1.table, A2.table], alpha, src, beta, out, nbatch_dims=1)
            li.sep_apply_tables(True, [A1.table, A2.table], alpha, y, beta, back, nbatch_dims=1)
            res[mode] = (li.to_host(out).astype(np.float64), li.to_host(back).astype(np.float64))
        li.set_option("no_fused", 0)
        tol = (2e-5 if T == np.float32 else 1e-11)
        for idx, name in ((0, "forward"), (1, "adjoint")):
            a, b = res["fused"][idx], res["passes"][idx]
            scale = max(np.max(np.abs(b)), 1e-30)
            err = np.max(np.abs(a - b)) / scale
            worst = max(worst, err / tol)
            if not np.isfinite(err) or err > tol:
                print(f"MISMATCH case {case} {name}: err {err:.3e} T={np.dtype(T).name} k1={type(k1).__name__}/{k1.S} k2={type(k2).__name__}/{k2.S} "
                      f"m=({m1},{m2}) n=({n1},{n2}) nb={nb} styles=({s1},{s2}) alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0

if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 1))

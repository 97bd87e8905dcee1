"""Randomised cross-check of the fused 3-D separable kernels (li_sep3d.cuh) against the library's own two-pass path
(option exp bit 2), both on the GPU: shapes, kernel sizes, rates, clamped edges, batches, alpha / beta, element types."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li


def kernels(T):
    return [li.LinearSpline(T), li.CatmullRomSpline(T), li.CubicSpline(T), li.LanczosKernel(T, 4), li.LanczosKernel(T, 6), li.QuadraticSpline(T),
            li.SafeFlat(li.LanczosKernel(T, 4)), li.SafeFlat(li.CatmullRomSpline(T)), li.SafeFlat(li.LinearSpline(T)), li.MitchellNetravaliSpline(T)]


def positions(style, m, n, rng):
    if style == 0: return np.linspace(1.0, float(m), n)
    if style == 1: return np.linspace(-rng.uniform(0, 4), m + rng.uniform(0, 4), n)
    if style == 2: return np.sort(rng.uniform(-1.0, m + 2.0, n))
    step = rng.choice([0.5, 1.0, 0.98, 0.49994, 2.0, 1.5, 0.3333])
    return rng.uniform(-2, 2) + 1 + np.arange(n) * step


def main(ncases, seed):
    rng = np.random.default_rng(seed)
    worst, fused_f, fused_a = 0.0, 0, 0
    for case in range(ncases):
        T = np.float32 if rng.random() < 0.7 else np.float64
        ks = kernels(T)
        k = [ks[rng.integers(len(ks))] for _ in range(3)]
        m = [int(rng.integers(6, 160)), int(rng.integers(6, 90)), int(rng.integers(4, 40))]
        n = [int(rng.integers(4, 200)), int(rng.integers(4, 110)), int(rng.integers(3, 48))]
        if rng.random() < 0.8: m[0] = (m[0] + 3) & ~3
        if rng.random() < 0.8: n[0] = (n[0] + 3) & ~3
        nb = int(rng.choice([1, 1, 2]))
        xs = [positions(int(rng.integers(4)), m[d], n[d], rng) for d in range(3)]
        A = [li.TabulatedInterpolator(T, d + 1, k[d], xs[d], m[d]) for d in range(3)]
        tabs = [a.table for a in A]
        src = li.to_device(np.asfortranarray(rng.standard_normal((*m, nb)).astype(T)))
        y = li.to_device(np.asfortranarray(rng.standard_normal((*n, nb)).astype(T)))
        alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.choice([0.0, 1.0, -0.5])))
        res = {}
        desc = (f"T={np.dtype(T).name} k={[type(q).__name__ + '/' + str(q.S) for q in k]} m={m} n={n} nb={nb} alpha={alpha} beta={beta}")
        if os.environ.get("FUZZ_VERBOSE"): print(case, desc, flush=True)
        for mode in ("fused", "passes"):
            li.set_option("exp", 2 if mode == "passes" else 0)
            out = li.to_device(np.asfortranarray(np.ones((*n, nb), dtype=T)))
            back = li.to_device(np.asfortranarray(np.ones((*m, nb), dtype=T)))
            c0 = li.launch_count()
            try:
                if mode == "fused":   # (a first call builds the lazy table metadata: count the launches of a second one)
                    li.sep_apply_tables(False, tabs, alpha, src, 0, li.to_device(np.asfortranarray(np.ones((*n, nb), dtype=T))), nbatch_dims=1)
                    li.sep_apply_tables(True, tabs, alpha, y, 0, li.to_device(np.asfortranarray(np.ones((*m, nb), dtype=T))), nbatch_dims=1)
                    torch.cuda.synchronize()
                    c0 = li.launch_count()
                li.sep_apply_tables(False, tabs, alpha, src, beta, out, nbatch_dims=1)
                torch.cuda.synchronize()
                c1 = li.launch_count()
                li.sep_apply_tables(True, tabs, alpha, y, beta, back, nbatch_dims=1)
                torch.cuda.synchronize()
                c2 = li.launch_count()
            except Exception as exc:
                print(f"EXCEPTION case {case} mode {mode} after {'forward' if 'c1' in dir() and False else '?'}: {exc}\n   {desc}", flush=True)
                return 2
            if mode == "fused" and case >= 0:
                fused_f += (c1 - c0 == 1)
                fused_a += (c2 - c1 == 1)
            res[mode] = (li.to_host(out).astype(np.float64), li.to_host(back).astype(np.float64))
        li.set_option("exp", 0)
        tol = (3e-5 if T == np.float32 else 1e-11)
        for idx, name in ((0, "forward"), (1, "adjoint")):
            a, b = res["fused"][idx], res["passes"][idx]
            scale = max(np.max(np.abs(b)), 1e-30)
            err = np.max(np.abs(a - b)) / scale
            worst = max(worst, err / tol)
            if not np.isfinite(err) or err > tol:
                print(f"MISMATCH case {case} {name}: err {err:.3e} T={np.dtype(T).name} k={[type(q).__name__ + '/' + str(q.S) for q in k]} m={m} n={n} nb={nb} "
                      f"alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}; single-launch forward in {fused_f}, adjoint in {fused_a} cases (first calls build table metadata)")
    return 0


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 1))

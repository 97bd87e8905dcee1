"""Randomised check of TabulatedInterpolator applies along one dimension (all kernels, any dimension, sorted and
unsorted positions, both element types, alpha / beta, device arrays) against the CPU oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import li_b200 as li
import li_oracle as O
from util import KERNELS, CONDITIONS, oracle_kernel, product_kernel

def main(ncases, seed):
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
        tol = 4e-6 if T == np.float32 else 1e-11
        for got, want, name in ((li.to_host(dd), ref, "direct"), (li.to_host(dc), refa, "adjoint")):
            scale = float(np.max(np.abs(want))) + 1.0
            err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) / scale
            worst = max(worst, err / tol)
            if not np.isfinite(err) or err > tol:
                print(f"MISMATCH case {case} {name}: err {err:.3e} T={np.dtype(T).name} {entry[0]} {cname} shape={shape} d={d} nrows={nrows} style={style} alpha={alpha} beta={beta}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0

if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1))

"""Random affine transforms: tiled forward and atomic adjoint against the CPU oracle."""
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
            # 1e-6 / 1e-12 relative times the factors of tests/test_gpu_parity.py::test_affine (8 direct, 64 adjoint:
            # border elements of the adjoint accumulate thousands of clamped contributions, in a different order)
            tol = (1e-6 if T == np.float32 else 1e-12) * (8 if name == "direct" else 64)
            scale = float(np.max(np.abs(want))) + 1.0
            err = float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) / scale
            worst = max(worst, err / tol)
            if not np.isfinite(err) or err > tol:
                print(f"MISMATCH case {case} {name}: err {err:.3e} T={np.dtype(T).name} {e1[0]}/{c1n} {e2[0]}/{c2n} rows={rows} cols={cols} alpha={alpha} beta={beta} R={Rc}", flush=True)
                return 1
    print(f"{ncases} cases ok, worst err / tol = {worst:.3f}")
    return 0

if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2]) if len(sys.argv) > 2 else 1))

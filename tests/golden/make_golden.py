"""Generate tests/golden/golden_v1.npz -- frozen input/output vectors of the hot path.

The reference is Julia and cannot run in this image (no julia binary), so these vectors are produced by the CPU
restatement in oracle/ (which is itself pinned to the reference's own known-answer tests by
tests/test_oracle_reference_suite.py).  Freezing them guards against silent drift of the oracle *and* gives the
GPU suite a checker that does not need the oracle's shared object at run time.  Regenerate with

    python tests/golden/make_golden.py

and commit the result only when a change of behaviour is intended.  Status: J (indices) is pinned to the reference
for every case its tests cover; W of the InterpolationKernels families is "parity unpinned" (see oracle/li_oracle.h).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import li_oracle as O  # noqa: E402
from util import CONDITIONS, DTYPES, KERNELS, hostile_positions, oracle_kernel  # noqa: E402

NCOLS = 23
NROWS = 64


def main(path=os.path.join(HERE, "golden_v1.npz")):
    out = {}
    for ki, entry in enumerate(KERNELS):
        for cname, cond in CONDITIONS:
            for T in DTYPES:
                tag = f"{ki:02d}_{cname}_{np.dtype(T).name}"
                rng = np.random.default_rng(1000 + ki)
                x = hostile_positions(NCOLS, T, rng, NROWS)
                k = oracle_kernel(entry, cond)
                J, W = O.maketable(k, x, NCOLS, T)
                src = rng.standard_normal((NCOLS, 3)).astype(T)
                y = rng.standard_normal((NROWS, 3)).astype(T)
                out[f"x_{tag}"] = x
                out[f"J_{tag}"] = J.astype(np.int32)
                out[f"W_{tag}"] = W
                out[f"src_{tag}"] = src
                out[f"y_{tag}"] = y
                out[f"dir_{tag}"] = O.tab_direct(J, W, NCOLS, 1, np.asfortranarray(src))
                out[f"adj_{tag}"] = O.tab_adjoint(J, W, NCOLS, 1, np.asfortranarray(y))
    # 2-D separable (the headline path at toy size) and the affine operator, one case per dtype
    for T in DTYPES:
        rng = np.random.default_rng(77)
        tn = np.dtype(T).name
        k = O.kernel(O.LANCZOS, 4)
        x1 = np.linspace(1, 40, 80).astype(T)
        x2 = np.linspace(1, 33, 66).astype(T)
        src = np.asfortranarray(rng.standard_normal((40, 33)).astype(T))
        y = np.asfortranarray(rng.standard_normal((80, 66)).astype(T))
        out[f"sep_src_{tn}"] = src
        out[f"sep_y_{tn}"] = y
        out[f"sep_dir_{tn}"] = O.sep2d(False, k, x1, k, x2, src, (80, 66), T)
        out[f"sep_adj_{tn}"] = O.sep2d(True, k, x1, k, x2, y, (40, 33), T)
        kc = O.kernel(O.CATMULLROM, 4)
        th = 0.3
        R = np.array([np.cos(th), -np.sin(th), 3.25, np.sin(th), np.cos(th), -1.5])
        asrc = np.asfortranarray(rng.standard_normal((31, 29)).astype(T))
        ay = np.asfortranarray(rng.standard_normal((37, 41)).astype(T))
        out[f"aff_R_{tn}"] = R
        out[f"aff_src_{tn}"] = asrc
        out[f"aff_y_{tn}"] = ay
        out[f"aff_dir_{tn}"] = O.affine2d(False, kc, kc, R, (37, 41), (31, 29), asrc, T)
        out[f"aff_adj_{tn}"] = O.affine2d(True, kc, kc, R, (37, 41), (31, 29), ay, T)
    np.savez_compressed(path, **out)
    return path


if __name__ == "__main__":
    p = main()
    print(p, os.path.getsize(p), "bytes")

"""Write the INPUTS of tests/golden/golden_v1.npz as raw little-endian files that Julia reads without any package:

    python baseline/julia/export_golden_inputs.py            # -> baseline/_ref/golden_inputs/
    julia --project=<env with LinearInterpolators> baseline/julia/dump_goldens.jl

`dump_goldens.jl` then runs the REAL reference (emmt/LinearInterpolators.jl + InterpolationKernels) on them and writes
J, W and the applied results to baseline/_ref/goldens/, which tests/test_golden.py picks up when present: that turns
"parity unpinned" (the weight formulas are recalled, Julia is not in the build image) into a
bit-level check of the oracle AND of the CUDA tables against Julia.  Nothing here is needed by the product."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from util import CONDITIONS, DTYPES, KERNELS  # noqa: E402

# Julia constructor of every entry of tests/util.py:KERNELS (InterpolationKernels 0.1.x)
JULIA_CTOR = ["RectangularSpline(T)", "LinearSpline(T)", "QuadraticSpline(T)", "CubicSpline(T)", "CatmullRomSpline(T)",
              "CardinalCubicSpline(T, 0.3)", "KeysSpline(T, -0.4)", "MitchellNetravaliSpline(T, 1/3, 1/3)",
              "LanczosKernel(T, 4)", "LanczosKernel(T, 6)", "LanczosKernel(T, 8)"]


def main():
    out = os.path.join(ROOT, "baseline", "_ref", "golden_inputs")
    os.makedirs(out, exist_ok=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    lines = []
    for ki in range(len(KERNELS)):
        for cname, _ in CONDITIONS:
            for T in DTYPES:
                tn = np.dtype(T).name
                tag = f"{ki:02d}_{cname}_{tn}"
                for part in ("x", "src", "y"):
                    np.asfortranarray(gold[f"{part}_{tag}"]).T.astype(T).tofile(os.path.join(out, f"{tag}.{part}.bin"))
                nrows, ncols = gold[f"x_{tag}"].size, gold[f"src_{tag}"].shape[0]
                lines.append(f"{tag}\t{JULIA_CTOR[ki]}\t{cname}\t{'Float32' if T == np.float32 else 'Float64'}\t{ncols}\t{nrows}\t{gold[f'src_{tag}'].shape[1]}")
    with open(os.path.join(out, "cases.tsv"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print(out, len(lines), "cases")


if __name__ == "__main__":
    main()

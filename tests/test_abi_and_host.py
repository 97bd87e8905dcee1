"""CPU-only checks: the C-ABI library loads and exports every symbol include/li_b200.h declares, the host
mirror parses the reference's constructor forms, and the product fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "li_b200.h")).read()
    return sorted(set(re.findall(r"^LI_API [a-z0-9_ ]+\*? ?\*?(li_[a-z0-9_]+)\(", text, flags=re.M)))


def test_header_symbols_are_exported():
    import li_b200
    names = _declared_symbols()
    assert len(names) >= 30
    lib = ctypes.CDLL(li_b200.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in li_b200.h but not exported"
    assert set(names) == set(li_b200._lib.EXPORTED), "python binding and header disagree"
    assert li_b200.version() == 200


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "linearinterpolators.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "li_oracle" not in text and "oracle/" not in text, f"{f} references the oracle"


def test_kernel_constructors():
    import li_b200 as li
    k = li.LinearSpline(np.float32)
    assert k.T == np.float32 and k.S == 2 and k.B is li.Flat
    k = li.SafeFlat(li.LanczosKernel(6))
    assert k.T == np.float64 and k.S == 6 and k.B is li.SafeFlat and k.desc().boundary == 1
    k = li.KeysSpline(-0.4)
    assert k.params == (-0.4,) and k.S == 4
    k = li.MitchellNetravaliSpline(np.float32, li.SafeFlat)
    assert k.params == (1 / 3, 1 / 3) and k.with_eltype(np.float64).T == np.float64
    with pytest.raises(li.ArgumentError):
        li.LanczosKernel(5)


def test_affine_transform_algebra():
    import li_b200 as li
    c = (5.6, 6.4)
    R = c + li.rotate(li.AffineTransform2D() - c, 0.2)
    x, y = R(1.0, 2.0)
    np.testing.assert_allclose((x, y), (np.cos(0.2) * 1 - np.sin(0.2) * 2, np.sin(0.2) * 1 + np.cos(0.2) * 2))
    A = li.AffineTransform2D(1.5, 0.2, 3, -0.1, 0.9, -2)
    ix, iy = li.intercept(A)
    np.testing.assert_allclose(A(ix, iy), (0, 0), atol=1e-12)
    B = A * li.AffineTransform2D(0, 1, 0, 1, 0, 0)
    np.testing.assert_allclose(B(2.0, 5.0), A(5.0, 2.0))


def test_fractional_index_matches_reference_formula():
    import li_b200 as li
    g = li.Range(0.5, 10.5, length=21)
    pos = np.array([0.5, 3.25, 10.5])
    f = li.fractional_index(np.float64, pos, g)
    np.testing.assert_allclose(f, 1 + (pos - 0.5) / 0.5)
    assert li.fractional_index(np.float32, pos).dtype == np.float32


def test_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import li_b200 as li
    with pytest.raises(li.CudaError):
        li.TabulatedInterpolator(li.LinearSpline(), np.linspace(1, 5, 9), 5)

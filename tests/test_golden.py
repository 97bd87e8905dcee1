"""Frozen vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).

CPU leg: the oracle still reproduces them bit for bit (guards the checker against drift).
GPU leg: the CUDA path reproduces them -- tables bit for bit, values to 1e-6 / 1e-12 -- without the oracle."""
import os

import numpy as np
import pytest

from util import CONDITIONS, DTYPES, KERNEL_IDS, KERNELS, assert_close, oracle_kernel, product_kernel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz")
NCOLS = 23


@pytest.fixture(scope="module")
def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


def _bits_equal(a, b):
    U = np.uint32 if a.dtype == np.float32 else np.uint64
    nan = np.isnan(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), nan) and np.array_equal(a.view(U)[~nan], b.view(U)[~nan])


CASES = [(ki, cname, cond, T) for ki in range(len(KERNELS)) for cname, cond in CONDITIONS for T in DTYPES]
CASE_IDS = [f"{KERNEL_IDS[ki]}-{cname}-{np.dtype(T).name}" for ki, cname, cond, T in CASES]


@pytest.mark.parametrize("ki,cname,cond,T", CASES, ids=CASE_IDS)
def test_oracle_reproduces_golden(gold, ki, cname, cond, T):
    import li_oracle as O
    tag = f"{ki:02d}_{cname}_{np.dtype(T).name}"
    J, W = O.maketable(oracle_kernel(KERNELS[ki], cond), gold[f"x_{tag}"], NCOLS, T)
    assert np.array_equal(J, gold[f"J_{tag}"])
    assert _bits_equal(W, gold[f"W_{tag}"])
    d = O.tab_direct(J, W, NCOLS, 1, np.asfortranarray(gold[f"src_{tag}"]))
    a = O.tab_adjoint(J, W, NCOLS, 1, np.asfortranarray(gold[f"y_{tag}"]))
    assert _bits_equal(np.asarray(d), gold[f"dir_{tag}"])
    assert _bits_equal(np.asarray(a), gold[f"adj_{tag}"])


def test_golden_tables_are_wellformed(gold):
    """Properties every reference table has (src/tabulated.jl:206-215): 1 <= J <= ncols; interior rows of a
    normalised kernel sum to 1."""
    for key, J in gold.items():
        if key.startswith("J_"):
            assert J.min() >= 1 and J.max() <= NCOLS, key
    for T in DTYPES:
        W = gold[f"W_04_Flat_{np.dtype(T).name}"]      # Catmull-Rom
        assert np.allclose(W.sum(axis=0), 1, atol=8 * np.finfo(T).eps)


@pytest.mark.gpu
@pytest.mark.parametrize("ki,cname,cond,T", CASES, ids=CASE_IDS)
def test_cuda_reproduces_golden_tables_and_applies(gold, ki, cname, cond, T):
    import li_b200 as li
    tag = f"{ki:02d}_{cname}_{np.dtype(T).name}"
    A = li.TabulatedInterpolator(T, product_kernel(li, KERNELS[ki], cond, T), gold[f"x_{tag}"], NCOLS)
    assert np.array_equal(A.J, gold[f"J_{tag}"])
    assert _bits_equal(np.ascontiguousarray(A.W), np.ascontiguousarray(gold[f"W_{tag}"]))
    src = np.asfortranarray(gold[f"src_{tag}"])
    y = np.asfortranarray(gold[f"y_{tag}"])
    ok = ~np.isnan(gold[f"dir_{tag}"]).any()
    d = li.apply(li.Direct, A, 1, src)
    a = li.apply(li.Adjoint, A, 1, y)
    if ok:
        assert_close(d, gold[f"dir_{tag}"], T, factor=4)
        assert_close(a, gold[f"adj_{tag}"], T, factor=16)
    else:
        assert np.array_equal(np.isnan(d), np.isnan(gold[f"dir_{tag}"]))


@pytest.mark.gpu
@pytest.mark.parametrize("T", DTYPES)
def test_cuda_reproduces_golden_sep2d_and_affine(gold, T):
    import li_b200 as li
    tn = np.dtype(T).name
    ker = li.LanczosKernel(T, 4)
    x1 = np.linspace(1, 40, 80).astype(T)
    x2 = np.linspace(1, 33, 66).astype(T)
    dst = np.zeros((80, 66), dtype=T, order="F")
    li.apply_b(1, li.Direct, ker, x1, ker, x2, np.asfortranarray(gold[f"sep_src_{tn}"]), 0, dst)
    assert_close(dst, gold[f"sep_dir_{tn}"], T, factor=4)
    back = np.zeros((40, 33), dtype=T, order="F")
    li.apply_b(1, li.Adjoint, ker, x1, ker, x2, np.asfortranarray(gold[f"sep_y_{tn}"]), 0, back)
    assert_close(back, gold[f"sep_adj_{tn}"], T, factor=16)
    cr = li.CatmullRomSpline(T)
    R = li.AffineTransform2D(*gold[f"aff_R_{tn}"], T=T)
    out = np.zeros((37, 41), dtype=T, order="F")
    li.apply_b(1, li.Direct, cr, cr, R, np.asfortranarray(gold[f"aff_src_{tn}"]), 0, out)
    assert_close(out, gold[f"aff_dir_{tn}"], T, factor=4)
    outa = np.zeros((31, 29), dtype=T, order="F")
    li.apply_b(1, li.Adjoint, cr, cr, R, np.asfortranarray(gold[f"aff_y_{tn}"]), 0, outa)
    assert_close(outa, gold[f"aff_adj_{tn}"], T, factor=16)


# ------------------------------------------------------------------------------------------------
# Vectors of the REAL reference, when somebody has produced them (baseline/julia/dump_goldens.jl; needs Julia, which the
# build image does not have): baseline/_ref/goldens/<tag>.{J,W,dir,adj}.bin.  With them the oracle -- and through it the
# CUDA tables -- are pinned to Julia bit for bit instead of "parity unpinned" (SURVEY.md section 8c).
JULIA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "goldens")
have_julia = os.path.exists(os.path.join(JULIA, "cases.tsv"))


def _julia_case(tag, T, S, nrows, ncols, nb):
    rd = lambda part, dt, shape: np.fromfile(os.path.join(JULIA, f"{tag}.{part}.bin"), dtype=dt).reshape(shape, order="F")
    return rd("J", np.int64, (S, nrows)), rd("W", T, (S, nrows)), rd("dir", T, (nrows, nb)), rd("adj", T, (ncols, nb))


@pytest.mark.skipif(not have_julia, reason="no Julia goldens (baseline/julia/dump_goldens.jl has not been run)")
@pytest.mark.parametrize("ki,cname,cond,T", CASES, ids=CASE_IDS)
def test_oracle_matches_julia_goldens(gold, ki, cname, cond, T):
    import li_oracle as O
    tag = f"{ki:02d}_{cname}_{np.dtype(T).name}"
    if os.path.exists(os.path.join(JULIA, f"{tag}.error.txt")):
        pytest.skip("the reference throws on these positions")
    S = KERNELS[ki][2]
    Jj, Wj, dj, aj = _julia_case(tag, T, S, gold[f"x_{tag}"].size, NCOLS, gold[f"src_{tag}"].shape[1])
    J, W = O.maketable(oracle_kernel(KERNELS[ki], cond), gold[f"x_{tag}"], NCOLS, T)
    assert np.array_equal(J, Jj), "index table differs from Julia"
    assert _bits_equal(W, Wj), "weights differ from Julia (InterpolationKernels formula or operation order)"
    assert_close(O.tab_direct(J, W, NCOLS, 1, np.asfortranarray(gold[f"src_{tag}"])), dj, T, factor=4)
    assert_close(O.tab_adjoint(J, W, NCOLS, 1, np.asfortranarray(gold[f"y_{tag}"])), aj, T, factor=16)


@pytest.mark.gpu
@pytest.mark.skipif(not have_julia, reason="no Julia goldens (baseline/julia/dump_goldens.jl has not been run)")
@pytest.mark.parametrize("ki,cname,cond,T", CASES, ids=CASE_IDS)
def test_cuda_matches_julia_goldens(gold, ki, cname, cond, T):
    import li_b200 as li
    tag = f"{ki:02d}_{cname}_{np.dtype(T).name}"
    if os.path.exists(os.path.join(JULIA, f"{tag}.error.txt")):
        pytest.skip("the reference throws on these positions")
    S = KERNELS[ki][2]
    Jj, Wj, dj, aj = _julia_case(tag, T, S, gold[f"x_{tag}"].size, NCOLS, gold[f"src_{tag}"].shape[1])
    A = li.TabulatedInterpolator(T, 1, product_kernel(li, KERNELS[ki], cond, T), gold[f"x_{tag}"], NCOLS)
    assert np.array_equal(A.J, Jj) and _bits_equal(np.asarray(A.W), Wj)
    assert_close(A * np.asfortranarray(gold[f"src_{tag}"]), dj, T, factor=4)
    assert_close(A.adj * np.asfortranarray(gold[f"y_{tag}"]), aj, T, factor=16)

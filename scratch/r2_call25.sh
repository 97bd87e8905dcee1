#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2ae_pytest.log
python - <<'PY'
import numpy as np, torch, time
import li_b200 as li
# Flat clamping: a quarter of 2e6 unsorted rows hangs over the edges -> two columns of ~1.5e6 entries each
T=np.float32; rng=np.random.default_rng(1); n=2_000_000; m=10_000
xs=rng.uniform(-m/3, m+m/3, n)
A=li.SparseInterpolator(li.CatmullRomSpline(T), xs, m)
y=li.to_device(rng.standard_normal(n).astype(T))
out=A.adj*y; torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): out=A.adj*y
torch.cuda.synchronize(); print("gather adjoint, 2e6 unsorted rows with heavy edge columns: %.3f ms per apply" % ((time.perf_counter()-t0)/5*1e3))
PY

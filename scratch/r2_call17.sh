#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "aligned or sep3d_on or affine or on_the_fly" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r2w_pytest.log
python - <<'PY'
import numpy as np, torch, time
import li_b200 as li
T=np.float32; n=8192
ker=li.CatmullRomSpline(T)
R=li.AffineTransform2D(0.98,0.0,3.3,0.0,1.01,-2.2,T=T)
A=li.TwoDimensionalTransformInterpolator((n,n),(n,n),ker,ker,R)
x=li.device_empty((n,n),T); x.normal_(); y=li.device_empty((n,n),T)
def t(fn,reps=5):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
for e in (0,4):
    li.set_option("exp",e)
    print("exp",e,"fwd ms",t(lambda: li.apply_b(1,li.Direct,A,x,False,0,y)),"adj ms",t(lambda: li.apply_b(1,li.Adjoint,A,y,False,0,x)))
PY

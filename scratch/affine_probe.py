"""Config #3 in small (8192^2 x NB images): times the affine direct / adjoint kernels."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
NB = int(os.environ.get("NB", "4")); n = int(os.environ.get("N", "8192")); T = np.float32
dev = torch.device("cuda", 0)
ker = li.CatmullRomSpline(T)
c = ((n + 1) / 2, (n + 1) / 2)
R = c + li.rotate(li.AffineTransform2D(T=T) - c, float(os.environ.get("THETA", "0.2")))
src = li.device_empty((n, n, NB), T, dev); src.normal_()
dst = li.device_empty((n, n, NB), T, dev)
for name, P in (("forward", li.Direct), ("adjoint", li.Adjoint)):
    fn = lambda: li.affine_apply(1, P, ker, ker, R, src, 0, dst, nbatch_dims=1)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/3/NB*1e3:.0f} us per image", flush=True)

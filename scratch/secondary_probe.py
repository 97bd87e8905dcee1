"""One launch each of the secondary kernels (for ncu): strided-dimension streaming kernels (1024^3 along dim 3), tiled
affine forward and atomic affine adjoint (8192^2 x 4)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
T = np.float32
n = 1024
A3 = li.TabulatedInterpolator(T, 3, li.LinearSpline(T), 1 + np.arange(n) * 0.98 + 0.73, n)
x = li.device_empty((n, n, n), T); x.normal_()
y = li.device_empty((n, n, n), T)
for _ in range(2):
    li.apply_b(1, li.Direct, A3, x, False, 0, y)
    li.apply_b(1, li.Adjoint, A3, x, False, 0, y)
del x, y
torch.cuda.empty_cache()
n = 8192; nb = 4
ker = li.CatmullRomSpline(T)
c = ((n + 1) / 2, (n + 1) / 2)
R = c + li.rotate(li.AffineTransform2D(T=T) - c, 0.2)
src = li.device_empty((n, n, nb), T); src.normal_()
dst = li.device_empty((n, n, nb), T)
for _ in range(2):
    li.affine_apply(1, li.Direct, ker, ker, R, src, 0, dst, nbatch_dims=1)
    li.affine_apply(1, li.Adjoint, ker, ker, R, src, 0, dst, nbatch_dims=1)
torch.cuda.synchronize()
print("done")

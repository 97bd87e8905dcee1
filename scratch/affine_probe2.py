import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32
n = 8192
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = (4096.5, 4096.5)
R = c + li.rotate(li.AffineTransform2D(T=T) - c, 0.2)
ker = li.CatmullRomSpline(T)
src = li.device_empty((n, n, nb), T); src.normal_()
dst = li.device_empty((n, n, nb), T)
for _ in range(3):
    li.affine_apply(1, li.Adjoint, ker, ker, R, src, 0, dst, nbatch_dims=1)
torch.cuda.synchronize()

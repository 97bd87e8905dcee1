"""Per-CTA start / end times of the headline streaming adjoint (probe build: LI_NVCC_EXTRA=-DADS_PROBE, adj_debug = 5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32; M, N = 4096, 8192
x = np.linspace(1.0, float(M), N)
ker = li.LanczosKernel(T, 4)
H = li.TabulatedInterpolator(T, 1, ker, x, M) * li.TabulatedInterpolator(T, 2, ker, x, M)
ys = [li.device_empty((N, N), T) for _ in range(2)]
for y in ys: y.normal_()
z = li.device_empty((M, M), T)
for k in range(4): li.apply_b(1, li.Adjoint, H, ys[k % 2], False, 0, z)
torch.cuda.synchronize()
li.set_option("adj_debug", 5)
li.apply_b(1, li.Adjoint, H, ys[0], False, 0, z)
torch.cuda.synchronize()
li.set_option("adj_debug", 0)

"""Per-CTA cycle counts of the multi-frame accumulate kernel (config #5 shape, fewer frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32
m = 2048
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ker = li.SafeFlat(li.LanczosKernel(T, 6))
xs = np.arange(1, m + 1, dtype=np.float64)
shifts = np.random.default_rng(5).uniform(-8, 8, (nf, 2))
tabs = [(li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table, li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table) for k in range(nf)]
dev = torch.device("cuda:0")
frames = li.device_empty((m, m, nf), T, dev)
frames.normal_()
views = [frames[:, :, k] for k in range(nf)]
model = li.device_empty((m, m), T, dev)
for _ in range(3):
    li.adjoint_accumulate(tabs, views, model)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); li.adjoint_accumulate(tabs, views, model); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1), "us/frame", 1e3 * e0.elapsed_time(e1) / nf)
li.set_option("adj_debug", 9)
li.adjoint_accumulate(tabs, views, model)
torch.cuda.synchronize()

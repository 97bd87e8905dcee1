"""Multi-frame accumulate kernel (config #5 shape): time per frame for forced numbers of frame groups."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32
m = 2048
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ker = li.SafeFlat(li.LanczosKernel(T, 6))
xs = np.arange(1, m + 1, dtype=np.float64)
dev = torch.device("cuda:0")
frames = li.device_empty((m, m, nf), T, dev)
frames.normal_()
views = [frames[:, :, k] for k in range(nf)]
model = li.device_empty((m, m), T, dev)
ref = None
for g in [int(v) for v in sys.argv[2:]] or [0]:
    shifts = np.random.default_rng(5).uniform(-8, 8, (nf, 2))
    tabs = [(li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table, li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table) for k in range(nf)]
    li.set_option("adj_debug", 10 + g if g else 0)
    batch = li.FrameBatch(tabs, views)
    for _ in range(3):
        li.adjoint_accumulate(batch, model)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); li.adjoint_accumulate(batch, model); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out = model.clone()
    if ref is None:
        ref = out
    print(f"G {g}: ms {min(ts):.3f} us/frame {1e3 * min(ts) / nf:.2f}  max rel diff vs first {float((out - ref).abs().max() / ref.abs().max()):.2e}", flush=True)

"""Config #5 local kernel at few frames per rank: time of li_adjoint_accumulate for forced numbers of frame groups G."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
T = np.float32; m = 2048
ker = li.SafeFlat(li.LanczosKernel(T, 6))
xs = np.arange(1, m + 1, dtype=np.float64)
for nf in (64, 128, 512):
    shifts = np.random.default_rng(5).uniform(-8, 8, (nf, 2))
    tabs = [(li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table, li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table) for k in range(nf)]
    frames = li.device_empty((m, m, nf), T); frames.normal_()
    model = li.device_empty((m, m), T)
    batch = li.FrameBatch(tabs, [frames[:, :, k] for k in range(nf)])
    res = {}
    for g in (0, 1, 2, 3, 4, 6, 8, 12, 16):
        li.set_option("adj_debug", 10 + g if g else 0)
        # a forced G changes the plan: drop cached plans by rebuilding the batch key? plans are keyed by tables+rows: G is read at build time
        # -> use fresh tables per G
        tb = [(li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table, li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table) for k in range(nf)]
        b = li.FrameBatch(tb, [frames[:, :, k] for k in range(nf)])
        for _ in range(2): li.adjoint_accumulate(b, model)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); li.adjoint_accumulate(b, model); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[g] = min(ts)
    li.set_option("adj_debug", 0)
    print(nf, "frames:", {g: round(v, 3) for g, v in res.items()}, "us/frame best", round(1e3 * min(res.values()) / nf, 2), flush=True)

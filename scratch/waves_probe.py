"""Forward kernel: waves of CTAs (exp option, bits 8..11) across kernels / rates; CUDA events around back-to-back launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32
cases = [("Lanczos4 2x", li.LanczosKernel(T, 4), 4096, 8192), ("Lanczos6 2x", li.LanczosKernel(T, 6), 4096, 8192),
         ("CatmullRom 2x", li.CatmullRomSpline(T), 4096, 8192), ("Linear 2x", li.LinearSpline(T), 4096, 8192),
         ("Lanczos4 1x", li.LanczosKernel(T, 4), 8192, 8192), ("Lanczos4 1.5x", li.LanczosKernel(T, 4), 5464, 8192),
         ("Lanczos4 0.5x", li.LanczosKernel(T, 4), 8192, 4096), ("Lanczos4 3x", li.LanczosKernel(T, 4), 2730, 8190),
         ("Lanczos4 2x f64", li.LanczosKernel(np.float64, 4), 4096, 8192), ("Lanczos4 2x 2048->4096", li.LanczosKernel(T, 4), 2048, 4096)]
for name, ker, m, n in cases:
    TT = ker.T
    x = np.linspace(1.0, float(m), n)
    H = li.TabulatedInterpolator(TT, 1, ker, x, m) * li.TabulatedInterpolator(TT, 2, ker, x, m)
    srcs = [li.device_empty((m, m), TT) for _ in range(3)]
    dsts = [li.device_empty((n, n), TT) for _ in range(3)]
    for s in srcs: s.normal_()
    res = {}
    for w in (2, 3, 4, 5, 6):
        li.set_option("exp", w << 8)
        for k in range(3): li.apply_b(1, li.Direct, H, srcs[k], False, 0, dsts[k])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(30): li.apply_b(1, li.Direct, H, srcs[k % 3], False, 0, dsts[k % 3])
        e1.record(); torch.cuda.synchronize()
        res[w] = e0.elapsed_time(e1) / 30 * 1e3
    li.set_option("exp", 0)
    print(f"{name:24s}", {w: round(v, 1) for w, v in res.items()}, flush=True)

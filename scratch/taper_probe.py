"""Headline adjoint: chunk taper strength (exp bits 12..15 = strength + 1), back-to-back launches, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
T = np.float32
for name, ker, M, N in (("Lanczos4 2x", li.LanczosKernel(T, 4), 4096, 8192), ("CatmullRom 2x", li.CatmullRomSpline(T), 4096, 8192),
                        ("Lanczos4 1x", li.LanczosKernel(T, 4), 8192, 8192), ("Lanczos6 2x", li.LanczosKernel(T, 6), 4096, 8192),
                        ("Lanczos4 1.5x", li.LanczosKernel(T, 4), 5464, 8192)):
    x = np.linspace(1.0, float(M), N)
    ys = [li.device_empty((N, N), T) for _ in range(3)]
    for y in ys: y.normal_()
    zs = [li.device_empty((M, M), T) for _ in range(3)]
    res = {}
    for s in (0, 4, 6, 8, 10, 12, 14):
        li.set_option("exp", (s + 1) << 12)
        H = li.TabulatedInterpolator(T, 1, ker, x, M) * li.TabulatedInterpolator(T, 2, ker, x, M)
        for k in range(3): li.apply_b(1, li.Adjoint, H, ys[k], False, 0, zs[k])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                for k in range(3): li.apply_b(1, li.Adjoint, H, ys[k], False, 0, zs[k])
        torch.cuda.synchronize()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        res[s] = e0.elapsed_time(e1) / 30 * 1e3
    li.set_option("exp", 0)
    print(f"{name:16s}", {k: round(v, 1) for k, v in res.items()}, flush=True)

"""Fused 2-D separable forward / adjoint across kernels, rates and element types (device resident, CUDA events)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = [("Lanczos4 2x f32", np.float32, lambda T: li.LanczosKernel(T, 4), 4096, 8192),
         ("Lanczos4 2x f64", np.float64, lambda T: li.LanczosKernel(T, 4), 4096, 8192),
         ("Lanczos6 2x f32", np.float32, lambda T: li.LanczosKernel(T, 6), 4096, 8192),
         ("CatmullRom 2x f32", np.float32, lambda T: li.CatmullRomSpline(T), 4096, 8192),
         ("Linear 2x f32", np.float32, lambda T: li.LinearSpline(T), 4096, 8192),
         ("Lanczos4 1x f32", np.float32, lambda T: li.LanczosKernel(T, 4), 8192, 8192),
         ("Lanczos4 1.5x f32", np.float32, lambda T: li.LanczosKernel(T, 4), 5464, 8192),
         ("Lanczos4 0.5x f32 (down)", np.float32, lambda T: li.LanczosKernel(T, 4), 8192, 4096),
         ("Quadratic 2x f32 (S=3: passes)", np.float32, lambda T: li.QuadraticSpline(T), 4096, 8192)]
only = os.environ.get("ONLY")
for name, T, mk, m, n in cases:
    if only and only not in name: continue
    ker = mk(T)
    x = np.linspace(1.0, float(m), n)
    A1, A2 = li.TabulatedInterpolator(T, 1, ker, x, m), li.TabulatedInterpolator(T, 2, ker, x, m)
    H = A1 * A2
    src = li.device_empty((m, m), T); src.normal_()
    y = li.device_empty((n, n), T); y.normal_()
    out = li.device_empty((n, n), T); back = li.device_empty((m, m), T)
    tf = timeit(lambda: li.apply_b(1, li.Direct, H, src, False, 0, out))
    ta = timeit(lambda: li.apply_b(1, li.Adjoint, H, y, False, 0, back))
    byt = (m * m + n * n) * np.dtype(T).itemsize
    print(f"{name:34s} fwd {tf*1e3:7.1f} us ({byt/tf/1e6:6.0f} GB/s)  adj {ta*1e3:7.1f} us ({byt/ta/1e6:6.0f} GB/s)", flush=True)

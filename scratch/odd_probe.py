import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
T = np.float32
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1023, 1022, 1024):
    A3 = li.TabulatedInterpolator(T, 3, li.CubicSpline(T), 1 + np.arange(n) * 0.98 + 0.3, n)
    x = li.device_empty((n, n, n), T); x.normal_()
    y = li.device_empty((n, n, n), T)
    tf = timeit(lambda: li.apply_b(1, li.Direct, A3, x, False, 0, y))
    ta = timeit(lambda: li.apply_b(1, li.Adjoint, A3, x, False, 0, y))
    print(f"n={n}: dim3 fwd {tf:.3f} ms ({8*n**3/tf/1e6:.0f} GB/s) adj {ta:.3f} ms ({8*n**3/ta/1e6:.0f} GB/s)", flush=True)
    del x, y; torch.cuda.empty_cache()

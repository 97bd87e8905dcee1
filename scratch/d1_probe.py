import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
T = np.float32
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for ncols, nrows, npost in ((32, 64, 1 << 21), (96, 192, 1 << 20), (256, 512, 1 << 18), (1000, 2000, 1 << 16), (4096, 8192, 8192)):
    A = li.TabulatedInterpolator(T, 1, li.LanczosKernel(T, 4), np.linspace(1, ncols, nrows), ncols)
    x = li.device_empty((ncols, npost), T); x.normal_()
    y = li.device_empty((nrows, npost), T); y.normal_()
    o = li.device_empty((nrows, npost), T); b = li.device_empty((ncols, npost), T)
    tf = timeit(lambda: li.apply_b(1, li.Direct, A, x, False, 0, o))
    ta = timeit(lambda: li.apply_b(1, li.Adjoint, A, y, False, 0, b))
    byt = (ncols + nrows) * npost * 4
    print(f"{ncols}->{nrows} x {npost}: fwd {tf:.3f} ms ({byt/tf/1e6:.0f} GB/s)  adj {ta:.3f} ms ({byt/ta/1e6:.0f} GB/s)  [{'gather' if os.environ.get('LI_B200_NO_STREAM_DN') else 'fused'}]", flush=True)

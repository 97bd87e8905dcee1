import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
T = np.float32; n = 1024
kers = (li.CubicSpline(T), li.MitchellNetravaliSpline(T), li.LinearSpline(T))
deltas = (0.37, 0.11, 0.73)
A = [li.TabulatedInterpolator(T, d + 1, kers[d], 1 + np.arange(n) * 0.98 + deltas[d], n) for d in range(3)]
x = li.device_empty((n, n, n), T); x.normal_()
y = li.device_empty((n, n, n), T)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for P, name in ((li.Direct, "fwd"), (li.Adjoint, "adj")):
    for label, op in (("dim3 only", A[2]), ("dim2 only", A[1]), ("dim1 only", A[0]), ("dims 1-2 fused", A[0] * A[1]), ("all", A[0] * A[1] * A[2])):
        t = timeit(lambda: li.apply_b(1, P, op, x, False, 0, y))
        print(f"{name} {label}: {t:.3f} ms  ({2*4*n**3/t/1e6:.0f} GB/s)", flush=True)

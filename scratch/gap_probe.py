"""Per-call device time of the fused forward/adjoint with no intermediate events, vs host time per call."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import li_b200 as li
from bench import make_operator, M, N_OUT
dev = torch.device("cuda", 0)
A1, A2, H = make_operator(li)
NB = 2
xs = [li.device_empty((M, M), np.float32, dev) for _ in range(NB)]
ys = [li.device_empty((N_OUT, N_OUT), np.float32, dev) for _ in range(NB)]
zs = [li.device_empty((M, M), np.float32, dev) for _ in range(NB)]
for t in xs + ys: t.normal_()
def fwd(k): li.apply_b(1, li.Direct, H, xs[k % NB], False, 0, ys[k % NB])
def adj(k): li.apply_b(1, li.Adjoint, H, ys[k % NB], False, 0, zs[k % NB])
def both(k): fwd(k); adj(k)
K = 200
for name, f in (("fwd", fwd), ("adj", adj), ("both", both)):
    for k in range(10): f(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for k in range(K): f(k)
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name}: device {1e3*e0.elapsed_time(e1)/K:.1f} us/call, host issue {1e6*th/K:.1f} us/call  (dbg={os.environ.get('LI_B200_ADJ_DEBUG','0')})", flush=True)

"""e2e leg of bench.py on its own (pinned host buffers, forward and adjoint on two host streams), 10 steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import li_b200 as li
M, N = 4096, 8192
x = np.linspace(1.0, float(M), N)
ker = li.LanczosKernel(np.float32, 4)
H = li.TabulatedInterpolator(np.float32, 1, ker, x, M) * li.TabulatedInterpolator(np.float32, 2, ker, x, M)
xh = li.pinned_empty((M, M), np.float32); yh_in = li.pinned_empty((N, N), np.float32)
xh[...] = 1.0; yh_in[...] = 1.0
yh = li.vcreate(li.Direct, H, xh); zh = li.vcreate(li.Adjoint, H, yh_in)
sa, sb = li.HostStream(), li.HostStream()
def step():
    with li.async_host(sa): li.apply_b(1, li.Direct, H, xh, False, 0, yh)
    with li.async_host(sb): li.apply_b(1, li.Adjoint, H, yh_in, False, 0, zh)
    sa.synchronize(); sb.synchronize()
for nb in (0, 2, 3, 4, 6, 8, 0):
  li.set_option("no_bands", nb)
  for _ in range(3): step()
  ts = []
  for _ in range(10):
    t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
  print(f"bands {nb or 'default'}: e2e step: min {min(ts)*1e3:.2f} ms, median {np.median(ts)*1e3:.2f} ms -> {2*N*N/np.median(ts)/1e9:.2f} G samples/s, {(8*M*M+8*N*N)/np.median(ts)/1e9:.1f} GB/s over PCIe")

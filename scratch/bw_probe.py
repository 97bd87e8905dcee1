"""Ceilings for the headline shapes: a fill of 268 MB, a copy of 168+168 MB, a read of 268 MB (sum), CUDA events."""
import torch
dev = torch.device("cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
n_out, n_in = 8192 * 8192, 4096 * 4096
ys = [torch.empty(n_out, device=dev) for _ in range(3)]
xs = [torch.randn(n_in, device=dev) for _ in range(3)]
k = [0]
def fill():
    k[0] += 1; ys[k[0] % 3].fill_(1.0)
def copy():
    k[0] += 1; ys[k[0] % 3][: n_out // 2].copy_(ys[(k[0] + 1) % 3][n_out // 2:])
def read():
    k[0] += 1; ys[k[0] % 3].sum()
us = t(fill); print(f"fill 268 MB: {us:.1f} us = {n_out*4/us/1e3:.0f} GB/s")
us = t(copy); print(f"copy 134 -> 134 MB: {us:.1f} us = {n_out*4/us/1e3:.0f} GB/s")
us = t(read); print(f"sum 268 MB: {us:.1f} us = {n_out*4/us/1e3:.0f} GB/s")

"""PCIe ceilings of the box for the e2e leg: H2D alone, D2H alone, both at once (pinned host memory, 268 MB each)."""
import time, torch
n = 8192 * 8192
h_in = torch.empty(n, dtype=torch.float32).pin_memory(); h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, device="cuda"); d_out = torch.randn(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
for name, a, b in (("H2D alone", 1, 0), ("D2H alone", 0, 1), ("both at once", 1, 1)):
    run(a, b, 2)
    t = run(a, b)
    print(f"{name}: {t*1e3:.2f} ms per 268 MB each -> {(a + b) * n * 4 / t / 1e9:.1f} GB/s aggregate")

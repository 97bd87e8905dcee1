#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs #1, #3, #4, #5 on ONE GPU (bench.py is the headline, config #2).
Prints one JSON line per config; numbers go to BASELINE.md section 4.  Device-resident buffers, CUDA events,
5 warm-ups, inputs larger than L2 or rotated."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PEAK = 6531.9
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps, warm=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.min(ts)), float(np.median(ts))


def report(name, samples, algo_bytes, tmin, tmed, **extra):
    line = {"config": name, "ms_min": tmin, "ms_median": tmed, "Gsamples_s": samples / (tmed * 1e-3) / 1e9,
            "GB_s": algo_bytes / (tmed * 1e-3) / 1e9, "frac_measured_peak": algo_bytes / (tmed * 1e-3) / 1e9 / PEAK,
            "frac_8TBs": algo_bytes / (tmed * 1e-3) / 1e9 / 8000.0, **extra}
    print(json.dumps(line), flush=True)


def cfg1(li, torch):
    T = np.float64
    ncols, nrows = 10_000, 20_000
    pos = np.linspace(1, ncols, nrows)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for tag, p in (("monotone", pos), ("shuffled", np.random.default_rng(1).permutation(pos))):
        A = li.SparseInterpolator(li.CatmullRomSpline(T), p, ncols)
        x = li.to_device(np.random.default_rng(0).standard_normal(ncols))
        y = li.to_device(np.random.default_rng(0).standard_normal(nrows))
        out, back = li.vcreate(li.Direct, A, x), li.vcreate(li.Adjoint, A, y)
        bytes_ = nrows * 4 * (4 + 8) + ncols * 8 + nrows * 8          # int32 J + f64 C, x, y
        for variant, vname in ((li.VARIANT_GATHER, "gather"), (li.VARIANT_ATOMIC, "atomic")):
            A.variant = variant
            def fwd():
                flush.zero_()
                li.apply_b(1, li.Direct, A, x, False, 0, out)
            def adj():
                flush.zero_()
                li.apply_b(1, li.Adjoint, A, y, False, 0, back)
            def base():
                flush.zero_()
            b0, _ = timeit(base, 20)
            if variant == li.VARIANT_GATHER:
                t = timeit(fwd, 20)
                report(f"#1 sparse 1-D CatmullRom F64 {tag} forward (incl. launch, L2 flushed)", nrows, bytes_, t[0] - b0, t[1] - b0)
            t = timeit(adj, 20)
            report(f"#1 sparse 1-D CatmullRom F64 {tag} adjoint {vname}", nrows, bytes_, t[0] - b0, t[1] - b0)

            # the same calls replayed from a CUDA graph (100 launches back to back: no host in the loop, no L2 flush):
            # config #1 is launch-latency bound, the graph shows the kernel itself
            def graph_us(fn):
                st, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
                fn()
                torch.cuda.synchronize()
                with torch.cuda.stream(st):
                    with torch.cuda.graph(g, stream=st):
                        for _ in range(100):
                            fn()
                torch.cuda.synchronize()
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / 100
            try:
                if variant == li.VARIANT_GATHER:
                    tg = graph_us(lambda: li.apply_b(1, li.Direct, A, x, False, 0, out))
                    report(f"#1 {tag} forward, CUDA-graph replay (L2-resident)", nrows, bytes_, tg, tg)
                tg = graph_us(lambda: li.apply_b(1, li.Adjoint, A, y, False, 0, back))
                report(f"#1 {tag} adjoint {vname}, CUDA-graph replay (L2-resident)", nrows, bytes_, tg, tg)
            except Exception as exc:   # pragma: no cover
                print(json.dumps({"config": f"#1 {tag} graph replay", "error": f"{type(exc).__name__}: {exc}"}), flush=True)


def cfg3(li, torch, nb):
    T = np.float32
    n = 8192
    c = (4096.5, 4096.5)
    R = c + li.rotate(li.AffineTransform2D(T=T) - c, 0.2)
    ker = li.CatmullRomSpline(T)
    src = li.device_empty((n, n, nb), T)
    src.normal_()
    dst = li.device_empty((n, n, nb), T)
    for P, name, variant in ((li.Direct, "forward", li.VARIANT_AUTO), (li.Adjoint, "adjoint (gather form, deterministic)", li.VARIANT_AUTO),
                             (li.Adjoint, "adjoint (atomic variant)", li.VARIANT_ATOMIC)):
        fn = lambda: li.affine_apply(1, P, ker, ker, R, src, 0, dst, variant=variant, nbatch_dims=1)
        t = timeit(fn, 5, warm=2)
        report(f"#3 affine rotation CatmullRom F32 8192^2 x {nb} {name}", n * n * nb, 2 * n * n * 4 * nb, *t)
    # one image at a time (the reference's own call shape: the type is strictly 2-D)
    one_s, one_d = src[:, :, 0], dst[:, :, 0]
    for variant, name in ((li.VARIANT_AUTO, "gather form"), (li.VARIANT_ATOMIC, "atomic variant")):
        fn = lambda: li.affine_apply(1, li.Adjoint, ker, ker, R, one_s, 0, one_d, variant=variant)
        t = timeit(fn, 5, warm=2)
        report(f"#3 affine rotation, adjoint of a single 8192^2 image, {name}", n * n, 2 * n * n * 4, *t)


def cfg4(li, torch):
    T = np.float32
    n = 1024
    kers = (li.CubicSpline(T), li.MitchellNetravaliSpline(T), li.LinearSpline(T))
    deltas = (0.37, 0.11, 0.73)
    A = [li.TabulatedInterpolator(T, d + 1, kers[d], 1 + np.arange(n) * 0.98 + deltas[d], n) for d in range(3)]
    H = A[0] * A[1] * A[2]
    x = li.device_empty((n, n, n), T)
    x.normal_()
    y = li.device_empty((n, n, n), T)
    for P, name in ((li.Direct, "forward"), (li.Adjoint, "adjoint")):
        fn = lambda: li.apply_b(1, P, H, x, False, 0, y)
        t = timeit(fn, 3, warm=1)
        report(f"#4 3-D separable 1024^3 F32 (CubicSpline, MitchellNetravali, Linear) {name}, one fused pass (li_sep3d)",
               n ** 3, 2 * 4 * n ** 3, *t)


def cfg5(li, torch, nframes):
    T = np.float32
    m = 2048
    ker = li.SafeFlat(li.LanczosKernel(T, 6))
    xs = np.arange(1, m + 1, dtype=np.float64)
    rng = np.random.default_rng(5)
    ops = []
    for k in range(nframes):
        d1, d2 = rng.uniform(-8, 8, 2)
        ops.append(li.TabulatedInterpolator(T, 1, ker, xs + d1, m) * li.TabulatedInterpolator(T, 2, ker, xs + d2, m))
    frames = li.device_empty((m, m, nframes), T)
    frames.normal_()
    views = [frames[:, :, k] for k in range(nframes)]
    acc = li.device_empty((m, m), T)

    def run():
        acc.zero_()
        for k in range(nframes):
            li.apply_b(1, li.Adjoint, ops[k], views[k], False, 1, acc)
    t = timeit(run, 3, warm=1)
    report(f"#5 adjoint of {nframes} x 2048^2 frames (Lanczos-6, SafeFlat) accumulated into one model, 1 GPU, one apply! per frame",
           nframes * m * m, nframes * m * m * 4 + m * m * 4, *t)
    tabs = [[f.table for f in op.factors] for op in ops]

    def run_batch():
        acc.zero_()
        li.sep_apply_batch(True, tabs, 1, views, 1, [acc] * nframes)
    t = timeit(run_batch, 3, warm=1)
    report(f"#5 same, li_sep_apply_batch (one library call for all {nframes} frames; handles marshalled per call)",
           nframes * m * m, nframes * m * m * 4 + m * m * 4, *t)
    batch = li.FrameBatch([tuple(t) for t in tabs], views)   # handles and pointers marshalled once, as a fit would

    def run_multi():
        li.adjoint_accumulate(batch, acc)
    t = timeit(run_multi, 3, warm=1)
    report(f"#5 same, li_adjoint_accumulate_allreduce on a FrameBatch (multi-frame kernel, one launch)",
           nframes * m * m, nframes * m * m * 4 + m * m * 4, *t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--batch3", type=int, default=16)
    ap.add_argument("--frames5", type=int, default=128)
    args = ap.parse_args()
    import torch
    import li_b200 as li
    which = set(args.configs.split(","))
    if "1" in which:
        cfg1(li, torch)
    if "3" in which:
        cfg3(li, torch, args.batch3)
    if "4" in which:
        cfg4(li, torch)
    if "5" in which:
        cfg5(li, torch, args.frames5)


if __name__ == "__main__":
    main()

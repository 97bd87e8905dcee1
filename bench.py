#!/usr/bin/env python
"""Headline benchmark: 2-D separable Lanczos-4 resampling, Float32 4096^2 -> 8192^2, forward + adjoint
(BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One *step* = one forward apply (H*x, 8192^2 output samples) plus one adjoint apply (H'*y, 8192^2 input
samples) of H = A1*A2 on one image; with N GPUs every rank processes its own image (the path shards over
the batch dimension with no collective; "scaling": "weak").  One sample = one element on the
interpolated ("rows") side.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of the reference's loops (oracle/, all host threads) on
the same workload: the reference itself is Julia and cannot run in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = 4096           # source is M x M
N_OUT = 8192       # result is N_OUT x N_OUT
S = 4              # Lanczos-4
SAMPLES_PER_STEP = 2 * N_OUT * N_OUT                 # forward outputs + adjoint inputs
# algorithmic (compulsory) bytes of ONE apply, forward or adjoint:
# |src| + |dst| + tables (int32 base + Float32 weights per row, both dimensions; beta = 0)
TABLE_BYTES = 2 * N_OUT * (4 + 4 * S)
ALGO_BYTES_PER_APPLY = 4 * M * M + 4 * N_OUT * N_OUT + TABLE_BYTES
METRIC = "Gsamples/s fwd+adjoint, 2D separable Lanczos-4 (fraction of HBM roofline)"
WORKLOAD = "2D separable TabulatedInterpolator Lanczos-4, Float32 4096^2->8192^2, forward + adjoint"


def common_config():
    """The `config` object both arms print (identical dicts, so the driver can compare them)."""
    return {"workload": WORKLOAD, "alpha": 1, "beta": 0, "images_per_gpu": 1}


def host_threads() -> int:
    """Threads the CPU arm uses: every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it
    to 1 for N > 1, which made the round-1 ratios incomparable across N)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_operator(li):
    x = np.linspace(1.0, float(M), N_OUT)       # range(1, 4096, length=8192), end-point aligned
    ker = li.LanczosKernel(np.float32, S)
    A1 = li.TabulatedInterpolator(np.float32, 1, ker, x, M)
    A2 = li.TabulatedInterpolator(np.float32, 2, ker, x, M)
    return A1, A2, A1 * A2


# ------------------------------------------------------------------------------------------------------
def cpu_reference_step(O, J, W, src, y, nthreads, rows_frac=1.0):
    """The reference's two-pass schedule on the CPU (A2 then A1 forward; A1' then A2' adjoint), loops of
    src/tabulated.jl:409-416,455-463, multi-threaded over the non-interpolated index."""
    t0 = time.perf_counter()
    tmp = O.tab_direct(J, W, M, 2, src, nthreads=nthreads)
    out = O.tab_direct(J, W, M, 1, tmp, nthreads=nthreads)
    t1 = O.tab_adjoint(J, W, M, 1, y, nthreads=nthreads)
    back = O.tab_adjoint(J, W, M, 2, t1, nthreads=nthreads)
    return time.perf_counter() - t0, out, back


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import li_oracle as O
    nthreads = host_threads()
    rng = np.random.default_rng(0xB200)
    x = np.linspace(1.0, float(M), N_OUT)
    J, W = O.maketable(O.kernel(O.LANCZOS, S), x, M, np.float32)
    src = np.asfortranarray(rng.standard_normal((M, M), dtype=np.float32))
    y = np.asfortranarray(rng.standard_normal((N_OUT, N_OUT), dtype=np.float32))
    # bound the whole run to a few minutes: calibrate on one step
    t_one, _, _ = cpu_reference_step(O, J, W, src, y, nthreads)
    steps, warmup = args.steps, args.warmup
    budget = 120.0
    if t_one * (steps + warmup) > budget:
        steps = max(1, int(budget / t_one) - 1)
        warmup = min(warmup, 1)
    for _ in range(warmup):
        cpu_reference_step(O, J, W, src, y, nthreads)
    total = 0.0
    for _ in range(steps):
        dt, _, _ = cpu_reference_step(O, J, W, src, y, nthreads)
        total += dt
    ms = 1e3 * total / steps
    value = SAMPLES_PER_STEP / (ms * 1e-3) / 1e9
    # the reference itself is single-threaded (no Threads.@threads anywhere in /root/reference/src): one step on one core
    t_single, _, _ = cpu_reference_step(O, J, W, src, y, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": common_config(),
        "notes": {"schedule": "two passes per apply, as the reference composes A1*A2",
                  "what": "CPU restatement of the reference loops (oracle/); Julia is not installed in this image",
                  "threads": f"{nthreads} OpenMP threads (all cores of the box, independent of OMP_NUM_THREADS)",
                  "single_thread_value": SAMPLES_PER_STEP / t_single / 1e9,
                  "single_thread_note": "1 thread = what the single-threaded reference does; 1 full-size step"},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": nthreads, "kind": "port",
                         "sample": f"full workload, {steps} steps (requested {args.steps})"},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def run_multi(li, torch, dist, rank, world, dev, frames_total=512, reps=5):
    """BASELINE.json config #5, strong scaling: the adjoint of `frames_total` observed 2048^2 frames (per-frame
    Lanczos-6 SafeFlat interpolators, shifts U(-8, 8)) accumulated into ONE shared 2048^2 model.  Frames are
    block-sharded over the ranks; every rank makes ONE call of the C ABI's li_adjoint_accumulate_allreduce (multi-frame
    accumulate kernel, model produced in 4 bands, NCCL all-reduce of each band on the library's own communicator
    overlapping the next band).  Times: CUDA events, max over ranks.  `allreduce_ms_exposed` = that time minus the
    time of the same local work without a communicator."""
    T = np.float32
    m = 2048
    lo, hi = li.shard_range(frames_total, rank, world)
    ker = li.SafeFlat(li.LanczosKernel(T, 6))
    xs = np.arange(1, m + 1, dtype=np.float64)
    shifts = np.random.default_rng(5).uniform(-8, 8, (frames_total, 2))
    tabs = []
    for k in range(lo, hi):
        tabs.append((li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m).table,
                     li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m).table))
    frames = li.device_empty((m, m, hi - lo), T, dev)
    frames.normal_(generator=torch.Generator(device=dev).manual_seed(505 + rank))
    views = [frames[:, :, k] for k in range(hi - lo)]
    model = li.device_empty((m, m), T, dev)
    comm = li.Comm.from_torch_distributed() if world > 1 else None

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        return min(ts), float(np.median(ts))

    batch = li.FrameBatch(tabs, views)        # (handles and pointers marshalled once, as an iterative fit would)
    l0 = li.launch_count()
    li.adjoint_accumulate(batch, model, comm=comm, nbands=4)
    torch.cuda.synchronize()
    launches = li.launch_count() - l0        # (includes the one-off table analysis of the first call)
    l0 = li.launch_count()
    li.adjoint_accumulate(batch, model, comm=comm, nbands=4)
    launches = li.launch_count() - l0
    # bands of model rows: the all-reduce of a band overlaps the computation of the next one, but every band is its own
    # launch over all local frames; which wins depends on the frames per rank -- measured, the best one is reported
    by_bands = {}
    for nb in ((1, 2, 4) if comm is not None else (1,)):
        by_bands[nb] = timed(lambda: li.adjoint_accumulate(batch, model, comm=comm, nbands=nb))
    best_nb = min(by_bands, key=lambda k: by_bands[k][0])
    t_all, t_all_med = by_bands[best_nb]
    t_loc, t_loc_med = timed(lambda: li.adjoint_accumulate(batch, model, comm=None))
    t_ar = None
    if comm is not None:       # the all-reduce of the whole model on its own (NCCL, in place)
        t_ar, _ = timed(lambda: comm.allreduce_sum(model))
    # the per-frame path of round 1 (one launch per frame), for the record
    li.set_option("no_multi", 1)
    try:
        t_old, _ = timed(lambda: li.adjoint_accumulate(batch, model, comm=None))
    finally:
        li.set_option("no_multi", 0)
    finite = bool(torch.isfinite(model).all().item())
    peak, _ = measured_peak_gbs()
    nloc = hi - lo
    algo = nloc * m * m * 4 + m * m * 4 + nloc * 2 * m * (4 + 4 * 6)     # local frames + model + tables (base + W)
    if comm is not None:
        comm.destroy()
    return {"workload": f"config #5: adjoint of {frames_total} x 2048^2 frames (Lanczos-6, SafeFlat, shifts +-8) into one shared "
                        f"2048^2 model; frames sharded over {world} GPU(s); NCCL all-reduce inside libli_b200.so, "
                        f"{best_nb} band(s)",
            "ms_by_bands": {str(k): v[0] for k, v in by_bands.items()}, "allreduce_alone_ms": t_ar,
            "scaling": "strong", "n_gpus": world, "frames": frames_total, "frames_per_gpu": nloc,
            "ms": t_all, "ms_median": t_all_med, "ms_local_only": t_loc, "allreduce_ms_exposed": max(0.0, t_all - t_loc),
            "us_per_frame_local": 1e3 * t_loc / max(nloc, 1), "ms_one_launch_per_frame": t_old,
            "Gsamples_s": frames_total * m * m / (t_all * 1e-3) / 1e9,
            "frac": algo / (t_loc * 1e-3) / 1e9 / peak, "achieved_gbs_local": algo / (t_loc * 1e-3) / 1e9,
            "algorithmic_bytes_per_gpu": algo, "kernel": "sep2d_adj_multi_kernel", "launches_per_step": int(launches),
            "allreduce_bytes": m * m * 4, "results_finite": finite}


# ------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import li_b200 as li

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    # pinned host buffers of the e2e leg: first-touched on the NUMA node of this rank's GPU
    numa = li.bind_host_to_device(local_rank) if world > 1 and not args.no_numa_bind else {"bound": False, "why": "single rank"}
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    A1, A2, H = make_operator(li)
    gen = torch.Generator(device=dev).manual_seed(0xB200 + rank)
    NBUF = 2   # rotate buffer sets so that no timed kernel finds its input in L2 (one set = 671 MB > 126 MB L2)
    xs = [li.device_empty((M, M), np.float32, dev) for _ in range(NBUF)]
    ys_in = [li.device_empty((N_OUT, N_OUT), np.float32, dev) for _ in range(NBUF)]
    ys_out = [li.device_empty((N_OUT, N_OUT), np.float32, dev) for _ in range(NBUF)]
    zs = [li.device_empty((M, M), np.float32, dev) for _ in range(NBUF)]
    for t in xs + ys_in:
        t.normal_(generator=gen)

    def step(k, events=None):
        b = k % NBUF
        if events is not None:
            events[0].record()
        li.apply_b(1, li.Direct, H, xs[b], False, 0, ys_out[b])       # forward:  8192^2 outputs
        if events is not None:
            events[1].record()
        li.apply_b(1, li.Adjoint, H, ys_in[b], False, 0, zs[b])       # adjoint:  8192^2 inputs
        if events is not None:
            events[2].record()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(args.warmup, 3)):
        step(k)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = li.launch_count()
    barrier()
    t_host0 = time.perf_counter()
    for k in range(args.steps):
        step(k, evs[k])
    barrier()
    t_host = time.perf_counter() - t_host0
    launches = li.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    fwd_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    adj_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    eager_ms_per_step = total_ms / args.steps

    # The same K steps replayed from CUDA graphs (one captured step per buffer set): the two kernels of a step take
    # ~140 us while Python needs ~85 us to issue them, so with 8 ranks sharing the host cores the eager loop can
    # become host-bound; the graphs take the host out of the timed region.  Falls back to the eager time if the
    # capture is refused.
    ms_per_step, timed_how = eager_ms_per_step, "eager calls through the public API"
    fwd_ms_eager, adj_ms_eager = fwd_ms, adj_ms
    per_launch_how = "CUDA events around each apply in the eager pass of the same K steps"
    try:
        graphs = []
        cap_stream = torch.cuda.Stream(dev)
        cap_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(cap_stream):
            for b in range(NBUF):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream):
                    step(b)
                graphs.append(g)
        torch.cuda.current_stream(dev).wait_stream(cap_stream)
        for k in range(4):
            graphs[k % NBUF].replay()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for k in range(args.steps):
            graphs[k % NBUF].replay()
        g1.record()
        barrier()
        ms_per_step = g0.elapsed_time(g1) / args.steps
        timed_how = "CUDA-graph replay of the public-API calls (one captured step per buffer set)"
        launches = 2 * args.steps
        # per-kernel launch durations: the same launches, one kernel type per graph (back to back on the stream, the
        # buffer sets rotating), CUDA events around K launches.  An event pair around every single eager launch adds
        # the event records and the launch gap to a 65 us kernel (kept as *_eager_events).
        def per_launch(which):
            with torch.cuda.stream(cap_stream):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=cap_stream):
                    for b in range(NBUF):
                        if which == "fwd":
                            li.apply_b(1, li.Direct, H, xs[b], False, 0, ys_out[b])
                        else:
                            li.apply_b(1, li.Adjoint, H, ys_in[b], False, 0, zs[b])
            torch.cuda.current_stream(dev).wait_stream(cap_stream)
            reps = max(2, args.steps // NBUF)
            for _ in range(2):
                g.replay()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            barrier()
            return e0.elapsed_time(e1) / (reps * NBUF)
        fwd_ms_eager, adj_ms_eager = fwd_ms, adj_ms
        fwd_ms, adj_ms = per_launch("fwd"), per_launch("adj")
        per_launch_how = "CUDA events around K back-to-back graph-replayed launches of one kernel (buffer sets rotating)"
    except Exception as exc:  # pragma: no cover
        timed_how += f" (graph capture refused: {type(exc).__name__})"
    if dist is not None:
        t = torch.tensor([ms_per_step, fwd_ms, adj_ms, eager_ms_per_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, fwd_ms, adj_ms, eager_ms_per_step = (float(v) for v in t.tolist())
    value = world * SAMPLES_PER_STEP / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region
    xh = li.pinned_empty((M, M), np.float32)
    yh_in = li.pinned_empty((N_OUT, N_OUT), np.float32)
    xh[...] = np.asfortranarray(np.random.default_rng(rank).standard_normal((M, M), dtype=np.float32))
    yh_in[...] = 1.0
    yh = li.vcreate(li.Direct, H, xh)          # pinned, like its argument
    zh = li.vcreate(li.Adjoint, H, yh_in)
    e2e_steps = max(2, min(args.steps, 5))
    # (a) blocking calls, one after the other -- the reference's calling convention
    for _ in range(2):
        li.apply_b(1, li.Direct, H, xh, False, 0, yh)
        li.apply_b(1, li.Adjoint, H, yh_in, False, 0, zh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        li.apply_b(1, li.Direct, H, xh, False, 0, yh)
        li.apply_b(1, li.Adjoint, H, yh_in, False, 0, zh)
    torch.cuda.synchronize()
    e2e_blocking_s = (time.perf_counter() - t0) / e2e_steps
    # (b) the two independent applies of a step on two host streams: the download of the forward result overlaps
    #     the upload of the adjoint's input (PCIe is full duplex); same public API, `with li.async_host(stream)`
    sa, sb = li.HostStream(), li.HostStream()

    def e2e_step():
        with li.async_host(sa):
            li.apply_b(1, li.Direct, H, xh, False, 0, yh)
        with li.async_host(sb):
            li.apply_b(1, li.Adjoint, H, yh_in, False, 0, zh)
        sa.synchronize()
        sb.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s, e2e_blocking_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_blocking_s = (float(v) for v in t.tolist())
    e2e_value = world * SAMPLES_PER_STEP / e2e_s / 1e9
    e2e_ok = bool(np.isfinite(yh[::997, ::991]).all() and np.isfinite(zh[::97, ::91]).all())

    # ---- config #5 (the path's one collective), through the C ABI; every rank takes part
    multi = None
    if not args.no_multi:
        del xs, ys_in, ys_out, zs
        torch.cuda.empty_cache()
        try:
            multi = run_multi(li, torch, dist, rank, world, dev, frames_total=args.frames)
        except Exception as exc:  # pragma: no cover - reported, never hidden
            multi = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- CPU baseline on rank 0 (bounded sample: one full-size step, all host threads)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import li_oracle as O
        nthreads = host_threads()
        J, W = O.maketable(O.kernel(O.LANCZOS, S), np.linspace(1.0, float(M), N_OUT), M, np.float32)
        y_np = np.asfortranarray(np.ones((N_OUT, N_OUT), dtype=np.float32))
        dt, out_ref, back_ref = cpu_reference_step(O, J, W, np.asarray(xh), y_np, nthreads)
        # the same step doubles as a full-size parity check of the e2e results
        err_f = float(np.max(np.abs(out_ref[::61, ::67] - yh[::61, ::67])) / np.max(np.abs(out_ref)))
        err_a = float(np.max(np.abs(back_ref[::31, ::37] - zh[::31, ::37])) / np.max(np.abs(back_ref)))
        cpu = {"value": SAMPLES_PER_STEP / dt / 1e9, "unit": "Gsamples/s", "cores": nthreads, "kind": "port",
               "sample": "1 full-size step (fwd+adj, two-pass schedule of the reference), all host threads",
               "parity_rel_err": {"forward": err_f, "adjoint": err_a}}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        dom_ms, dom = (adj_ms, "sep2d_adj_stream_kernel") if adj_ms >= fwd_ms else (fwd_ms, "sep2d_fwd_kernel")
        achieved = ALGO_BYTES_PER_APPLY / (dom_ms * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/), if any
        traffic, traffic_from = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_from = tj.get(dom), tj.get("captured_at")
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(),
            "notes": {"l2": "inputs rotate over 2 buffer sets of 671 MB each (> 126 MB L2); no timed kernel re-reads fresh data",
                      "timed_region": timed_how, "eager_ms_per_step": eager_ms_per_step,
                      "per_launch_times": per_launch_how,
                      "fwd_ms_eager_events": fwd_ms_eager, "adj_ms_eager_events": adj_ms_eager,
                      "host_wall_ms_per_step": 1e3 * t_host / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_captured_at": traffic_from,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_APPLY,
                         "fwd_ms": fwd_ms, "adj_ms": adj_ms,
                         "fwd_frac": ALGO_BYTES_PER_APPLY / (fwd_ms * 1e-3) / 1e9 / peak,
                         "adj_frac": ALGO_BYTES_PER_APPLY / (adj_ms * 1e-3) / 1e9 / peak,
                         "frac_of_8TBs_nominal": achieved / 8000.0},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": 4 * M * M + 4 * N_OUT * N_OUT,
                    "d2h_bytes_per_step": 4 * N_OUT * N_OUT + 4 * M * M, "ms_per_step": 1e3 * e2e_s,
                    "how": "pinned host buffers; forward and adjoint of a step on two host streams (li.async_host)",
                    "pcie_gbs_per_rank": (8 * M * M + 8 * N_OUT * N_OUT) / e2e_s / 1e9, "host_numa_binding": numa,
                    "blocking_ms_per_step": 1e3 * e2e_blocking_s,
                    "blocking_value": world * SAMPLES_PER_STEP / e2e_blocking_s / 1e9, "results_finite": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "multi": multi,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin each rank to the cores local to its GPU")
    ap.add_argument("--no-multi", action="store_true", help="skip the config #5 block (multi-frame accumulate + all-reduce)")
    ap.add_argument("--frames", type=int, default=512, help="frames of the config #5 block (all GPUs together)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

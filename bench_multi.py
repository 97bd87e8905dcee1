"""Sharded BASELINE configs on N GPUs of one node (one process per GPU, NCCL):

  #5  iterative-fitting step: adjoint of 512 observed 2048^2 frames (per-frame Lanczos-6 interpolators) accumulated
      into one shared 2048^2 model.  Frames are block-sharded over the ranks (strong scaling); each rank accumulates
      its frames with ONE li_sep_apply_batch call, then a single NCCL sum all-reduce of the 16.8 MB model.
  #3  affine rotation of a batch of 64 x 8192^2 images, images block-sharded over the ranks, no collective.

    python bench_multi.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_multi.py

Rank 0 prints one JSON line per config; times are CUDA-event times, max over ranks."""
import argparse
import json
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="5,3")
    args = ap.parse_args()
    import torch
    import li_b200 as li
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    T = np.float32

    def timed(fn, reps):
        fn()
        best = []
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
            best.append(float(t.item()))
        return min(best), float(np.median(best))

    which = set(args.configs.split(","))
    if "5" in which:
        m = 2048
        lo, hi = li.shard_range(args.frames, rank, world)
        ker = li.SafeFlat(li.LanczosKernel(T, 6))
        xs = np.arange(1, m + 1, dtype=np.float64)
        rng = np.random.default_rng(5)
        shifts = rng.uniform(-8, 8, (args.frames, 2))
        tabs = []
        for k in range(lo, hi):
            A1 = li.TabulatedInterpolator(T, 1, ker, xs + shifts[k, 0], m)
            A2 = li.TabulatedInterpolator(T, 2, ker, xs + shifts[k, 1], m)
            tabs.append([A1.table, A2.table])
        frames = li.device_empty((m, m, hi - lo), T, dev)
        frames.normal_()
        views = [frames[:, :, k] for k in range(hi - lo)]
        acc = li.device_empty((m, m), T, dev)
        tcomp = [0.0]

        def step():
            acc.zero_()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            li.sep_apply_batch(True, tabs, 1, views, 1, [acc] * len(views))
            c1.record()
            if dist is not None:
                dist.all_reduce(li.distributed._dense(acc), op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
            tcomp[0] = c0.elapsed_time(c1)
        tmin, tmed = timed(step, args.reps)
        if rank == 0:
            print(json.dumps({"config": f"#5 adjoint of {args.frames} x 2048^2 frames (Lanczos-6, SafeFlat) into one shared model, "
                                        f"frames sharded over {world} GPU(s), NCCL sum all-reduce of the model",
                              "n_gpus": world, "ms_min": tmin, "ms_median": tmed, "ms_local_accumulate_rank0": tcomp[0],
                              "frames_per_s": args.frames / (tmin * 1e-3), "Gsamples_s": args.frames * m * m / (tmin * 1e-3) / 1e9,
                              "allreduce_bytes": m * m * 4, "scaling": "strong"}), flush=True)
        del frames, views, tabs, acc
        torch.cuda.empty_cache()
    if "3" in which:
        n = 8192
        lo, hi = li.shard_range(args.images, rank, world)
        nb = hi - lo
        ker = li.CatmullRomSpline(T)
        c = ((n + 1) / 2, (n + 1) / 2)
        R = c + li.rotate(li.AffineTransform2D(T=T) - c, 0.2)
        src = li.device_empty((n, n, nb), T, dev)
        src.normal_()
        dst = li.device_empty((n, n, nb), T, dev)
        for P, name in ((li.Direct, "forward"), (li.Adjoint, "adjoint")):
            tmin, tmed = timed(lambda: li.affine_apply(1, P, ker, ker, R, src, 0, dst, nbatch_dims=1), args.reps)
            if rank == 0:
                print(json.dumps({"config": f"#3 affine rotation CatmullRom F32 8192^2 x {args.images} images {name}, images sharded over "
                                            f"{world} GPU(s)", "n_gpus": world, "ms_min": tmin, "ms_median": tmed,
                                  "Gsamples_s": args.images * n * n / (tmin * 1e-3) / 1e9, "scaling": "strong"}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

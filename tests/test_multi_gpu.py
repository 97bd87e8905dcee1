"""N > 1 on real GPUs (run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`): frames sharded
over ranks, adjoint accumulated into a shared model with an NCCL all-reduce (BASELINE config #5 in small), and a
batch of images sharded over ranks for the forward operator (config #3 in small).  Skipped with fewer than 2 GPUs."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame_ops(li, T, m, k):
    # frame k observes the model shifted by (d1, d2) in (-8, 8): positions x_d(i) = i + d, Lanczos-6, SafeFlat
    rng = np.random.default_rng(1000 + k)
    d1, d2 = rng.uniform(-8, 8, 2)
    ker = li.SafeFlat(li.LanczosKernel(T, 6))
    x = np.arange(1, m + 1, dtype=np.float64)
    return (li.TabulatedInterpolator(T, 1, ker, x + d1, m), li.TabulatedInterpolator(T, 2, ker, x + d2, m), d1, d2)


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import li_b200 as li
    import li_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    T, m, nframes = np.float32, 96, 10
    rng = np.random.default_rng(7)
    frames = rng.standard_normal((nframes, m, m)).astype(T)
    ops = {}

    def adjoint_into(k, acc):
        A1, A2, _, _ = ops.setdefault(k, _frame_ops(li, T, m, k))
        y = li.to_device(np.asfortranarray(frames[k]))
        li.apply_b(1, li.Adjoint, A1 * A2, y, False, 1, acc)

    ok = True
    for groups in (1, 2):
        acc = li.device_empty((m, m), T)
        acc.zero_()
        li.accumulate_adjoint(nframes, adjoint_into, acc, groups=groups)
        got = li.to_host(acc)
        if rank == 0:
            want = np.zeros((m, m), dtype=np.float64)
            for k in range(nframes):
                _, _, d1, d2 = _frame_ops(li, T, m, k)
                x = np.arange(1, m + 1, dtype=np.float64)
                kd = O.kernel(O.LANCZOS, 6, O.SAFEFLAT)
                J1, W1 = O.maketable(kd, (x + d1).astype(T), m, T)
                J2, W2 = O.maketable(kd, (x + d2).astype(T), m, T)
                yk = np.asfortranarray(frames[k])
                want += O.tab_adjoint(J2, W2, m, 2, O.tab_adjoint(J1, W1, m, 1, yk)).astype(np.float64)
            err = np.max(np.abs(got - want)) / np.max(np.abs(want))
            ok = ok and err <= 1e-5
    # the same exchange step through the C ABI: the multi-frame kernel on the local frames + the library's own
    # all-reduce (peer memory over NVLink when the ranks can map each other's buffers, else NCCL)
    lo, hi = li.shard_range(nframes, rank, world)
    loc = [_frame_ops(li, T, m, k) for k in range(lo, hi)]
    tabs = [(A1.table, A2.table) for A1, A2, _, _ in loc]
    views = [li.to_device(np.asfortranarray(frames[k])) for k in range(lo, hi)]
    results = {}
    fails = []
    model0 = np.asfortranarray(rng.standard_normal((m, m)).astype(T))         # same on every rank (same seed)
    for name, exp in (("peer", 0), ("nccl", 8)):
        li.set_option("exp", exp)                 # (read when the communicator first needs its exchange buffers)
        comm = li.Comm.from_torch_distributed()
        for beta in (0.0, 0.5):
            model = li.to_device(model0.copy(order="F"))
            li.adjoint_accumulate(tabs, views, model, alpha=1, beta=beta, comm=comm)
            results[(name, beta)] = li.to_host(model)
            if rank == 0:
                ref = beta * model0.astype(np.float64) + want
                err = np.max(np.abs(results[(name, beta)] - ref)) / np.max(np.abs(ref))
                if not err <= 1e-5:
                    fails.append(f"{name} beta={beta}: rel err {err:.3e}")
        big = torch.full((1 << 19,), float(rank + 1), device="cuda", dtype=torch.float32)      # 2 MB: the peer path
        small = torch.full((1000,), float(rank + 1), device="cuda", dtype=torch.float64)      # NCCL
        comm.allreduce_sum(big)
        comm.allreduce_sum(small)
        tot = world * (world + 1) / 2
        if not bool((big == tot).all().item()):
            fails.append(f"{name}: all-reduce of the large buffer")
        if not bool((small == tot).all().item()):
            fails.append(f"{name}: all-reduce of the small buffer")
        torch.cuda.synchronize()
        comm.destroy()
    li.set_option("exp", 0)
    for beta in (0.0, 0.5):      # two ranks: a + b in either order, so both all-reduces give the same bits
        if not np.array_equal(results[("peer", beta)], results[("nccl", beta)]):
            fails.append(f"peer and nccl results differ, beta={beta}: {np.max(np.abs(results[('peer', beta)] - results[('nccl', beta)])):.3e}")
    # batch of images sharded over ranks (forward, no communication)
    nb = 6
    lo, hi = li.shard_range(nb, rank, world)
    imgs = rng.standard_normal((m, m, nb)).astype(T)
    A1, A2, _, _ = _frame_ops(li, T, m, 0)
    out = li.device_empty((m, m, hi - lo), T)
    li.sep_apply_tables(False, [A1.table, A2.table], 1, li.to_device(np.asfortranarray(imgs[:, :, lo:hi])), 0, out,
                        nbatch_dims=1)
    ref = O.tab_direct(A1.J, A1.W, m, 1, O.tab_direct(A2.J, A2.W, m, 2, np.asfortranarray(imgs[:, :, lo:hi])))
    ok = ok and np.max(np.abs(li.to_host(out) - ref)) <= 2e-6 * np.max(np.abs(ref))
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write("ok" if ok and not fails else "bad: " + "; ".join(fails))
    dist.barrier()
    dist.destroy_process_group()


def test_accumulate_and_shard_on_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"r{r}.txt").read_text() == "ok"

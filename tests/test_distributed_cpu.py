"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: block sharding and the accumulate + all-reduce
exchange step of iterative fitting (BASELINE config #5).  The per-frame adjoint is a numpy stand-in here (the CUDA
kernels need a GPU); the -m gpu suite runs the same driver with the real operators."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    import li_b200 as li
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 4, 8):
            blocks = [li.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        li.shard_range(4, 2, 2)


def _worker(rank, world, port, groups, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import li_b200 as li
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nframes, shape = 11, (6, 5)
    rng = np.random.default_rng(0)
    frames = rng.standard_normal((nframes,) + shape)
    seen = []

    def adjoint_into(k, acc):            # stand-in for apply_b(1, Adjoint, A_k, y_k, False, 1, acc)
        seen.append(k)
        acc.add_(torch.from_numpy(frames[k] * (k + 1)))

    acc = torch.zeros(shape, dtype=torch.float64)
    li.accumulate_adjoint(nframes, adjoint_into, acc, groups=groups)
    want = sum(frames[k] * (k + 1) for k in range(nframes))
    ok = np.allclose(acc.numpy(), want, rtol=1e-13, atol=1e-13)
    lo, hi = li.shard_range(nframes, rank, world)
    ok = ok and seen == list(range(lo, hi))
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("groups", [1, 3])
def test_accumulate_adjoint_allreduce_gloo(tmp_path, groups):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000) + groups
    mp.spawn(_worker, args=(world, port, groups, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"r{r}.txt").read_text() == "ok"

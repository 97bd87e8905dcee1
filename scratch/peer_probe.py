"""All-reduce of the config #5 model (16.8 MB) on its own: the library's peer-memory all-reduce against NCCL (exp bit 8).
torchrun --nproc-per-node N scratch/peer_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import li_b200 as li
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for name, e in (("peer", 0), ("nccl", 8), ("peer", 0)):
    li.set_option("exp", e)
    comm = li.Comm.from_torch_distributed()
    x = torch.ones(2048 * 2048, device="cuda", dtype=torch.float32)
    for _ in range(3):
        comm.allreduce_sum(x)
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(20):
        x.fill_(1.0)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); comm.allreduce_sum(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ok = bool((x == world).all().item())
    t = torch.tensor([float(np.median(ts)), float(min(ts))], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} ranks, {name}: median {t[0].item()*1e3:.1f} us, min {t[1].item()*1e3:.1f} us (max over ranks), correct {ok}", flush=True)
    comm.destroy()
dist.destroy_process_group()

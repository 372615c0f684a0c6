"""Aggregate host <-> device copy ceiling of the box with N GPUs copying at once (torchrun, one process per GPU):
pinned 1 GiB buffers, H2D alone, D2H alone, both directions together; max time over ranks.  One JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe_multi.py
"""
import json, os, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)

a = timed(lambda: d.copy_(h, non_blocking=True)); b = timed(lambda: h2.copy_(d2, non_blocking=True)); c = timed(both)
if rank == 0:
    print(json.dumps({"gpus": world, "h2d_GBps_aggregate": world * n / a / 1e9, "d2h_GBps_aggregate": world * n / b / 1e9,
                      "bidirectional_GBps_each_way_aggregate": world * n / c / 1e9, "host_cores": os.cpu_count()}), flush=True)
if world > 1: dist.destroy_process_group()

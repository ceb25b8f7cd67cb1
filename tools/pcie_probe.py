"""PCIe D2H / H2D bandwidth with pinned buffers, on ONE GPU or on N GPUs CONCURRENTLY (the e2e floor of bench.py):

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py

Under torchrun every rank copies on its own GPU at the same time; rank 0 prints the per-GPU minimum and the aggregate —
the host side (root complexes, memory channels) is shared, so the aggregate is what bounds e2e at N > 1.
"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 100_000_000  # 400 MB of int32 per buffer
d1 = torch.empty(n, dtype=torch.int32, device=dev); d2 = torch.empty(n, dtype=torch.int32, device=dev)
h1 = torch.empty(n, dtype=torch.int32, pin_memory=True); h2 = torch.empty(n, dtype=torch.int32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return dt


def one_d2h():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True); h2.copy_(d2, non_blocking=True)


def two_d2h():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def one_h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True); d2.copy_(h2, non_blocking=True)


def bidir():
    with torch.cuda.stream(s1):
        h1.copy_(d1, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)


res = {"gpus": world, "bytes_per_gpu_per_test": 800_000_000}
for name, fn in (("d2h_one_stream", one_d2h), ("d2h_two_streams", two_d2h), ("h2d_one_stream", one_h2d), ("d2h_plus_h2d", bidir)):
    t = run(fn)
    res[name] = {"per_gpu_gbs_slowest": round(0.8 / t, 2), "aggregate_gbs": round(0.8 * world / t, 2), "ms": round(t * 1e3, 2)}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.destroy_process_group()

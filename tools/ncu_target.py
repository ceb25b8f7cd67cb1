"""Small single-GPU run for ncu: one build + one search (auto radius) of a bench workload.

    python tools/ncu_target.py cfg2|cfg3|cfg4|<n_points> [k]

The first traverse_kernel launch of the search (round 1) is the one to capture:
    ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -c 1 -o gpurun_out/<name> python tools/ncu_target.py cfg2
(the start-radius sample runs the warp-per-query kernel, so `-k regex:traverse_kernel` skips it).
"""
import sys

import torch

sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets  # noqa: E402

CFG = {"cfg2": (10_000_000, 10, "uniform", 42), "cfg3": (10_000_000, 64, "lidar", 7), "cfg4": (100_000_000, 10, "uniform", 42)}
arg = sys.argv[1] if len(sys.argv) > 1 else "2000000"
if arg in CFG:
    n, k, cloud, seed = CFG[arg]
else:
    n, k, cloud, seed = int(arg), int(sys.argv[2]) if len(sys.argv) > 2 else 10, "uniform", 42
t = TrueKNN(0)
if cloud == "uniform":
    x = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    t.generate_uniform(seed, 0, n, out=x)
else:
    x = torch.from_numpy(datasets.lidar_like(n, seed)).cuda()
t.build(x)
cap = t.shard_capacity(1)
out = (torch.empty((cap,), dtype=torch.int32, device="cuda"), torch.empty((cap, k), dtype=torch.int32, device="cuda"),
       torch.empty((cap, k), dtype=torch.float32, device="cuda"))
t.search_shard(k, 0, 1, out=out)
s = t.stats()
print("ok", arg, n, k, s["search_ms"], s["rounds"], s["round_queries"], s["kernel_ms"])

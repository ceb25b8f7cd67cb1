"""End-to-end step (tknn_build + tknn_search with pinned host buffers) against the number of file-order slices:

    python tools/e2e_sweep.py [--cfg3] [--idx-only] <slices> ...      (cfg2 unless --cfg3)
"""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets
cfg3 = "--cfg3" in sys.argv
n, k = (10_000_000, 64) if cfg3 else (10_000_000, 10)
t = TrueKNN(0)
stream = torch.cuda.current_stream()
t.set_stream(stream.cuda_stream)
if cfg3:
    xd = torch.from_numpy(datasets.lidar_like(n, 7)).cuda()
else:
    xd = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    t.generate_uniform(42, 0, n, out=xd)
xh = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); xh.copy_(xd)
idx_h = torch.empty((n, k), dtype=torch.int32, pin_memory=True)
dst_h = torch.empty((n, k), dtype=torch.float32, pin_memory=True)
idx_only = "--idx-only" in sys.argv  # dist_out = NULL: indices only, half the device->host bytes
xn, out = xh.numpy(), (idx_h.numpy(), None if idx_only else dst_h.numpy())
for chunks in [int(a) for a in sys.argv[1:] if not a.startswith("--")] or [1, 3, 4, 5, 6, 8]:
    t.set_option("file_order_chunks", chunks)
    best = 1e9
    for it in range(3 if cfg3 else 4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t.build(xn)
        t.search(k, 0.0, out=out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if it > 0:
            best = min(best, dt)
    s = t.stats()
    print("cfg3" if cfg3 else "cfg2", "idx_only" if idx_only else "idx+dist", "chunks", chunks, "e2e_ms", round(best, 2), "search_ms", round(s["search_ms"], 2), "d2h_ms", round(s["d2h_ms"], 2), flush=True)

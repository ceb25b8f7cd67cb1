"""cfg2 / cfg3 search time and cfg2 end-to-end time against TKNN_OPT_WARP_ROUND_MAX."""
import sys
import time
import torch
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets
which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = 10_000_000
if which == "cfg2":
    k = 10
    x = torch.empty((n, 3), dtype=torch.float32, device="cuda")
else:
    k = 64
    x = torch.from_numpy(datasets.lidar_like(n, 7)).cuda()
for wmax in (0, 8192, 32768, 65536, 262144):
    t = TrueKNN(0, warp_round_max=wmax)
    stream = torch.cuda.current_stream()
    t.set_stream(stream.cuda_stream)
    if which == "cfg2":
        t.generate_uniform(42, 0, n, out=x)
    t.build(x)
    idx = torch.empty((n, k), dtype=torch.int32, device="cuda"); dst = torch.empty((n, k), dtype=torch.float32, device="cuda")
    best = None
    for _ in range(4):
        t.search(k, 0.0, out=(idx, dst))
        s = t.stats()
        if best is None or s["search_ms"] < best["search_ms"]:
            best = s
    line = f"{which} warp_round_max {wmax}: search {best['search_ms']:.3f} ms, estimate {best['estimate_ms']:.3f}, rounds {[round(v, 3) for v in best['round_ms'][:4]]} {best['round_queries'][:4]}"
    if which == "cfg2":
        xh = torch.empty((n, 3), dtype=torch.float32, pin_memory=True); xh.copy_(x)
        ih = torch.empty((n, k), dtype=torch.int32, pin_memory=True); dh = torch.empty((n, k), dtype=torch.float32, pin_memory=True)
        e = 1e9
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            t.build(xh.numpy()); t.search(k, 0.0, out=(ih.numpy(), dh.numpy()))
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
            if it:
                e = min(e, dt)
        line += f", e2e {e:.2f} ms"
        del xh, ih, dh
    print(line, flush=True)
    t.close()

"""cfg2 search time against the start-radius quantile (per mille)."""
import sys
import torch
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN
n, k = 10_000_000, 10
x = torch.empty((n, 3), dtype=torch.float32, device="cuda")
idx = torch.empty((n, k), dtype=torch.int32, device="cuda"); dst = torch.empty((n, k), dtype=torch.float32, device="cuda")
for q in (970, 980, 990, 993, 995, 997, 999):
    t = TrueKNN(0, radius_quantile=q)
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    t.generate_uniform(42, 0, n, out=x)
    t.build(x)
    best = None
    for _ in range(5):
        t.search(k, 0.0, out=(idx, dst))
        s = t.stats()
        if best is None or s["search_ms"] < best["search_ms"]:
            best = s
    print(f"quantile {q}: search {best['search_ms']:.3f} ms rounds {[round(v, 3) for v in best['round_ms'][:3]]} {best['round_queries'][:3]} r0 {best['start_radius']:.6f}", flush=True)
    t.close()

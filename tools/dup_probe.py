"""How bad are big clusters of exact duplicates? (boxdist == bound ties force visits)"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets
t = TrueKNN(0)
for n, d in ((2_000_000, 2_000), (2_000_000, 100_000), (2_000_000, 400_000)):
    x = datasets.uniform(n, 5)
    x[:d] = x[0]
    xd = torch.from_numpy(x).cuda()
    t.build(xd)
    for rep in range(2):
        t.search(10)
        s = t.stats()
        print(f"dev n={n} identical={d}: search {s['search_ms']:.2f} ms est {s['estimate_ms']:.2f} rounds={s['rounds']} round_ms={[round(v,2) for v in s['round_ms']]} kernel_ms={[round(v,2) for v in s['kernel_ms']]} q={s['round_queries']}", flush=True)
    t.search(10, float('inf')); s = t.stats(); print("   r=inf:", round(s['search_ms'],2))
    t.search(10, 0.01); s = t.stats(); print("   r=0.01:", round(s['search_ms'],2), s['rounds'])

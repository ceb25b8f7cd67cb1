"""Small single-GPU run for ncu: one build + one search (auto radius) on N uniform points."""
import sys

import torch

sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
x = torch.from_numpy(datasets.uniform(n, 42)).cuda()
t = TrueKNN(0)
t.build(x)
idx, dist = t.search(k)
s = t.stats()
print("ok", n, k, s["search_ms"], s["rounds"], s["round_ms"])

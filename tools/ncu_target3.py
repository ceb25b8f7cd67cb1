"""ncu target for the k = 64 LiDAR-like shape (cfg3 at reduced size): one build + one search."""
import sys
import torch
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
x = torch.from_numpy(datasets.lidar_like(n, 7)).cuda()
t = TrueKNN(0)
t.build(x)
t.search(64)
s = t.stats()
print("ok", n, s["search_ms"], s["rounds"], s["round_ms"])

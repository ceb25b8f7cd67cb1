"""Small all-paths run for compute-sanitizer: build, search (several k / radii), shard, query, range count, brute force."""
import sys

import numpy as np

sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets  # noqa: E402

x = datasets.lidar_like(20_000, seed=3)
t = TrueKNN(0)
t.build(x)
for k, r0 in ((5, 0.0), (10, 0.01), (30, 0.0), (64, float("inf"))):
    t.search(k, r0)
t.set_option("approx_filter", 1)
t.search(10)
t.set_option("approx_filter", 0)
t.search_shard(10, 1, 3)
q = datasets.uniform(3000, seed=1) * 50
t.query(q, 7)
t.query(q, 7, init_radius2=np.full(3000, 4.0, np.float32))
t.range_count(0.5)
t.brute_force(np.arange(0, 20_000, 50, dtype=np.int32), 10)
t.build(datasets.uniform(33, seed=1)).search(32)
t.close()
print("sanitize target ok")

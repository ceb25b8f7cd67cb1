"""Generates tests/golden/knn_golden.npz.

The reference ships no golden vectors for kNN and cannot be run here (OptiX), so
these fixtures are produced by the BRUTE-FORCE oracle (oracle/knn_oracle.c: tko_knn_brute — the
ground truth) and pin (a) the kd-tree oracle, (b) the distance formula and tie-break, (c) the
hash generator, against silent change.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from owlraytracing_b200 import datasets  # noqa: E402

out = {}
cases = {
    "uniform": (datasets.uniform(1500, seed=42), 10),
    "lidar": (datasets.lidar_like(1500, seed=7), 16),
    "lattice": (datasets.lattice(6), 7),
    "dups": (np.concatenate([datasets.uniform(300, seed=3)] * 3), 5),
}
for name, (x, k) in cases.items():
    idx, dist = O.knn_brute(x, k)
    out[f"{name}_x"], out[f"{name}_k"], out[f"{name}_idx"], out[f"{name}_dist"] = x, np.int32(k), idx, dist
out["hash_uniform_seed42_first64"] = datasets.uniform(64, seed=42)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "knn_golden.npz"), **out)
print("wrote", {k: v.shape for k, v in out.items()})

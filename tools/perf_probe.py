"""Quick single-GPU probe: build + search timings and traversal counters for a few configurations."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN, datasets  # noqa: E402


def probe(name, x, k, reps=3, **opts):
    xd = torch.from_numpy(x).cuda()
    t = TrueKNN(0, **opts)
    t.set_stream(torch.cuda.current_stream().cuda_stream)
    idx = torch.empty((x.shape[0], k), dtype=torch.int32, device="cuda")
    dist = torch.empty((x.shape[0], k), dtype=torch.float32, device="cuda")
    best = None
    for r in range(reps):
        t.build(xd)
        sb = t.stats()
        t.search(k, 0.0, out=(idx, dist))
        s = t.stats()
        if best is None or s["search_ms"] < best["search_ms"]:
            best = dict(s)
            best.update({kk: sb[kk] for kk in ("build_ms", "bounds_ms", "morton_ms", "sort_ms", "leaves_ms", "hierarchy_ms",
                                                "refit_ms", "n_leaves")})
    t.search(k, float("inf"), out=(idx, dist))
    sinf = t.stats()
    t.set_option("counters", 1)
    t.search(k, 0.0, out=(idx, dist))
    c = t.stats()
    n = x.shape[0]
    out = {
        "name": name, "opts": opts, "n": n, "k": k, "qps": n / (best["search_ms"] * 1e-3), "search_ms": best["search_ms"],
        "search_inf_ms": sinf["search_ms"],
        "estimate_ms": best["estimate_ms"], "rounds": best["rounds"], "round_ms": best["round_ms"],
        "round_queries": best["round_queries"], "r0": best["start_radius"],
        "build_ms": best["build_ms"], "phases": {p: round(best[p], 3) for p in ("bounds_ms", "morton_ms", "sort_ms", "leaves_ms",
                                                                                 "hierarchy_ms", "refit_ms")},
        "n_leaves": best["n_leaves"], "pts_per_leaf": n / best["n_leaves"],
        "nodes_per_q": c["nodes_visited"] / n, "tests_per_q": c["points_tested"] / n, "inserts_per_q": c["heap_inserts"] / n,
        "warp_nodes_per_group": c["warp_node_visits"] / (n / 32), "warp_leaves_per_group": c["warp_leaf_visits"] / (n / 32),
    }
    print(json.dumps(out), flush=True)
    t.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        x = datasets.uniform(1_000_000, 42)
        probe("uni1M", x, 10)
    elif which == "cfg2":
        x = datasets.uniform(10_000_000, 42)
        for leaf in (32, 16, 8):
            probe("cfg2", x, 10, leaf_size=leaf)
        probe("cfg2-fixed", x, 10, leaf_size=32, leaf_policy=1)
    elif which == "cfg3":
        x = datasets.lidar_like(10_000_000, 7)
        for leaf in (32, 16):
            probe("cfg3", x, 64, reps=2, leaf_size=leaf)

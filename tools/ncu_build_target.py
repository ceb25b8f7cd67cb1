"""Single-GPU run for ncu captures of the BUILD kernels: three builds of the 10 M-point uniform cloud (cfg2).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/<name> python tools/ncu_build_target.py
"""
import sys

import torch

sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t = TrueKNN(0)
x = torch.empty((n, 3), dtype=torch.float32, device="cuda")
t.generate_uniform(42, 0, n, out=x)
for _ in range(3):
    t.build(x)
s = t.stats()
print("ok", n, s["build_ms"], {p: round(s[p], 4) for p in ("bounds_ms", "morton_ms", "sort_ms", "leaves_ms", "hierarchy_ms", "refit_ms")})

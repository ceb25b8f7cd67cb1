"""cfg2 probe with option sweeps: python tools/probe2.py '{"leaf_size":32}' '{"leaf_size":16}' ..."""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from perf_probe import probe  # noqa: E402
from owlraytracing_b200 import datasets  # noqa: E402

n = 10_000_000
x = datasets.uniform(n, 42)
for arg in sys.argv[1:] or ["{}"]:
    probe("cfg2", x, 10, reps=2, **json.loads(arg))

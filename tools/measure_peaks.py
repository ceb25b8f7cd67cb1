"""L2 / HBM read bandwidth of this B200 from libtrueknn's read probe (BASELINE.md §2 asks for the L2 figure)."""
import json, sys
sys.path.insert(0, ".")
from owlraytracing_b200 import TrueKNN
t = TrueKNN(0)
r = t.measure_bandwidth()
r["how"] = ("tknn_measure_bandwidth: read_probe_kernel (uint4 __ldcg loads, 148*16 blocks x 256 threads), best of 5; "
            "L2: 32 MiB buffer read 64 times per launch; HBM: 2 GiB buffer read once")
print(json.dumps(r))

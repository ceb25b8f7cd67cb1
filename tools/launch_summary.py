"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/<name>_summary.txt
"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("tknn::", "").replace("void ", "").strip()
    v = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, {sys.argv[2] if len(sys.argv) > 2 else ''} (cold-cache, serialised: compare shares)")
print(f"# launches {sum(cnt.values())}, total {total:.1f} us")
for name, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:12.1f} us  {100 * v / total:5.2f}%  n={cnt[name]:4d}  avg {v / cnt[name]:9.1f} us  {name}")

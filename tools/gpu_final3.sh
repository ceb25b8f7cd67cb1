#!/bin/bash
# final call 3: the whole GPU suite on the final build, smoke, the cfg3 line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --workload cfg3 --no-cpu > gpurun_out/r2_bench_cfg3_n1_v12.json 2> gpurun_out/r2_bench_cfg3_n1_v12.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_cfg3_n1_v12.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("cfg3", d["value"], d["ms_per_step"], d["verified"], d["build_ms"], e["value"], e["ms_per_step"], e["indices_only"]["ms_per_step"])
PY

#!/bin/bash
# final call 2: ncu capture of the final onesweep kernel, the bench lines, the reference arm, the launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:onesweep_kernel -s 12 -c 1 -o gpurun_out/r2_build_onesweep_v12 python tools/ncu_build_target.py > gpurun_out/ncu_build_onesweep_v12.log 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench_n1_v12.json 2> gpurun_out/r2_bench_n1_v12.err
timeout 600 python bench.py --workload cfg3 --no-cpu > gpurun_out/r2_bench_cfg3_n1_v12.json 2> gpurun_out/r2_bench_cfg3_n1_v12.err
timeout 300 python bench.py --workload cfg1 --no-cpu > gpurun_out/r2_bench_cfg1_n1_v12.json 2> gpurun_out/r2_bench_cfg1_n1_v12.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n1_v12.json 2> gpurun_out/r2_bench_ref_n1_v12.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_v12.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench_v12.log 2>&1
python - <<'PY'
import json
for f in ("r2_bench_n1_v12", "r2_bench_cfg3_n1_v12", "r2_bench_cfg1_n1_v12", "r2_bench_ref_n1_v12"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        r = d.get("roofline") or {}
        print(f, d.get("value"), d.get("ms_per_step"), d.get("verified"), d.get("build_ms"), e.get("value"), e.get("ms_per_step"),
              (e.get("indices_only") or {}).get("ms_per_step"), r.get("frac"), (r.get("issue_slots") or {}).get("frac"), d.get("clocks"))
    except Exception as ex:
        print(f, "FAILED", ex)
PY

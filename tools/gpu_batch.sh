#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab_ties_cfg3.jsonl
: > $OUT
timeout 300 python tools/ab.py cfg3 libtrueknn.so libtrueknn.so:tie_pruning=2 >> $OUT 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_ties_cfg3.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'rc' in d: print(d); continue
    print(d['workload'], d['lib'], d['opts'], 'search', d['search_ms'], 'kernels', d['kernel_ms'], 'tests/q', d['tests_per_q'], 'nodes/q', d['nodes_per_q'], d['brute_ok'], d['checksum'])
PY

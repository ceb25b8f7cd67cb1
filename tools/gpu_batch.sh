#!/bin/bash
# one gpurun call: parity tests + A/B of the sibling-leaf pair path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
OUT=gpurun_out/r2_ab_pair.jsonl
: > $OUT
timeout 200 python tools/ab.py cfg2 v_pair0.so libtrueknn.so >> $OUT 2>&1
timeout 300 python tools/ab.py cfg3 v_pair0.so libtrueknn.so v_pair2.so >> $OUT 2>&1
timeout 200 python tools/ab.py cfg4 v_pair0.so libtrueknn.so >> $OUT 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_pair.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'rc' in d: print(d); continue
    print(d['workload'], d['lib'], 'search', d['search_ms'], 'kernels', d['kernel_ms'], 'ins/q', d['inserts_per_q'], 'tests/q', d['tests_per_q'], 'nodes/q', d['nodes_per_q'], d['brute_ok'], d['checksum'])
PY

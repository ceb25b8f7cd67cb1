#!/bin/bash
# one gpurun call: A/B of the filter-loop forms of the dense kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab_pairfilter.jsonl
: > $OUT
timeout 300 python tools/ab.py cfg2 libtrueknn.so v_pf8.so v_sf4.so v_pf8sf4.so >> $OUT 2>&1
timeout 300 python tools/ab.py cfg4 libtrueknn.so v_pf8.so v_sf4.so >> $OUT 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_pairfilter.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'rc' in d: print(d); continue
    print(d['workload'], d['lib'], 'search', d['search_ms'], 'kernels', d['kernel_ms'], 'ins/q', d['inserts_per_q'], 'tests/q', d['tests_per_q'], d['brute_ok'], d['checksum'])
PY

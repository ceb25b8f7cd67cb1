#!/bin/bash
# one gpurun call: A/B of the sort variants (results -> gpurun_out/r2_ab_sortkeys4.jsonl)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/r2_ab_sortkeys4.jsonl
: > $OUT
timeout 300 python tools/ab.py cfg2 libtrueknn.so v_lb4.so v_lb2.so v_lb1.so v_ne_lb4.so v_ne_lb2.so v_lb2mb5.so >> $OUT 2>&1
timeout 300 python tools/ab.py cfg4 libtrueknn.so v_lb4.so v_lb2.so v_lb1.so v_ne_lb4.so v_ne_lb2.so v_lb2mb5.so >> $OUT 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_sortkeys4.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'rc' in d: print(d); continue
    print(d['workload'], d['lib'], d['opts'], 'build', d['build_ms'], 'sort', d['build']['sort_ms'], d['brute_ok'], d['checksum'])
PY

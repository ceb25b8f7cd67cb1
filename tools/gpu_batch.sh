#!/bin/bash
# one gpurun call: build tests + A/B of the PTX match_digit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_build.py -x -q 2>&1 | tail -2
OUT=gpurun_out/r2_ab_match.jsonl
: > $OUT
timeout 300 python tools/ab.py cfg2 v_head.so libtrueknn.so libtrueknn.so:sort_mode=1 >> $OUT 2>&1
timeout 300 python tools/ab.py cfg4 v_head.so libtrueknn.so >> $OUT 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2_ab_match.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l[:300]); continue
    if 'rc' in d: print(d); continue
    print(d['workload'], d['lib'], d['opts'], 'search', d['search_ms'], 'build', d['build_ms'], 'sort', d['build']['sort_ms'], d['brute_ok'], d['checksum'])
PY

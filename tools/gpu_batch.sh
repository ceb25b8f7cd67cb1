#!/bin/bash
# one gpurun call: new tests + the indices-only e2e sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_build.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 200 python tools/e2e_sweep.py 4 > gpurun_out/r2_e2e_idxonly.txt 2>&1
timeout 200 python tools/e2e_sweep.py --idx-only 1 2 3 4 >> gpurun_out/r2_e2e_idxonly.txt 2>&1
cat gpurun_out/r2_e2e_idxonly.txt

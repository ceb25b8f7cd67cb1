#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python tools/e2e_sweep.py --cfg3 1 2 3 4 > gpurun_out/r2_e2e_cfg3.txt 2>&1
timeout 300 python tools/e2e_sweep.py --cfg3 --idx-only 1 2 3 >> gpurun_out/r2_e2e_cfg3.txt 2>&1
cat gpurun_out/r2_e2e_cfg3.txt

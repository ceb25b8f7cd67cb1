#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:traverse_kernel -c 1 -o gpurun_out/r2_trav_cfg3_v12 python tools/ncu_target.py cfg3 > gpurun_out/ncu_trav_cfg3_v12.log 2>&1
tail -1 gpurun_out/ncu_trav_cfg3_v12.log; ls -la gpurun_out/r2_trav_cfg3_v12.ncu-rep

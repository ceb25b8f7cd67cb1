#!/bin/bash
# final call 1: the whole GPU suite, smoke, then the ncu captures whose summaries bench.py cites (reps -> gpurun_out/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 python tools/ncu_target.py cfg2 > gpurun_out/ncu_plain_cfg2.log 2>&1 && \
timeout 300 $NCU -k regex:traverse_kernel -c 1 -o gpurun_out/r2_trav_cfg2_v11 python tools/ncu_target.py cfg2 > gpurun_out/ncu_trav_cfg2_v11.log 2>&1
timeout 300 $NCU -k regex:traverse_kernel -c 1 -o gpurun_out/r2_trav_cfg4_v11 python tools/ncu_target.py cfg4 > gpurun_out/ncu_trav_cfg4_v11.log 2>&1
timeout 200 python tools/ncu_build_target.py > gpurun_out/ncu_plain_build.log 2>&1 && \
timeout 300 $NCU -k regex:onesweep_kernel -s 12 -c 1 -o gpurun_out/r2_build_onesweep_v11 python tools/ncu_build_target.py > gpurun_out/ncu_build_onesweep_v11.log 2>&1
timeout 300 $NCU -k regex:morton_kernel -s 2 -c 1 -o gpurun_out/r2_build_morton_v11 python tools/ncu_build_target.py > gpurun_out/ncu_build_morton_v11.log 2>&1
timeout 300 $NCU -k regex:gather_points_kernel -s 2 -c 1 -o gpurun_out/r2_build_gather_v11 python tools/ncu_build_target.py > gpurun_out/ncu_build_gather_v11.log 2>&1
tail -2 gpurun_out/ncu_plain_cfg2.log gpurun_out/ncu_plain_build.log
ls -la gpurun_out/*.ncu-rep

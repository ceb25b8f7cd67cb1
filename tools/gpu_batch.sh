#!/bin/bash
# one gpurun call: tests + e2e sweeps after the mailbox read-back
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_build.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 200 python tools/e2e_sweep.py 3 4 5 6 8 > gpurun_out/r2_e2e_mailbox.txt 2>&1
timeout 200 python tools/e2e_sweep.py --idx-only 1 2 3 >> gpurun_out/r2_e2e_mailbox.txt 2>&1
cat gpurun_out/r2_e2e_mailbox.txt
timeout 100 python tools/ab.py cfg1 libtrueknn.so 2>&1 | cut -c1-400

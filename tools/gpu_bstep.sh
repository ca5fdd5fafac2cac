#!/bin/bash
# First hardware runs of the many-row step kernel: phase-by-phase bisect at two shapes, then the decode-loop parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 180 python tools/bstep_bisect.py --chunks 2 --beam 5 > gpurun_out/bisect_r10.log 2>&1; echo "bisect r10 exit $?"; tail -n 30 gpurun_out/bisect_r10.log | cut -c1-200
timeout 180 python tools/bstep_bisect.py --chunks 16 --beam 5 --d 192 --layers 3 > gpurun_out/bisect_r80.log 2>&1; echo "bisect r80 exit $?"; tail -n 36 gpurun_out/bisect_r80.log | cut -c1-200
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_bstep.py -x > gpurun_out/test_bstep.log 2>&1; echo "test_bstep exit $?"; tail -n 25 gpurun_out/test_bstep.log | cut -c1-300

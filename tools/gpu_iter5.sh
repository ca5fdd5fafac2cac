#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -s -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "large_v3" > gpurun_out/test_large.log 2>&1
echo "large-v3 parity exit $?"; grep -E "large-v3|passed|failed|Error|assert" gpurun_out/test_large.log | cut -c1-300 | tail -12
B2W_DSTEP_PROF=1 timeout 300 python tools/profile_step.py --batch 1 --new-tokens 24 > gpurun_out/dsprof.log 2>&1
grep -E "dstep prof" gpurun_out/dsprof.log | tail -18

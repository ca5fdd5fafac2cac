#!/bin/bash
# persistent decode-step kernel: engine parity tests, then the per-phase profile at large-v3 geometry
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "not search and not sampling" > gpurun_out/test_engine.log 2>&1
echo "tests exit $?"; tail -n 12 gpurun_out/test_engine.log
B2W_DSTEP_PROF=1 timeout 300 python tools/profile_step.py --batch 1 --new-tokens 24 > gpurun_out/dsprof.log 2>&1
echo "prof exit $?"; grep -E "dstep prof|tokens|rror" gpurun_out/dsprof.log | tail -24

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2W_DSTEP_PROF=1 timeout 300 python tools/profile_step.py --batch 1 --new-tokens 24 > gpurun_out/dsprof.log 2>&1
echo "exit $?"; grep -E "dstep prof|tokens" gpurun_out/dsprof.log | tail -14

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "persistent_step_matches and 1-8" > gpurun_out/test_multi_task.log 2>&1
echo "multi-task exit $?"; tail -n 12 gpurun_out/test_multi_task.log | cut -c1-300

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "long_context" > gpurun_out/test_long.log 2>&1
echo "long-context exit $?"; tail -n 12 gpurun_out/test_long.log | cut -c1-300

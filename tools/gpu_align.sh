#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "align or word_timestamps" > gpurun_out/test_align.log 2>&1
echo "align tests exit $?"; tail -n 30 gpurun_out/test_align.log | cut -c1-300

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_bstep.py -k "long_context or other_beam" -rs > gpurun_out/test_bstep_extra.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/test_bstep_extra.log | cut -c1-300
timeout 900 python -m pytest -q -s -m gpu -p no:cacheprovider tests/test_gpu_engine.py -k "large_v3" > gpurun_out/test_large.log 2>&1; echo "large exit $?"; grep "large-v3" gpurun_out/test_large.log | cut -c1-400

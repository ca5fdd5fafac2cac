#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 60 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "(persistent_step_matches and 5-1) or generate_beam5 or teacher_forced" > gpurun_out/test_sanity.log 2>&1
echo "default path sanity exit $?"; tail -n 3 gpurun_out/test_sanity.log | cut -c1-300
B2W_TEST_EXPERIMENTAL=1 timeout 60 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "int8" > gpurun_out/test_int8.log 2>&1
echo "int8 exit $?"; tail -n 12 gpurun_out/test_int8.log | cut -c1-300

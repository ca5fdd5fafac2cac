#!/bin/bash
# Runs the -m gpu test groups in separate processes (a hung kernel in one group cannot block the others).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
run() { # name timeout pytest-args...
  local name=$1; local to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/tests.log
  timeout $to python -m pytest -q -m gpu -p no:cacheprovider "$@" > gpurun_out/test_$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/tests.log
  tail -n 25 gpurun_out/test_$name.log | tee -a gpurun_out/tests.log
}
: > gpurun_out/tests.log
run mel 300 tests/test_gpu_kernels.py -k logmel
run gemm_ref 300 tests/test_gpu_kernels.py -k "gemm_vs_numpy and -1]"
run gemm_tc 300 tests/test_gpu_kernels.py -k "(gemm_vs_numpy and -0]) or gelu"
run attn_ref 300 tests/test_gpu_kernels.py -k "attention and -1]"
run attn_tc 300 tests/test_gpu_kernels.py -k "attention and -0]"
run gemv 600 tests/test_gpu_kernels.py -k skinny
run search 600 tests/test_gpu_engine.py -k "search or sampling"
run engine 900 tests/test_gpu_engine.py -k "not search and not sampling"

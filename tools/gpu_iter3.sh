#!/bin/bash
# decode GEMMs with narrow tiles / split-K: engine parity (both step implementations), GEMM unit tests, batched + single bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x > gpurun_out/test_engine.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/test_engine.log
B2W_DSTEP=0 timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "not search and not sampling" > gpurun_out/test_engine_multi.log 2>&1
echo "tests (multi-kernel step) exit $?"; tail -n 3 gpurun_out/test_engine_multi.log
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "gemm or gelu" > gpurun_out/test_gemm.log 2>&1; echo "gemm tests exit $?"; tail -n 2 gpurun_out/test_gemm.log
timeout 600 python bench.py --workload batched --no-cpu-baseline --steps 2 > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; echo "bench batched exit $?"
python tools/show_bench.py gpurun_out/bench_batched.json 2>&1 | grep -E "^==|value|roofline" | cut -c1-420
timeout 300 python tools/profile_step.py --batch 16 --new-tokens 8 > gpurun_out/prof_b16.log 2>&1; tail -n 3 gpurun_out/prof_b16.log | cut -c1-400

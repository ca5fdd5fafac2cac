#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "(gemm_vs_numpy and -0]) or gelu" > gpurun_out/test_gemm_tc.log 2>&1; echo "gemm_tc exit $?"; tail -n 2 gpurun_out/test_gemm_tc.log
timeout 600 python bench.py > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
python tools/show_bench.py gpurun_out/bench_single.json gpurun_out/bench_reference.json 2>&1 | tail -20
B2W_DSTEP_PROF=1 timeout 300 python tools/profile_step.py --batch 1 --new-tokens 24 > gpurun_out/dsprof.log 2>&1
grep -E "dstep prof" gpurun_out/dsprof.log | tail -20

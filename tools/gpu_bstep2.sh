#!/bin/bash
# bstep at the production shape: remaining parity tests, then per-phase profile + bench lines (batched; single via dstep and via bstep).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_bstep.py > gpurun_out/test_bstep.log 2>&1; echo "test_bstep exit $?"; tail -n 12 gpurun_out/test_bstep.log | cut -c1-300
B2W_DSTEP_PROF=1 timeout 900 python bench.py --workload batched --no-secondary --no-cpu-baseline --steps 3 > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo "bench b16 exit $?"
grep "bstep prof" gpurun_out/bench_b16.err | tail -n 11
B2W_BSTEP=all B2W_DSTEP_PROF=1 timeout 900 python bench.py --workload single --no-cpu-baseline --steps 3 > gpurun_out/bench_b1_bstep.json 2> gpurun_out/bench_b1_bstep.err; echo "bench b1 bstep exit $?"
grep "bstep prof" gpurun_out/bench_b1_bstep.err | tail -n 11
timeout 900 python bench.py --workload single --no-cpu-baseline --steps 3 > gpurun_out/bench_b1_dstep.json 2> gpurun_out/bench_b1_dstep.err; echo "bench b1 dstep exit $?"
python tools/show_bench.py gpurun_out/bench_b16.json gpurun_out/bench_b1_bstep.json gpurun_out/bench_b1_dstep.json 2>&1 | cut -c1-600 | tail -40

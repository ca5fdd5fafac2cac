#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -k "search or sampling" > gpurun_out/test_search.log 2>&1
echo "search exit $?"; tail -n 30 gpurun_out/test_search.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
echo "bench single exit $?"; tail -c 3000 gpurun_out/bench_single.json; tail -n 20 gpurun_out/bench_single.err
timeout 900 python bench.py --steps 2 --warmup 3 --workload batched --no-cpu-baseline > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err
echo "bench batched exit $?"; tail -c 3000 gpurun_out/bench_batched.json; tail -n 20 gpurun_out/bench_batched.err

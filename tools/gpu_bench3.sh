#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x > gpurun_out/test_engine.log 2>&1
echo "engine tests exit $?"; tail -n 25 gpurun_out/test_engine.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
echo "bench single exit $?"; python tools/show_bench.py gpurun_out/bench_single.json; tail -n 3 gpurun_out/bench_single.err

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_int8.py tests/test_gpu_bstep.py -rs > gpurun_out/test_int8.log 2>&1; echo "test int8+bstep exit $?"; tail -n 14 gpurun_out/test_int8.log | cut -c1-300
timeout 900 python bench.py --compute-type int8_float16 --no-secondary --no-cpu-baseline --steps 3 > gpurun_out/bench_b16_int8.json 2> gpurun_out/bench_b16_int8.err; echo "bench int8 exit $?"; tail -3 gpurun_out/bench_b16_int8.err | cut -c1-300
python tools/show_bench.py gpurun_out/bench_b16_int8.json 2>&1 | cut -c1-600 | tail -6

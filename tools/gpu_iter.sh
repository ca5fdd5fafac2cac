#!/bin/bash
# one perf iteration: search + engine parity tests, short bench of both workloads, decode-step phase profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x > gpurun_out/test_engine.log 2>&1
echo "tests exit $?"; tail -n 3 gpurun_out/test_engine.log
timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"
timeout 600 python bench.py --workload batched --no-cpu-baseline --steps 2 > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; echo "bench batched exit $?"
python tools/show_bench.py gpurun_out/bench_single.json gpurun_out/bench_batched.json 2>&1 | grep -E "^==|value|roofline" | cut -c1-420
B2W_DSTEP_PROF=1 timeout 300 python tools/profile_step.py --batch 1 --new-tokens 24 > gpurun_out/dsprof.log 2>&1
grep -E "dstep prof" gpurun_out/dsprof.log | tail -20

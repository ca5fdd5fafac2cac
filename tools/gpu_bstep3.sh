#!/bin/bash
# after a bstep change: bisect at two shapes, decode-loop parity, then the per-phase profile of the batched bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 180 python tools/bstep_bisect.py --chunks 2 --beam 5 > gpurun_out/bisect_r10.log 2>&1; echo "bisect r10 exit $?"; grep -E "worst|self|cross " gpurun_out/bisect_r10.log | cut -c1-150 | tail -8
timeout 180 python tools/bstep_bisect.py --chunks 16 --beam 5 --d 192 --layers 3 > gpurun_out/bisect_r80.log 2>&1; echo "bisect r80 exit $?"; grep -E "worst|self|cross " gpurun_out/bisect_r80.log | cut -c1-150 | tail -8
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_bstep.py > gpurun_out/test_bstep.log 2>&1; echo "test_bstep exit $?"; tail -n 6 gpurun_out/test_bstep.log | cut -c1-300
B2W_DSTEP_PROF=1 timeout 900 python bench.py --workload batched --no-secondary --no-cpu-baseline --steps 3 > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo "bench b16 exit $?"
grep "bstep prof" gpurun_out/bench_b16.err | tail -n 17
python tools/show_bench.py gpurun_out/bench_b16.json 2>&1 | cut -c1-700 | tail -8

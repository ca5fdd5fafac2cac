#!/bin/bash
# engine parity with the mma cross-attention in the multi-kernel step, batched bench, and the N>1 bench path (2 ranks, gloo, one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B2W_DSTEP=0 timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "not search and not sampling" > gpurun_out/test_engine_multi.log 2>&1
echo "tests (multi-kernel step) exit $?"; tail -n 3 gpurun_out/test_engine_multi.log
timeout 600 python bench.py --workload batched --no-cpu-baseline --steps 2 > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; echo "bench batched exit $?"
B2W_XATTN_IMPL=simt timeout 600 python bench.py --workload batched --no-cpu-baseline --steps 2 > gpurun_out/bench_batched_simt.json 2> gpurun_out/bench_batched_simt.err; echo "bench batched simt exit $?"
python tools/show_bench.py gpurun_out/bench_batched.json gpurun_out/bench_batched_simt.json 2>&1 | grep -E "^==|value|roofline" | cut -c1-420
B2W_DIST_BACKEND=gloo timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2rank.json 2> gpurun_out/bench_2rank.err; echo "2-rank bench exit $?"
tail -c 700 gpurun_out/bench_2rank.json; tail -n 3 gpurun_out/bench_2rank.err

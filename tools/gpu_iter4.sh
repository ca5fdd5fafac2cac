#!/bin/bash
# bench with device span timers (single, 2-rank gloo), batched decode launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"
python tools/show_bench.py gpurun_out/bench_single.json 2>&1 | grep -E "^==|value|roofline" | cut -c1-420
B2W_DIST_BACKEND=gloo timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2rank.json 2> gpurun_out/bench_2rank.err; echo "2-rank bench exit $?"
head -c 300 gpurun_out/bench_2rank.json; echo
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_b16.csv python tools/profile_step.py --batch 16 --new-tokens 3 > gpurun_out/ncu_b16.log 2>&1; echo "ncu exit $?"
python tools/ncu_summary.py gpurun_out/launches_b16.csv | head -24

#!/bin/bash
# Round-end GPU pass: full -m gpu suite, smoke(), bench lines (both workloads + reference arm), ncu launch list + one full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# exactly what the driver runs at round end: the whole -m gpu suite in one process
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all_gpu.log 2>&1; echo "pytest -m gpu exit $?"
tail -n 6 gpurun_out/test_all_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "bench single exit $?"
timeout 600 python bench.py --workload batched --no-cpu-baseline --steps 3 > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err; echo "bench batched exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?"
python tools/show_bench.py gpurun_out/bench_single.json gpurun_out/bench_batched.json gpurun_out/bench_reference.json 2>&1 | cut -c1-500 | tail -40
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_single.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k dstep_kernel -s 10 -c 1 -f -o gpurun_out/full_dstep \
  python tools/profile_step.py --batch 1 --new-tokens 16 > gpurun_out/full_dstep.log 2>&1; echo "ncu full exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ds_cross_attn_kernel -s 40 -c 1 -f -o gpurun_out/full_xattn_mma_b16 \
  python tools/profile_step.py --batch 16 --new-tokens 3 > gpurun_out/full_xattn_mma.log 2>&1; echo "ncu full xattn exit $?"
ls -la gpurun_out | tail -6

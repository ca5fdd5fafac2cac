#!/bin/bash
# tests + bench + `ncu --set full` captures of the hot kernels (short timeouts: a stuck ncu must not eat the budget)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest -q -m gpu -p no:cacheprovider tests > gpurun_out/test_all_gpu.log 2>&1
echo "pytest -m gpu exit $?"; tail -n 6 gpurun_out/test_all_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
echo "bench single exit $?"; python tools/show_bench.py gpurun_out/bench_single.json; tail -n 3 gpurun_out/bench_single.err
timeout 600 python bench.py --steps 2 --warmup 3 --workload batched --no-cpu-baseline > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err
echo "bench batched exit $?"; python tools/show_bench.py gpurun_out/bench_batched.json; tail -n 3 gpurun_out/bench_batched.err
export B2W_GRAPH=0
cap() { # name regex skip count batch
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/full_$1 \
     python tools/profile_step.py --batch $5 --new-tokens 2 > gpurun_out/full_$1.log 2>&1
  echo "ncu $1 exit $?"
}
cap skinny_b1 skinny_gemm 400 4 1
cap xattn_b1 dec_cross_attn 40 2 1
cap xattn_b16 dec_cross_attn 40 2 16
cap gemm_b4 gemm_tc 30 4 4
cap attn_b4 attn_tc 4 2 4
cap mel_b4 logmel_power 0 1 4
cap search_b1 search_rows 0 1 1
ls -la gpurun_out/*.ncu-rep

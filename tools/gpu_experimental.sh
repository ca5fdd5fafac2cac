#!/bin/bash
# First run of the opt-in kernels (search_v2.cu, dstep2_kernel): parity tests with short timeouts, then a timing comparison.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B2W_TEST_EXPERIMENTAL=1
timeout 120 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "split_search" > gpurun_out/test_search_v2.log 2>&1
echo "search_v2 exit $?"; tail -n 8 gpurun_out/test_search_v2.log | cut -c1-300
timeout 120 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py -x -k "head_pair" > gpurun_out/test_dstep2.log 2>&1
echo "dstep2 exit $?"; tail -n 8 gpurun_out/test_dstep2.log | cut -c1-300
for cfg in "" "B2W_SEARCH_V2=1" "B2W_DSTEP=2" "B2W_DSTEP=2 B2W_SEARCH_V2=1"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['ms_per_decode_step'])"
done

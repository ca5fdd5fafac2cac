#!/bin/bash
# First hardware run of the wave-pipelined many-row step: parity tests of the kernel, then the A/B of the wave count at large-v3.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_bstep.py tests/test_gpu_int8.py -rs > gpurun_out/test_bstep.log 2>&1; echo "pytest bstep exit $?"
tail -n 15 gpurun_out/test_bstep.log | cut -c1-300
timeout -s KILL 600 python tools/wave_ab.py --waves 1,2,3,4 > gpurun_out/wave_ab.log 2>&1; echo "wave_ab exit $?"; tail -n 8 gpurun_out/wave_ab.log | cut -c1-300
timeout -s KILL 300 python tools/wave_ab.py --waves 2 --repeat 1 --prof > gpurun_out/wave_prof.log 2>&1; echo "wave_prof exit $?"; grep "bstep prof" gpurun_out/wave_prof.log | tail -n 20 | cut -c1-300

#!/bin/bash
# Many-row step on hardware: A/B of wave counts / prefetch gates (tools/wave_ab.py), per-phase device timers, parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -n "$WITH_BASE" ] && [ -f build/libb200whisper_r2base.so ]; then
  B2W_LIBRARY=$PWD/build/libb200whisper_r2base.so timeout -s KILL 300 python tools/wave_ab.py --waves 1 > gpurun_out/wave_base.log 2>&1; echo "base exit $?"; tail -n 2 gpurun_out/wave_base.log | cut -c1-300
fi
timeout -s KILL 600 python tools/wave_ab.py --waves ${WAVES:-1,2} > gpurun_out/wave_ab.log 2>&1; echo "wave_ab exit $?"; tail -n 12 gpurun_out/wave_ab.log | cut -c1-300
for w in ${PROF_WAVES:-1 2}; do
  timeout -s KILL 300 python tools/wave_ab.py --waves $w --repeat 1 --prof > gpurun_out/wave_prof_$w.log 2>&1; echo "wave_prof $w exit $?"; grep "bstep prof" gpurun_out/wave_prof_$w.log | tail -n 18 | cut -c1-300
done
if [ -z "$NO_TESTS" ]; then
  timeout -s KILL 900 python -m pytest -q -m gpu -p no:cacheprovider ${TESTS:-tests/test_gpu_bstep.py tests/test_gpu_int8.py} -rs > gpurun_out/test_bstep.log 2>&1; echo "pytest bstep exit $?"
  tail -n 12 gpurun_out/test_bstep.log | cut -c1-300
fi

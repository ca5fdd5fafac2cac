#!/bin/bash
# Many-row decode step on hardware: this build (and, with WITH_BASE=1, build/libb200whisper_r2base.so) timed by tools/step_ab.py,
# per-phase device timers, parity tests of the step kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -n "$WITH_BASE" ] && [ -f build/libb200whisper_r2base.so ]; then
  B2W_LIBRARY=$PWD/build/libb200whisper_r2base.so timeout -s KILL 300 python tools/step_ab.py > gpurun_out/step_base.log 2>&1; echo "base exit $?"; tail -n 2 gpurun_out/step_base.log | cut -c1-300
fi
timeout -s KILL 600 python tools/step_ab.py --configs "${CONFIGS:-none}" > gpurun_out/step_ab.log 2>&1; echo "step_ab exit $?"; tail -n 12 gpurun_out/step_ab.log | cut -c1-300
if [ -z "$NO_PROF" ]; then
  timeout -s KILL 300 python tools/step_ab.py --repeat 1 --prof > gpurun_out/step_prof.log 2>&1; echo "step_prof exit $?"; grep "bstep prof" gpurun_out/step_prof.log | tail -n 18 | cut -c1-300
fi
if [ -z "$NO_TESTS" ]; then
  timeout -s KILL 900 python -m pytest -q -m gpu -p no:cacheprovider ${TESTS:-tests/test_gpu_bstep.py tests/test_gpu_int8.py} -rs > gpurun_out/test_bstep.log 2>&1; echo "pytest bstep exit $?"
  tail -n 12 gpurun_out/test_bstep.log | cut -c1-300
fi

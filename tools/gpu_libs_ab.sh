#!/bin/bash
# Same-box comparison of several builds of the C ABI (build/lib_*.so, made from other commits) against the in-tree library:
# tools/step_ab.py once per library, then (optionally) the step-kernel parity tests with the in-tree one.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/libs_ab.log
for lib in ${LIBS:-$(ls build/lib_*.so 2>/dev/null)} intree; do
  if [ "$lib" = intree ]; then unset B2W_LIBRARY; else export B2W_LIBRARY=$PWD/$lib; fi
  echo "== $lib" >> gpurun_out/libs_ab.log
  timeout -s KILL 300 python tools/step_ab.py --repeat ${REPEAT:-3} >> gpurun_out/libs_ab.log 2>&1; echo "$lib exit $?"
done
unset B2W_LIBRARY
grep -E "^==|decode" gpurun_out/libs_ab.log | cut -c1-200
if [ -n "$TESTS" ]; then
  timeout -s KILL 900 python -m pytest -q -m gpu -p no:cacheprovider $TESTS -rs > gpurun_out/test_bstep.log 2>&1; echo "pytest exit $?"
  tail -n 8 gpurun_out/test_bstep.log | cut -c1-300
fi
if [ -n "$SINGLE_LIBS" ]; then  # the single-chunk step (dstep_kernel): one chunk, beam 5
  for lib in $SINGLE_LIBS; do
    if [ "$lib" = intree ]; then unset B2W_LIBRARY; else export B2W_LIBRARY=$PWD/$lib; fi
    echo "== single-chunk $lib" >> gpurun_out/libs_ab.log
    timeout -s KILL 300 python tools/step_ab.py --batch 1 --repeat ${REPEAT:-3} >> gpurun_out/libs_ab.log 2>&1; echo "single $lib exit $?"
  done
  unset B2W_LIBRARY
  grep -A1 "^== single" gpurun_out/libs_ab.log | cut -c1-200
fi

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_engine.py > gpurun_out/test_engine.log 2>&1
echo "engine tests exit $?"; tail -n 30 gpurun_out/test_engine.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err
echo "bench single exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_single.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','stages_ms','clocks')}); print(d['e2e']); print(d['roofline']); print(d['roofline_encoder'])
PY
tail -n 5 gpurun_out/bench_single.err
timeout 900 python bench.py --steps 2 --warmup 3 --workload batched --no-cpu-baseline > gpurun_out/bench_batched.json 2> gpurun_out/bench_batched.err
echo "bench batched exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_batched.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','stages_ms','clocks')}); print(d['e2e']); print(d['roofline']); print(d['roofline_encoder'])
PY
tail -n 5 gpurun_out/bench_batched.err

#!/bin/bash
# the two bench arms exactly as the driver runs them, plus the int8 (configs[3]) line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "bench default exit $?"; grep real gpurun_out/bench_default.time
( time timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2> gpurun_out/bench_reference.time; echo "bench reference exit $?"; grep real gpurun_out/bench_reference.time
python tools/show_bench.py gpurun_out/bench_default.json gpurun_out/bench_reference.json 2>&1 | cut -c1-900 | tail -14
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print("single_chunk", {k: d.get("single_chunk", {}).get(k) for k in ("value", "ms_per_step", "e2e")})
print("single roofline", d.get("single_chunk", {}).get("roofline"))
PY

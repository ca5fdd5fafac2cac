#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B2W_GRAPH=0
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b1.csv python tools/profile_step.py --batch 1 --new-tokens 4 > gpurun_out/prof_b1.log 2>&1
echo "ncu b1 exit $?"; tail -n 3 gpurun_out/prof_b1.log
python tools/ncu_summary.py gpurun_out/launches_b1.csv | head -40
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b16.csv python tools/profile_step.py --batch 16 --new-tokens 3 > gpurun_out/prof_b16.log 2>&1
echo "ncu b16 exit $?"; tail -n 3 gpurun_out/prof_b16.log
python tools/ncu_summary.py gpurun_out/launches_b16.csv | head -40

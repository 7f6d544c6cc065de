#!/bin/bash
# ncu pass after a K5-only change: launch list + one full capture of K5 (1 GPU, short command)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v7.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k5_merge -s 1 -c 1 -f -o gpurun_out/prof_k5 $CMD > gpurun_out/ncu_k5.log 2>&1
echo "k5 rc=$?"
python scripts/ncu_summary.py gpurun_out/prof_k5.ncu-rep gpurun_out/ncu_k5_v7_summary.csv 2>&1 | tail -1

#!/bin/bash
# ncu pass for K4 only: launch list + one full capture (1 GPU, short command that already ran clean).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k4_tc -s 1 -c 1 -f -o gpurun_out/prof_k4 $CMD > gpurun_out/ncu_k4.log 2>&1
echo "k4 rc=$?"

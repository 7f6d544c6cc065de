#!/bin/bash
# ncu pass: launch list + full captures of the two hot kernels (1 GPU, short command).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k4_tc -s 1 -c 1 -f -o gpurun_out/prof_k4 $CMD > gpurun_out/ncu_k4.log 2>&1
echo "k4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k1_count -s 2 -c 1 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1
echo "k1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k5_merge -s 1 -c 1 -f -o gpurun_out/prof_k5 $CMD > gpurun_out/ncu_k5.log 2>&1
echo "k5 rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-600
ls -la gpurun_out

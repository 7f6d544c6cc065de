#!/bin/bash
# sanitizer logs, ncu captures (launch list, K4, fused K1, K5) of the headline bench, config-5 sweep
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-graph"
$CMD > gpurun_out/plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_v6.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k4_tc2 -s 2 -c 1 -f -o gpurun_out/prof_k4 $CMD > gpurun_out/ncu_k4.log 2>&1; echo "k4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k1_count_warp -s 2 -c 1 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1; echo "k1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k5_merge -s 2 -c 1 -f -o gpurun_out/prof_k5 $CMD > gpurun_out/ncu_k5.log 2>&1; echo "k5 rc=$?"
for r in k4 k1 k5; do python scripts/ncu_summary.py gpurun_out/prof_$r.ncu-rep gpurun_out/ncu_${r}_v6_summary.csv; done
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py count > gpurun_out/sanitizer_memcheck_count.log 2>&1; echo "memcheck count rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_count.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py count > gpurun_out/sanitizer_racecheck_count.log 2>&1; echo "racecheck count rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_count.log
KB_SAN_N=600 timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py knn > gpurun_out/sanitizer_memcheck_knn.log 2>&1; echo "memcheck knn rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_knn.log
KB_SAN_N=600 timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py knn > gpurun_out/sanitizer_racecheck_knn.log 2>&1; echo "racecheck knn rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_knn.log
timeout 900 python scripts/bench_config5.py 200000 > gpurun_out/bench_config5.jsonl 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; cut -c1-600 gpurun_out/bench_config5.jsonl; tail -3 gpurun_out/bench_config5.err

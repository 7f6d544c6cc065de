#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_knn.py -m gpu -q -x -k "certificate or any_n_neighbors or redundant or exact_side or dense_5120" 2>&1 | tail -15
for sched in 0 1; do
KB_KNN_SCHED=$sched timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/bench_s$sched.log 2> gpurun_out/bench_s$sched.err; echo "bench sched=$sched rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_s$sched.log").read().strip().split("\n")[-1])
    print("bench sched=$sched", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_s$sched.err").read()[-3000:])
PY
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
grep -E "k4_tc|k1_count|k3_norm|k5_merge|k4_init" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | tail -30

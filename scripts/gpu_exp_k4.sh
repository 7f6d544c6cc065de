#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
    print("bench", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"], "K4 TF", round(d["roofline"]["achieved"],1), "K1 GB/s", round(d["roofline_count"]["achieved"],1), "e2e", d.get("e2e",{}).get("ms_per_step"), "launches/step", d["gpu_launches_per_step"])
    for k in ("t2","neighbors15","roofline_count_dense5120","error","traceback"):
        if k in d: print(k, d[k])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench.err").read()[-3000:])
PY
for sched in 0 1; do
  KB_KNN_SCHED=$sched timeout 300 python scripts/exp_k4.py 5p6 2,15 1,2,4,8 2>&1 | grep "K4 "
done
KB_KNN_SCHED=0 KB_KNN_L2_MB=50 timeout 300 python scripts/exp_k4.py 5p6 2 1,8 2>&1 | grep "K4 "
timeout 300 python scripts/exp_k4.py 5+6 15 1,8 2>&1 | grep "K4 "

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_count.py tests/test_gpu_knn.py -m gpu -q -x -k "not redundant" 2>&1 | tail -8
for focc in 4 3; do
KB_K1_FOCC=$focc timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_f$focc.log 2> gpurun_out/bench_f$focc.err; echo "bench focc=$focc rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_f$focc.log").read().strip().split("\n")[-1])
    print("bench focc=$focc", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"], "e2e", d.get("e2e",{}).get("ms_per_step"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_f$focc.err").read()[-3000:])
PY
done

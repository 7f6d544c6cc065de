#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_all.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
    print("bench", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "K4 TF", round(d["roofline"]["achieved"],1), "K1 GB/s", round(d["roofline_count"]["achieved"],1), "frac", round(d["roofline_count"]["frac"],3), "e2e", d.get("e2e",{}).get("ms_per_step"), "launches/step", d["gpu_launches_per_step"])
    for k in ("t2","neighbors15","roofline_count_dense5120","cpu_baseline","error","traceback"):
        if k in d: print(k, json.dumps(d[k])[:400])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench.err").read()[-3000:])
PY
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench reference rc=$?"; cut -c1-300 gpurun_out/bench_ref.log

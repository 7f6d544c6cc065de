#!/bin/bash
# parity suite + headline bench (1 GPU)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1500 python -m pytest tests -m gpu -q -x ${KB_PYTEST_ARGS:-} 2>&1 | tail -25 > gpurun_out/pytest_gpu_all.log; echo "pytest rc=${PIPESTATUS[0]}"
tail -25 gpurun_out/pytest_gpu_all.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench.log").read().strip().split("\n")[-1])
    print("bench", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"], "K4 TF", round(d["roofline"]["achieved"],1), "K1 GB/s", round(d["roofline_count"]["achieved"],1), "e2e", d.get("e2e",{}).get("ms_per_step"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench.err").read()[-2500:])
PY

#!/bin/bash
# quick visit: count parity + headline bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_count.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_count.log; echo "count rc=${PIPESTATUS[0]}"
tail -3 gpurun_out/pytest_count.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 2 --kmer 5+6 --no-cpu-baseline --no-e2e > gpurun_out/bench_5120.log 2> gpurun_out/bench_5120.err; echo "bench 5120 rc=$?"
for f in bench bench_5120; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.log").read().strip().split("\n")[-1])
    print("$f", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"], "K4 TF", round(d["roofline"]["achieved"],1), "K1 GB/s", round(d["roofline_count"]["achieved"],1), "e2e", d.get("e2e",{}).get("ms_per_step"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/$f.err").read()[-1500:])
PY
done

#!/bin/bash
# BASELINE.json configs 3-4 and the north-star target shape on 8 GPUs (dense 5120 columns, n_neighbors=15)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
run() { # name args
  timeout 900 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus 8 --kmer 5+6 --neighbors 15 $2 > gpurun_out/$1.log 2> gpurun_out/$1.err; echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.log").read().strip().split("\n")[-1])
    print("$1", d["config"]["workload"], "| ms/step", round(d["ms_per_step"],2), "| contigs/s", round(d["value"]), "|", {k:round(v,2) for k,v in d["stage_ms"].items()}, "| K4 TF/GPU", round(d["roofline"]["achieved"],1), "| e2e ms", (d.get("e2e") or {}).get("ms_per_step"), "| clocks", d["clocks"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/$1.err").read()[-2500:])
PY
}
run big_500k "--contigs 500000 --steps 3 --warmup 1"
run big_1m "--contigs 1000000 --steps 3 --warmup 1"
run big_2m "--contigs 2000000 --synth S2 --steps 2 --warmup 1 --no-e2e"

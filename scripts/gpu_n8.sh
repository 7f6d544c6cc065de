#!/bin/bash
# 8-GPU visit: sharded == single parity, the 50k bench line with the 500k sub-record
set -u
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29511 scripts/multi_gpu_check.py --contigs 6001 --neighbors 15 > gpurun_out/multi_parity.log 2>&1; echo "parity k15 rc=$?"
grep -h "MULTI_GPU_PARITY\|Error\|error\|timed out" gpurun_out/multi_parity.log | tail -5
timeout 500 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 --big ${KB_BIG:-500000} > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print(round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity_sample",{}).get("ok"), d.get("error"), d.get("traceback"))
    for k in ("config3","north_star_1M"):
        if k in d: print(k, {kk: d[k][kk] for kk in ("ms_per_step","e2e_ms_per_step","k4_tflops_per_gpu","uncertified_rows")}, d[k]["parity_sample"]["ok"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale_$N.err").read()[-2000:])
PY

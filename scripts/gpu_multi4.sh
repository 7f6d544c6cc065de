#!/bin/bash
# 4-GPU visit: sharded == single parity with 4 ranks (three-stream push), benches at N=4 and N=2
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 4 --master-port 29511 scripts/multi_gpu_check.py --contigs 9001 --neighbors 15 > gpurun_out/multi_parity4.log 2>&1; echo "parity N=4 k15 rc=$?"
grep -h "MULTI_GPU_PARITY\|planned passes\|Error\|error" gpurun_out/multi_parity4.log | tail -8
for N in 4 2; do
  timeout 400 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
  echo "bench $N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print("N=$N", round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "parity", d.get("parity_sample",{}).get("ok"), "e2e", round(d["e2e"]["ms_per_step"],3))
    for k in d:
        if k in ("error","traceback"): print(k, d[k])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/scale_$N.err").read()[-3000:])
PY
done

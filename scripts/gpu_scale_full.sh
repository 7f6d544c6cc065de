#!/bin/bash
# one 8-GPU box: the driver's scaling run (N = 8 with the 500k / 1M sharded runs, then 4, 2, 1)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2 1; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
  else
    timeout 900 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
  fi
  echo "bench $N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print("N=$N", round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "parity", d.get("parity_sample",{}).get("ok"), "e2e", round(d["e2e"]["ms_per_step"],3))
    for k in d:
        if k.startswith("sharded") or k in ("config3","north_star_1M","error","traceback"):
            print(k, json.dumps(d[k])[:1400])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/scale_$N.err").read()[-3000:])
PY
done

#!/bin/bash
# one 8-GPU box: N=8 (with a modest sharded 5+6/k=15 run) and N=4 of the headline bench
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4; do
  EXTRA=""; [ "$N" = "8" ] && EXTRA="${KB_BIG8:---big 200000}"
  timeout 600 $TR --nproc-per-node $N --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 3 --no-e2e $EXTRA > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
  echo "bench $N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print("N=$N", round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "parity", d.get("parity_sample",{}).get("ok"), "launches", d["gpu_launches_per_step"])
    for k in d:
        if k.startswith("sharded") or k in ("config3","north_star_1M","error","traceback"):
            print(k, json.dumps(d[k])[:1500])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/scale_$N.err").read()[-3000:])
PY
done

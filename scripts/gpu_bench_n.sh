#!/bin/bash
# N-GPU bench line only (50k contigs, no big sub-records)
set -u
mkdir -p gpurun_out
N=${KB_NGPU:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 --big 0 > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_$N.log").read().strip().split("\n")[-1])
    print(round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity_sample",{}).get("ok"), d.get("error"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale_$N.err").read()[-2000:])
PY

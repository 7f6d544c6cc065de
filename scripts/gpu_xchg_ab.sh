#!/bin/bash
# N-GPU visit: sharded == single parity with the SM push, then the 50k bench with SM push vs copy engines
set -u
mkdir -p gpurun_out
N=${KB_NGPU:-4}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
KB_XCHG_SM=1 timeout 300 $TR --nproc-per-node $N --master-port 29511 scripts/multi_gpu_check.py --contigs 6001 --neighbors 15 > gpurun_out/multi_parity.log 2>&1; echo "parity k15 (SM push) rc=$?"
grep -h "MULTI_GPU_PARITY\|planned passes\|host pass\|Error\|error\|timed out" gpurun_out/multi_parity.log | tail -14
for sm in 1 0; do
  KB_XCHG_SM=$sm timeout 300 $TR --nproc-per-node $N --master-port $((29522+sm)) bench.py --gpus $N --steps 20 --warmup 3 --big 0 > gpurun_out/xchg_sm${sm}_$N.log 2> gpurun_out/xchg_sm${sm}_$N.err
  echo "bench N=$N sm=$sm rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/xchg_sm${sm}_$N.log").read().strip().split("\n")[-1])
    print("sm=$sm", round(d["value"]), "contigs/s", round(d["ms_per_step"],4), "ms", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="how"}, "e2e", round(d["e2e"]["ms_per_step"],3), "parity", d.get("parity_sample",{}).get("ok"), d.get("error"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/xchg_sm${sm}_$N.err").read()[-2000:])
PY
done

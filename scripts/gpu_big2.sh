#!/bin/bash
# 2-GPU reproduction of the large sharded runs (arena of the 1M case)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
KB_XCHG_DEBUG=1 timeout 1200 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --big ${KB_BIG:-1000000} > gpurun_out/big2.log 2> gpurun_out/big2.err
echo "bench rc=$?"; grep "bench big\|karma_b200 rank\|timed out\|never arrived\|Error" gpurun_out/big2.err gpurun_out/big2.log | head -20
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/big2.log").read().strip().split("\n")[-1])
    for k in d:
        if k.startswith("sharded") or k in ("config3","north_star_1M","error","traceback"):
            print(k, json.dumps(d[k])[:1600])
except Exception as e:
    print("failed", e); print(open("gpurun_out/big2.err").read()[-2000:])
PY

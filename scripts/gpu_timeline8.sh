#!/bin/bash
set -u
mkdir -p gpurun_out
N=${KB_NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for sm in 1 0; do
echo "KB_XCHG_SM=$sm"
KB_XCHG_SM=$sm timeout 200 $TR --nproc-per-node $N --master-port $((29541+sm)) scripts/exp_pass_timeline.py 50000 2>&1 | grep "rank " | cut -c1-600
done

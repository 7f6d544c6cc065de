#!/bin/bash
set -u
mkdir -p gpurun_out
N=${KB_NGPU:-2}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29541 scripts/exp_pass_timeline.py 50000 2>&1 | grep "rank " | cut -c1-700

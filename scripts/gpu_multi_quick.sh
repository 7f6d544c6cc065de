#!/bin/bash
# shortest 2-GPU visit: sharded == single parity once, then one short bench at N=2
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29511 scripts/multi_gpu_check.py --contigs 6001 --neighbors 15 > gpurun_out/multi_parity.log 2>&1; echo "parity k15 rc=$?"
grep -h "MULTI_GPU_PARITY\|rows \[" gpurun_out/multi_parity.log | tail -4
timeout 200 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/scale_2.log 2> gpurun_out/scale_2.err
echo "bench 2 rc=$?"; tail -c 1500 gpurun_out/scale_2.log | cut -c1-500; tail -3 gpurun_out/scale_2.err

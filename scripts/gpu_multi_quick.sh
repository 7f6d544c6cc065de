#!/bin/bash
# 2-GPU visit: sharded == single parity (eager/NCCL path and the peer-memory plan), then short benches at N=2
set -u
mkdir -p gpurun_out
N=${KB_NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29511 scripts/multi_gpu_check.py --contigs 6001 --neighbors 15 > gpurun_out/multi_parity.log 2>&1; echo "parity k15 rc=$?"
grep -h "MULTI_GPU_PARITY\|rows \[\|plan pass\|planned passes\|Error\|error" gpurun_out/multi_parity.log | tail -14
timeout 300 $TR --nproc-per-node $N --master-port 29517 scripts/multi_gpu_check.py --contigs 20011 --neighbors 2 > gpurun_out/multi_parity2.log 2>&1; echo "parity k2 rc=$?"
grep -h "MULTI_GPU_PARITY\|planned passes\|Error\|error" gpurun_out/multi_parity2.log | tail -6
timeout 400 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 ${KB_BENCH_ARGS:-} > gpurun_out/scale_$N.log 2> gpurun_out/scale_$N.err
echo "bench $N rc=$?"; tail -c 3000 gpurun_out/scale_$N.log | cut -c1-2500; tail -5 gpurun_out/scale_$N.err

#!/bin/bash
# N-GPU visit: sharded == single parity, then the scaling bench at 1..N.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29511 scripts/multi_gpu_check.py --contigs 6001 --neighbors 15 > gpurun_out/multi_parity.log 2>&1; echo "parity k15 rc=$?"
timeout 600 $TR --nproc-per-node $N --master-port 29512 scripts/multi_gpu_check.py --contigs 3000 --neighbors 2 --synth S2 > gpurun_out/multi_parity2.log 2>&1; echo "parity k2 rc=$?"
timeout 600 $TR --nproc-per-node $N --master-port 29513 scripts/multi_gpu_check.py --contigs 401 --neighbors 4 --synth S3 > gpurun_out/multi_parity3.log 2>&1; echo "parity S3 (flagged rows) rc=$?"
grep -h "MULTI_GPU_PARITY" gpurun_out/multi_parity3.log | tail -2
timeout 600 $TR --nproc-per-node $N --master-port 29514 scripts/multi_gpu_check.py --contigs 6 --neighbors 3 --synth S0 --exotic > gpurun_out/multi_parity4.log 2>&1; echo "parity exotic+compaction rc=$?"
grep -h "MULTI_GPU_PARITY" gpurun_out/multi_parity4.log | tail -2
grep -h "MULTI_GPU_PARITY\|rows \[" gpurun_out/multi_parity.log gpurun_out/multi_parity2.log | tail -20
for g in 1 2 4 8; do
  if [ $g -le $N ]; then
    if [ $g -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$g.log 2> gpurun_out/scale_$g.err
    else
      timeout 600 $TR --nproc-per-node $g --master-port $((29520+g)) bench.py --gpus $g --steps 20 --warmup 3 > gpurun_out/scale_$g.log 2> gpurun_out/scale_$g.err
    fi
    echo "bench $g rc=$?"; tail -c 1500 gpurun_out/scale_$g.log | cut -c1-700; tail -3 gpurun_out/scale_$g.err
  fi
done

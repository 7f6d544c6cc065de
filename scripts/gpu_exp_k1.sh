#!/bin/bash
set -u
mkdir -p gpurun_out
for e in "" "-DKB_K1_EXPERIMENT=1" "-DKB_K1_EXPERIMENT=2"; do
  if [ -z "$e" ]; then python -m karma_b200.build --force > /dev/null 2>&1; else KB_NVCC_EXTRA="$e" python -m karma_b200.build > /dev/null 2>&1; fi
  KB_NVCC_EXTRA="$e" timeout 300 python scripts/exp_k1.py 2>&1 | grep "K1 mode"
done
python -m karma_b200.build --force > /dev/null 2>&1

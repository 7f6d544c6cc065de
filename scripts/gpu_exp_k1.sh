#!/bin/bash
# K1 variants + parity + one ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_count.py -m gpu -q -x 2>&1 | tail -3
for occ in 5 4; do for first in 1024 2048 4096; do
  KB_K1_OCC=$occ KB_K1_FIRST=$first timeout 300 python scripts/exp_k1.py 5p6 2>&1 | grep "K1 "
done; done
timeout 300 python scripts/exp_k1.py 5+6,4+5,7 2>&1 | grep "K1 "
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k1_count_warp -s 3 -c 1 -f -o gpurun_out/prof_k1 python scripts/exp_k1.py 5p6 > gpurun_out/ncu_k1.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
set -u
mkdir -p gpurun_out
KB_NVCC_EXTRA="-DKB_TC_STATS" python -m karma_b200.build > gpurun_out/build_stats.log 2>&1; echo "build rc=$?"
for s in auto 1 2 4 8; do
  if [ $s = auto ]; then timeout 300 python scripts/exp_tc_stats.py 2>&1 | grep "k="; else KB_KNN_SPLITS=$s timeout 300 python scripts/exp_tc_stats.py 2>&1 | grep "k="; fi
done
python -m karma_b200.build --force > /dev/null 2>&1

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -x -k "not redundant and not simt" 2>&1 | tail -4
timeout 300 python scripts/exp_k4.py 5p6 2,15 1,8 2>&1 | grep "K4 "
KB_KNN_SCHED=0 KB_KNN_SPLITS=1 timeout 300 python scripts/exp_k4.py 5p6 2 8 2>&1 | grep "K4 "
KB_KNN_SCHED=1 KB_KNN_SPLITS=1 timeout 300 python scripts/exp_k4.py 5p6 2 8 2>&1 | grep "K4 "
KB_KNN_SCHED=1 KB_KNN_SPLITS=2 timeout 300 python scripts/exp_k4.py 5p6 2 8 2>&1 | grep "K4 "
KB_KNN_SCHED=1 KB_KNN_SPLITS=8 timeout 300 python scripts/exp_k4.py 5p6 2 8 2>&1 | grep "K4 "
timeout 300 python scripts/exp_k4.py 5+6 15 1 2>&1 | grep "K4 "

#!/bin/bash
# experiment: K4 with one 2-CTA MMA per CTA pair (KB_KNN_MMA2=1) against the default; parity first
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
KB_KNN_MMA2=1 timeout 300 python -m pytest tests/test_gpu_knn.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_knn_mma2.log; echo "knn(mma2) rc=${PIPESTATUS[0]}"
tail -6 gpurun_out/pytest_knn_mma2.log
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --no-cpu-baseline --no-e2e > gpurun_out/$1.log 2> gpurun_out/$1.err; echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.log").read().strip().split("\n")[-1])
    print("$1", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stage_ms"].items()}, "K4 TF", round(d["roofline"]["achieved"],1))
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/$1.err").read()[-1500:])
PY
}
run k2_mma2 KB_KNN_MMA2=1 "--steps 20 --warmup 3"
run k2_def KB_KNN_MMA2=0 "--steps 20 --warmup 3"
run k15_mma2 KB_KNN_MMA2=1 "--steps 10 --warmup 3 --neighbors 15"
run d5120_k15_mma2 KB_KNN_MMA2=1 "--steps 5 --warmup 2 --neighbors 15 --kmer 5+6"

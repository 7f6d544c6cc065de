#!/bin/bash
# A/B at 5120 columns: default pair kernel vs the 2-CTA MMA kernel (KB_KNN_MMA2=1), interleaved
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
run() { # name, env, args
  env $2 timeout 300 python bench.py $3 --no-cpu-baseline --no-e2e > gpurun_out/$1.log 2> gpurun_out/$1.err; echo "$1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$1.log").read().strip().split("\n")[-1])
    print("$1", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["stage_ms"].items()}, "K4 TF", round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/$1.err").read()[-1500:])
PY
}
run w_k15_def KB_KNN_MMA2=0 "--steps 12 --warmup 3 --neighbors 15 --kmer 5+6"
run w_k15_mma2 KB_KNN_MMA2=1 "--steps 12 --warmup 3 --neighbors 15 --kmer 5+6"
run w_k15_def_b KB_KNN_MMA2=0 "--steps 12 --warmup 3 --neighbors 15 --kmer 5+6"
run w_k15_mma2_b KB_KNN_MMA2=1 "--steps 12 --warmup 3 --neighbors 15 --kmer 5+6"
run w_k2_def KB_KNN_MMA2=0 "--steps 12 --warmup 3 --kmer 5+6"
run w_k2_mma2 KB_KNN_MMA2=1 "--steps 12 --warmup 3 --kmer 5+6"

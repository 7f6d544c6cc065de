#!/bin/bash
# One GPU-box visit: build check, parity tests (isolated per group), smoke, bench, launch list.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_count.py -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_count.log; echo "count rc=${PIPESTATUS[0]}"
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -k "simt" 2>&1 | tail -25 > gpurun_out/pytest_knn_simt.log; echo "knn simt rc=${PIPESTATUS[0]}"
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -k "not simt" 2>&1 | tail -40 > gpurun_out/pytest_knn_tc.log; echo "knn tc rc=${PIPESTATUS[0]}"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --knn-impl simt --no-cpu-baseline > gpurun_out/bench_simt.log 2> gpurun_out/bench_simt.err; echo "bench simt rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
tail -c 1500 gpurun_out/pytest_count.log; tail -c 3000 gpurun_out/pytest_knn_simt.log; tail -c 3000 gpurun_out/pytest_knn_tc.log; tail -c 600 gpurun_out/smoke.log; tail -c 3000 gpurun_out/bench.log; tail -c 800 gpurun_out/bench.err
# other shapes of BASELINE.json (not the headline): dense 5120 columns, n_neighbors=15
timeout 600 python bench.py --steps 5 --warmup 2 --kmer 5+6 --neighbors 15 --no-cpu-baseline > gpurun_out/bench_5120_k15.log 2> gpurun_out/bench_5120_k15.err; echo "bench 5120/k15 rc=$?"
timeout 600 python bench.py --steps 10 --warmup 2 --neighbors 15 --no-cpu-baseline > gpurun_out/bench_1088_k15.log 2> gpurun_out/bench_1088_k15.err; echo "bench 1088/k15 rc=$?"
for f in bench_5120_k15 bench_1088_k15; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.log").read().strip().split("\n")[-1])
    print("$f", round(d["value"]), "contigs/s", round(d["ms_per_step"],3), "ms", d["stage_ms"], "K4 TF", round(d["roofline"]["achieved"],1), "K1 GB/s", round(d["roofline_count"]["achieved"],1), "e2e", d.get("e2e",{}).get("ms_per_step"))
except Exception as e:
    print("$f failed", e)
PY
done

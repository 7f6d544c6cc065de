#!/bin/bash
# What the driver does at round end (one process per step), plus the ncu pass of the final kernels.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu_all.log; echo "pytest -m gpu rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/pytest_gpu_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench reference rc=$?"
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
if [ "${KB_WITH_NCU:-0}" = "1" ]; then bash scripts/gpu_profile.sh > gpurun_out/prof.txt 2>&1; grep -E "rc=" gpurun_out/prof.txt; fi

#!/bin/bash
# single GPU: the tests that involve flagged rows / the exact pass, then the S3 general path
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q -x -k "exact_side_path or any_n_neighbors or certificate or wide_lists or rerank_reads" 2>&1 | tail -25 > gpurun_out/pytest_flag.log; echo "tests rc=${PIPESTATUS[0]}"
tail -4 gpurun_out/pytest_flag.log
timeout 600 python scripts/bench_config5.py 200000 S3 > gpurun_out/bench_config5_s3.jsonl 2> gpurun_out/bench_config5_s3.err; echo "config5 S3 rc=$?"
python - <<PY
import json
for l in open("gpurun_out/bench_config5_s3.jsonl"):
    d=json.loads(l); print({k:d.get(k) for k in ("ms_per_pass","general_path_ms","general_path_first_call_ms","flagged_rows","general_path_stage_ms")})
PY
tail -3 gpurun_out/bench_config5_s3.err

#!/bin/bash
# rank-4 visit: link tests + secondary bench
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
timeout 400 python -m pytest tests/test_links.py tests/test_readgraph.py tests/test_cabi_load.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_links.log; echo "links rc=${PIPESTATUS[0]}"
tail -25 gpurun_out/pytest_links.log
timeout 300 python scripts/bench_links.py > gpurun_out/bench_links.json 2> gpurun_out/bench_links.err; echo "bench_links rc=$?"; cat gpurun_out/bench_links.json; tail -5 gpurun_out/bench_links.err

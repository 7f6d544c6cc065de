"""Secondary measurement: read-graph edge accumulation from synthetic salmon
equivalence classes.  GPU build (device-resident CSR in, edges on the host out) vs the oracle
port of read_graph.py:61-131 on the host (single thread, like the reference), same input.

    python scripts/bench_readgraph.py [--contigs 200000] [--classes 4000000]
Prints one JSON line (not the driver's bench line).
"""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import read_graph as rg  # noqa: E402
from karma_b200.engine import Engine  # noqa: E402
from oracle import readgraph_oracle as ro  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=200000)
    ap.add_argument("--classes", type=int, default=4000000)
    ap.add_argument("--cpu-classes", type=int, default=400000)
    a = ap.parse_args()
    names, classes = ro.synth_eq_classes(a.contigs, a.classes, seed=11, family=6, max_size=8)
    d = tempfile.mkdtemp()
    path = os.path.join(d, "eq_classes.txt")
    ro.write_eq_file(path, names, classes)
    eng = Engine(0)
    eng.enable_timing(True)
    t0 = time.perf_counter(); parsed = rg.parse(path); t_parse = time.perf_counter() - t0
    for _ in range(2):
        rg.build_edges(eng, parsed)
    eng.stage_ms("readgraph")
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        tot, ea, eb, w = rg.build_edges(eng, parsed)
    torch.cuda.synchronize()
    t_gpu = (time.perf_counter() - t0) / reps
    dev_ms, _ = eng.stage_ms("readgraph")
    pairs = int(sum(len(ids) * (len(ids) - 1) // 2 for f, ids, c in classes if f != "1"))
    # CPU: the reference's loops (oracle port) on a bounded prefix of the classes
    sub = classes[:a.cpu_classes]
    t0 = time.perf_counter(); ro.build(names, sub, ()); t_cpu = time.perf_counter() - t0
    cpu_rate = len(sub) / t_cpu
    print(json.dumps({"metric": "equivalence classes/s (read-graph edge accumulation)", "contigs": a.contigs, "classes": a.classes,
                      "pair_occurrences": pairs, "edges": int(len(ea)), "gpu_ms_device": dev_ms, "gpu_ms_host_to_host": t_gpu * 1e3,
                      "gpu_classes_per_s": a.classes / t_gpu, "parse_s": t_parse,
                      "cpu_port_classes_per_s": cpu_rate, "cpu_sample_classes": len(sub), "speedup_vs_cpu_port": (a.classes / t_gpu) / cpu_rate}))


if __name__ == "__main__":
    main()

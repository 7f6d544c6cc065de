"""Secondary measurement: connection weights between sub-clusters of a synthetic
read graph.  GPU (edge arrays on the host in, pair table on the host out) vs the oracle port of
karma.py:103-118 on a bounded sample of the same graph (single thread, like the reference).

    python scripts/bench_links.py [--nodes 400000] [--edges 3000000] [--group 8]
Prints one JSON line (not the driver's bench line).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import rearrange as rr  # noqa: E402
from karma_b200.engine import Engine  # noqa: E402
from oracle import links_oracle as lo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=400000)
    ap.add_argument("--edges", type=int, default=3000000)
    ap.add_argument("--group", type=int, default=8)
    ap.add_argument("--cpu-groups", type=int, default=1500)
    a = ap.parse_args()
    rng = np.random.default_rng(3)
    n = a.nodes
    # edges mostly between nearby nodes (families), unique, no self loops
    u = rng.integers(0, n, size=a.edges)
    v = np.clip(u + rng.integers(-40, 41, size=a.edges), 0, n - 1)
    far = rng.random(a.edges) < 0.1
    v[far] = rng.integers(0, n, size=int(far.sum()))
    keep = u != v
    lo_, hi_ = np.minimum(u, v)[keep], np.maximum(u, v)[keep]
    key = np.unique(lo_.astype(np.int64) * n + hi_)
    ea, eb = (key // n).astype(np.int32), (key % n).astype(np.int32)
    w = rng.random(len(key)) * 0.5 + 1e-3
    names = ["n%d" % i for i in range(n)]
    index = {nm: i for i, nm in enumerate(names)}
    perm = rng.permutation(n)
    groups = [[names[j] for j in perm[i:i + a.group]] for i in range(0, n, a.group)]
    eng = Engine(0)
    eng.enable_timing(True)
    arrays = (index, ea, eb, w)
    for _ in range(2):
        t = rr.link_table(eng, arrays, groups, cutoff=0.3)
    eng.stage_ms("links")
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        t = rr.link_table(eng, arrays, groups, cutoff=0.3)
    torch.cuda.synchronize()
    t_gpu = (time.perf_counter() - t0) / reps
    dev_ms, _ = eng.stage_ms("links")
    # CPU: the reference's product loops over the pairs of the first `cpu_groups` groups
    sub = groups[:a.cpu_groups]
    members = set(x for g in sub for x in g)
    adj = {}
    for x, y, ww in zip(ea.tolist(), eb.tolist(), w.tolist()):
        if names[x] in members and names[y] in members:
            adj.setdefault(names[x], {})[names[y]] = ww
            adj.setdefault(names[y], {})[names[x]] = ww
    lookup = lo.lookup_dict([sub])
    t0 = time.perf_counter(); ref = lo.connections_between_subclusters(adj, lookup, 0.3); t_cpu = time.perf_counter() - t0
    pairs_cpu = len(sub) * (len(sub) - 1) // 2
    pairs_all = len(groups) * (len(groups) - 1) // 2
    print(json.dumps({"metric": "sub-cluster pairs/s (connection weights, karma.py:103-118)", "nodes": n, "edges": int(len(key)),
                      "groups": len(groups), "linked_pairs": int(len(t["weight"])), "gpu_ms_device": dev_ms,
                      "gpu_ms_host_to_host": t_gpu * 1e3, "gpu_group_pairs_per_s": pairs_all / t_gpu,
                      "cpu_port_group_pairs_per_s": pairs_cpu / t_cpu, "cpu_sample_groups": len(sub), "cpu_sample_s": t_cpu,
                      "cpu_sample_appends": len(ref), "speedup_vs_cpu_port": (pairs_all / t_gpu) / (pairs_cpu / t_cpu)}))


if __name__ == "__main__":
    main()

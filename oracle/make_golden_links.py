"""Freeze golden vectors for the inter-group connection weights from the REAL reference
(authoring container only).

    python -m oracle.make_golden_links

Writes tests/golden/links_golden.json.  Per case: a seeded synthetic eq_classes input, the
sub-cluster nesting handed to create_lookup_dict, and
  * what the unmodified calc_connections_between_mcl_subclusters (karma.py:103-118, compiled
    from /root/reference with `full_graph` bound) returns for several cut-offs, run-length coded,
  * what the real ReadGraph.calc_distance_between_subgraphs (read_graph.py:359-373) returns for
    every linked pair of sub-clusters and for a few overlapping / partly unknown node lists,
    as float hex.
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import links_oracle as lo  # noqa: E402
from oracle import readgraph_oracle as ro  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "links_golden.json")


def nest(names, rng, max_group, max_per_cluster):
    """Random clusters_with_subcluster: shuffled nodes cut into sub-clusters, those into clusters."""
    order = [names[i] for i in rng.permutation(len(names))]
    subs, i = [], 0
    while i < len(order):
        k = int(rng.integers(1, max_group + 1))
        subs.append(order[i:i + k])
        i += k
    clusters, i = [], 0
    while i < len(subs):
        k = int(rng.integers(1, max_per_cluster + 1))
        clusters.append(subs[i:i + k])
        i += k
    return clusters


def rle(pairs):
    out = []
    for a, b in pairs:
        if out and out[-1][0] == a and out[-1][1] == b:
            out[-1][2] += 1
        else:
            out.append([a, b, 1])
    return out


def main():
    specs = [("small", 24, 120, 4, 4, 0.30, 5, 2), ("families", 240, 1500, 4, 6, 0.15, 12, 3),
             ("big_groups", 160, 2500, 8, 8, 0.25, 40, 2), ("singletons", 60, 400, 3, 5, 0.40, 1, 4)]
    cases = []
    d = tempfile.mkdtemp()
    for ci, (name, n, c, fam, ms, p_cross, max_group, per_cluster) in enumerate(specs):
        rng = np.random.default_rng(500 + ci)
        names, classes = ro.synth_eq_classes(n, c, seed=300 + ci, family=fam, max_size=ms, p_cross=p_cross)
        fasta_keys = [">" + x for x in names] + [">fasta_only_%d" % ci]
        path = os.path.join(d, name + ".txt")
        ro.write_eq_file(path, names, classes)
        graph = ro.reference_graph(path, fasta_keys)
        clusters = nest(list(graph.nodes()), rng, max_group, per_cluster)
        lookup = lo.lookup_dict(clusters)
        per_cutoff = {}
        for cutoff in (0, 0.25, 1.5):
            per_cutoff[repr(cutoff)] = rle(lo.reference_connections(graph, lookup, weight_cutoff=cutoff))
        linked = [(a, b) for a, b, _ in per_cutoff["0"]]
        weights = [[a, b, float(graph.calc_distance_between_subgraphs(lookup[a]["mcl_subcluster"], lookup[b]["mcl_subcluster"])).hex()]
                   for a, b in linked]
        # two-list calls: overlapping lists, nodes the graph does not know, an empty list
        nodes = list(graph.nodes())
        lists = []
        for t in range(4):
            la = [nodes[i] for i in rng.choice(len(nodes), size=min(len(nodes), 10 + 5 * t), replace=False)]
            lb = [nodes[i] for i in rng.choice(len(nodes), size=min(len(nodes), 8 + 7 * t), replace=False)]
            if t == 1:
                la.append("not_a_node")
                lb.insert(0, "neither")
            if t == 3:
                lb = []
            lists.append({"nodes_a": la, "nodes_b": lb, "weight": float(graph.calc_distance_between_subgraphs(la, lb)).hex()})
        cases.append({"name": name, "names": names, "classes": [[f, ids, k] for f, ids, k in classes], "fasta_keys": fasta_keys,
                      "clusters_with_subcluster": clusters, "connections": per_cutoff, "pair_weights": weights, "two_lists": lists})
        print(name, graph.number_of_nodes(), "nodes", graph.number_of_edges(), "edges", len(lookup), "sub-clusters",
              len(linked), "linked pairs", {k: sum(x[2] for x in v) for k, v in per_cutoff.items()})
    with open(OUT, "w") as f:
        json.dump({"generator": "oracle/make_golden_links.py",
                   "reference": "/root/reference/karma/karma.py:103-118 (function compiled alone, full_graph bound), "
                                "/root/reference/karma/read_graph.py:359-373", "cases": cases}, f)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()

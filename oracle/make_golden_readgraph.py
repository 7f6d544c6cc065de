"""Freeze read-graph golden vectors from the REAL reference (authoring container only).

    python -m oracle.make_golden_readgraph

Writes tests/golden/readgraph_golden.json: for each seeded synthetic eq_classes file the
input (names, classes, fasta-only keys) and what /root/reference/karma/read_graph.py:61-148
built from it (node order, edges in graph.edges() order with weights as float hex, adjacency
order per node).
"""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import readgraph_oracle as ro  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "readgraph_golden.json")


def main():
    cases = []
    hand = (["c0", "c1", "c2", "c3"], [("2", [0, 1], 5), ("1", [2], 7), ("3", [1, 0, 2], 2), ("2", [3, 2], 0), ("1", [0, 3], 9)])
    specs = [("hand", hand, [">" + x for x in hand[0]] + [">zz", ">aa_first"])]
    for i, (n, c, fam, ms) in enumerate([(12, 40, 4, 4), (40, 300, 4, 6), (64, 900, 8, 8), (30, 200, 3, 5), (100, 1500, 5, 6)]):
        names, classes = ro.synth_eq_classes(n, c, seed=100 + i, family=fam, max_size=ms)
        specs.append(("synth%d" % i, (names, classes), [">" + x for x in names] + [">missing_%d" % i, ">also_missing"]))
    d = tempfile.mkdtemp()
    for name, (names, classes), fasta_keys in specs:
        path = os.path.join(d, name + ".txt")
        ro.write_eq_file(path, names, classes)
        ref = ro.reference_build(path, fasta_keys)
        n = len(names)
        cases.append({"name": name, "names": names, "classes": [[f, ids, c] for f, ids, c in classes], "fasta_keys": fasta_keys,
                      "nodes": ref["nodes"],
                      "edges": [[a, b, float(w).hex()] for a, b, w in ref["edges"]],
                      "adj": {str(k): v for k, v in ref["adj"].items() if k < n}})
        print(name, len(ref["nodes"]), "nodes", len(ref["edges"]), "edges")
    with open(OUT, "w") as f:
        json.dump({"generator": "oracle/make_golden_readgraph.py",
                   "reference": "/root/reference/karma/read_graph.py:61-148 (matplotlib stubbed)", "cases": cases}, f)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()

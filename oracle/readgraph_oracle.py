"""CPU restatement of karma's read-graph construction (TEST INFRASTRUCTURE ONLY).

Restates ``ReadGraph.from_equivalence_classes`` (/root/reference/karma/read_graph.py:61-148)
without networkx: the result is described by the things a networkx consumer can observe --
node order, edge order as ``graph.edges()`` yields it, adjacency order per node, and the
float64 weights.  networkx keeps dict-of-dict adjacency in insertion order, so the same
insertion sequence is replayed on plain dicts.

Rules restated (read_graph.py):
  :75-82   line 1 = number of contigs, line 2 ignored, then the contig names, then one
           equivalence class per line: "eq_size<TAB>id...<TAB>count"
  :86-93   total reads per contig = sum of `count` over every class that lists it
  :99-115  for every class whose FIRST TOKEN is not "1": all itertools.combinations of the
           listed ids (in listed order) get `count` added to their edge
  :117-131 edges are re-inserted in `graph.edges()` order with weight
           ((shared/totA) + (shared/totB)) / 2; edges with shared == 0 are dropped
  :135-146 contigs of the FASTA that salmon did not list are appended as isolated nodes
"""
import itertools
import os
import sys
import tempfile
import types

import numpy as np


def parse_eq_file(path):
    with open(path, "r") as reader:
        n = int(reader.readline())
        reader.readline()
        names = [reader.readline().rstrip("\n") for _ in range(n)]
        classes = []
        for line in reader.readlines():
            first, *ids, count = line.rstrip("\n").split("\t")
            classes.append((first, [int(i) for i in ids], int(count)))
    return names, classes


def build(names, classes, fasta_keys=()):
    """Returns dict(nodes, totals, edges=[(a_idx, b_idx, weight)], adj={node_idx: [nbr_idx...]}).
    Indices refer to `nodes`; isolated FASTA-only nodes are appended in sorted order (the
    reference appends them in set-iteration order, which is not deterministic)."""
    n = len(names)
    totals = [0] * n
    for _, ids, count in classes:
        for i in ids:
            totals[i] += count
    adj = {i: {} for i in range(n)}                      # first graph: shared read counts
    for first, ids, count in classes:
        if first == "1":
            continue
        for a, b in itertools.combinations(ids, 2):
            if b in adj[a]:
                adj[a][b] += count
                adj[b][a] = adj[a][b]
            else:
                adj[a][b] = count
                adj[b][a] = count
    # graph.edges(data=True): nodes in insertion order, neighbours in insertion order, each edge once
    wadj = {i: [] for i in range(n)}
    edges = []
    seen = set()
    for u in range(n):
        for v, shared in adj[u].items():
            if v in seen:
                continue
            if shared == 0:
                continue
            w = ((shared / totals[u]) + (shared / totals[v])) / 2
            edges.append((u, v, w))
            wadj[u].append(v)
            if v != u:
                wadj[v].append(u)
        seen.add(u)
    nodes = list(names)
    have = set(names)
    for k in sorted(set(x.lstrip(">") for x in fasta_keys) - have):
        nodes.append(k)
        wadj[len(nodes) - 1] = []
    return {"nodes": nodes, "totals": totals, "edges": edges, "adj": wadj}


def synth_eq_classes(n_contigs, n_classes, seed=0, family=4, max_size=6, p_cross=0.05, p_zero=0.02, p_single_token=0.01):
    """Synthetic salmon eq_classes: contigs in families of `family`; classes are mostly
    subsets of one family (shared reads between isoforms), sometimes across families."""
    rng = np.random.default_rng(seed)
    names = ["TRINITY_DN%d_c0_g1_i%d" % (i // family, i % family + 1) for i in range(n_contigs)]
    classes = []
    for _ in range(n_classes):
        fam = int(rng.integers(0, max(1, n_contigs // family)))
        size = int(min(rng.integers(1, max_size + 1), family)) if rng.random() > p_cross else int(rng.integers(2, max_size + 1))
        if rng.random() < p_cross:
            ids = rng.choice(n_contigs, size=min(size, n_contigs), replace=False).tolist()
        else:
            base = fam * family
            pool = [i for i in range(base, min(base + family, n_contigs))]
            ids = rng.choice(pool, size=min(size, len(pool)), replace=False).tolist()
        count = 0 if rng.random() < p_zero else int(rng.integers(1, 60))
        first = "1" if (len(ids) == 1 or rng.random() < p_single_token) else str(len(ids))
        classes.append((first, [int(i) for i in ids], count))
    return names, classes


def write_eq_file(path, names, classes):
    with open(path, "w") as f:
        f.write("%d\n%d\n" % (len(names), len(classes)))
        for nm in names:
            f.write(nm + "\n")
        for first, ids, count in classes:
            f.write("\t".join([first] + [str(i) for i in ids] + [str(count)]) + "\n")


# ---------------------------------------------------------------------------
# the real reference (authoring container only)
# ---------------------------------------------------------------------------
REFERENCE_DIR = "/root/reference/karma"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "read_graph.py"))


def load_reference_module():
    """The real read_graph module (matplotlib stubbed; `from logs import logger` resolved from the
    reference directory, karma.log written to a scratch dir)."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="karma_rg_"))
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import read_graph as ref_rg
        import logs as ref_logs
        ref_logs.logger.setLevel(100)
    finally:
        sys.path.remove(REFERENCE_DIR)
        os.chdir(cwd)
    return ref_rg


def reference_graph(eq_path, fasta_keys=()):
    """The real ReadGraph.from_equivalence_classes result (a networkx graph subclass)."""
    return load_reference_module().ReadGraph.from_equivalence_classes(eq_path, {k: "" for k in fasta_keys})


def reference_build(eq_path, fasta_keys=()):
    """Run the real ReadGraph.from_equivalence_classes and describe what a consumer can observe."""
    g = reference_graph(eq_path, fasta_keys)
    nodes = list(g.nodes())
    idx = {k: i for i, k in enumerate(nodes)}
    edges = [(idx[a], idx[b], d["weight"]) for a, b, d in g.edges(data=True)]
    adj = {idx[k]: [idx[x] for x in g.adj[k]] for k in nodes}
    return {"nodes": nodes, "edges": edges, "adj": adj}

"""CPU restatement of karma's inter-group connection weights (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench scripts' CPU-baseline legs may import this.

Restates, on a plain dict-of-dict adjacency {node: {neighbour: weight}} (symmetric):
  ReadGraph.calc_distance_between_subgraphs          /root/reference/karma/read_graph.py:359-373
  calc_connections_between_mcl_subclusters           /root/reference/karma/karma.py:103-118
The second function reads a module-level `full_graph` that karma.py never defines (it is a local
of main(), karma.py:240), so the reference raises NameError as shipped; the restatement takes the
graph as an argument, everything else (iteration order, the float64 running sum, one append per
edge once the sum exceeds the cut-off) is the reference's.

Pinned: tests/golden/links_*.json were produced by oracle/make_golden_links.py from the real
ReadGraph method and from the unmodified source of the karma.py function (compiled from
/root/reference with `full_graph` supplied as a global).
"""
import ast
import itertools
import os

REFERENCE_KARMA = "/root/reference/karma/karma.py"


def adjacency(nodes, edges):
    """edges: iterable of (a, b, weight) over node names -> symmetric dict of dicts."""
    adj = {n: {} for n in nodes}
    for a, b, w in edges:
        adj[a][b] = w
        adj[b][a] = w
    return adj


def distance_between_subgraphs(adj, nodes_a, nodes_b):
    """read_graph.py:359-373."""
    weight = 0
    for node_a, node_b in itertools.product(nodes_a, nodes_b):
        if node_a in adj and node_b in adj[node_a]:
            weight += adj[node_a][node_b]
    return weight


def connections_between_subclusters(adj, mcl_subclusters, weight_cutoff=0):
    """karma.py:103-118 with the graph passed in.  Returns the list of [index_A, index_B]."""
    groups_to_combine = []
    for index_a, index_b in itertools.combinations(mcl_subclusters, 2):
        nodes_a = mcl_subclusters[index_a]["mcl_subcluster"]
        nodes_b = mcl_subclusters[index_b]["mcl_subcluster"]
        weight = 0
        for a, b in itertools.product(nodes_a, nodes_b):
            if a in adj and b in adj[a]:
                weight += adj[a][b]
                if weight > weight_cutoff:
                    groups_to_combine.append([index_a, index_b])
    return groups_to_combine


def pair_table(adj, mcl_subclusters, weight_cutoff=0):
    """Per pair of sub-clusters joined by an edge: (index_A, index_B, weight, edges, appended)."""
    out = []
    for index_a, index_b in itertools.combinations(mcl_subclusters, 2):
        nodes_a = mcl_subclusters[index_a]["mcl_subcluster"]
        nodes_b = mcl_subclusters[index_b]["mcl_subcluster"]
        weight, edges, over = 0, 0, 0
        for a, b in itertools.product(nodes_a, nodes_b):
            if a in adj and b in adj[a]:
                weight += adj[a][b]
                edges += 1
                over += weight > weight_cutoff
        if edges:
            out.append((index_a, index_b, float(weight), edges, over))
    return out


def lookup_dict(clusters_with_subcluster):
    """create_lookup_dict (karma.py:78-100) without the sequence-count assertion."""
    out, index = {}, 0
    for cl_no, cluster in enumerate(clusters_with_subcluster, 1):
        for mcl_cluster in cluster:
            out[index] = {"previous_cluster": cl_no, "mcl_subcluster": mcl_cluster}
            index += 1
    return out


# ---------------------------------------------------------------------------
# the real reference (authoring container only)
# ---------------------------------------------------------------------------
def reference_available():
    return os.path.isfile(REFERENCE_KARMA)


def reference_connections(full_graph, mcl_subclusters, weight_cutoff=0):
    """Compile the UNMODIFIED source of calc_connections_between_mcl_subclusters out of karma.py
    (the module itself cannot be imported: it parses sys.argv and needs salmon/mcl/dammit wrappers)
    and run it with `full_graph` bound as the global the function expects."""
    with open(REFERENCE_KARMA) as f:
        tree = ast.parse(f.read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "calc_connections_between_mcl_subclusters"]
    assert len(fn) == 1
    ns = {"itertools": itertools, "full_graph": full_graph}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REFERENCE_KARMA, "exec"), ns)
    return ns["calc_connections_between_mcl_subclusters"](mcl_subclusters, weight_cutoff=weight_cutoff)

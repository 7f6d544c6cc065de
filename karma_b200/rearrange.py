"""Connection weights between sub-clusters on the GPU (`--rearrange`).

Drop-ins for
  calc_connections_between_mcl_subclusters   /root/reference/karma/karma.py:103-118
  ReadGraph.calc_distance_between_subgraphs  /root/reference/karma/read_graph.py:359-373
whose itertools.product loops probe the graph once per pair of NODES for every pair of groups.
Here every edge of the graph is keyed by the groups of its end points and the weights are added
on the device in the reference's product order (``kb_links_build``), so the float64 sums and
the returned list are identical to the reference's.

karma.py's function reads a global ``full_graph`` that the script never defines (it is a local
of ``main``, karma.py:240); the mirror takes the graph as a keyword argument:

    from karma_b200.rearrange import calc_connections_between_mcl_subclusters
    pairs = calc_connections_between_mcl_subclusters(mcl_subclusters, weight_cutoff=args.THRESHOLD,
                                                     full_graph=full_graph)

There is no CPU fallback: without the CUDA library the call raises.
"""
import itertools
from ctypes import byref, c_int64, c_void_p

import numpy as np
import torch

from ._lib import check, ptr


def graph_arrays(graph):
    """networkx graph -> (node index dict, a int32[E], b int32[E], weight float64[E]); cached on the
    graph object and rebuilt when its node or edge count changed."""
    n, e = graph.number_of_nodes(), graph.number_of_edges()
    cached = getattr(graph, "_kb_link_arrays", None)
    if cached is not None and cached[0] == (n, e):
        return cached[1]
    index = {name: i for i, name in enumerate(graph.nodes())}
    a = np.empty(e, dtype=np.int32)
    b = np.empty(e, dtype=np.int32)
    w = np.empty(e, dtype=np.float64)
    for i, (u, v, weight) in enumerate(graph.edges(data="weight")):
        a[i] = index[u]
        b[i] = index[v]
        w[i] = weight
    out = (index, a, b, w)
    try:
        graph._kb_link_arrays = ((n, e), out)
    except AttributeError:
        pass
    return out


def _roles(index, groups, what):
    """Per node (group, position) from a list of node lists; nodes the graph does not know are skipped
    (has_edge is False for them).  A node listed twice has no single position: rejected."""
    group = np.full(len(index), -1, dtype=np.int32)
    pos = np.full(len(index), -1, dtype=np.int32)
    lens = np.fromiter(map(len, groups), dtype=np.int64, count=len(groups))
    total = int(lens.sum())
    if total == 0:
        return group, pos, 0
    flat = list(itertools.chain.from_iterable(groups))
    node = np.fromiter(map(index.get, flat, itertools.repeat(-1)), dtype=np.int64, count=total)
    starts = np.cumsum(lens) - lens
    gid = np.repeat(np.arange(len(groups), dtype=np.int32), lens)
    p = (np.arange(total, dtype=np.int64) - np.repeat(starts, lens)).astype(np.int32)
    known = node >= 0
    node, gid, p = node[known], gid[known], p[known]
    seen = np.bincount(node, minlength=len(index))
    if seen.size and seen.max() > 1:
        twice = int(np.argmax(seen > 1))
        name = next(k for k, v in index.items() if v == twice)
        raise ValueError("%s: node %r is listed more than once" % (what, name))
    group[node] = gid
    pos[node] = p
    return group, pos, int(lens.max())


def link_table(engine, arrays, row_groups, col_groups=None, n_groups=None, cutoff=0.0):
    """Device part.  Returns dict(group_a, group_b int32[P], weight float64[P], edges, over int64[P]) for the
    pairs of groups joined by at least one edge, ordered by (group_a, group_b)."""
    index, a, b, w = arrays
    rg, rp, m1 = _roles(index, row_groups, "row groups")
    if col_groups is None:
        cg, cp, m2 = rg, rp, m1
    else:
        cg, cp, m2 = _roles(index, col_groups, "column groups")
    if n_groups is None:
        n_groups = len(row_groups)
    dev = engine.device
    out = {"group_a": np.empty(0, np.int32), "group_b": np.empty(0, np.int32), "weight": np.empty(0, np.float64),
           "edges": np.empty(0, np.int64), "over": np.empty(0, np.int64)}
    if len(a) == 0 or len(index) == 0:
        return out
    d_a, d_b, d_w = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev), torch.from_numpy(w).to(dev)
    d_rg, d_rp = torch.from_numpy(rg).to(dev), torch.from_numpy(rp).to(dev)
    d_cg, d_cp = (d_rg, d_rp) if col_groups is None else (torch.from_numpy(cg).to(dev), torch.from_numpy(cp).to(dev))
    engine._bind_stream()
    n_pairs = c_int64()
    check(engine.lib.kb_links_build(engine.ctx, len(a), ptr(d_a), ptr(d_b), ptr(d_w), len(index), ptr(d_rg), ptr(d_rp),
                                    ptr(d_cg), ptr(d_cp), int(n_groups), int(max(m1, m2)), float(cutoff), byref(n_pairs)))
    p = n_pairs.value
    if p:
        out = {"group_a": np.empty(p, np.int32), "group_b": np.empty(p, np.int32), "weight": np.empty(p, np.float64),
               "edges": np.empty(p, np.int64), "over": np.empty(p, np.int64)}
        check(engine.lib.kb_links_fetch(engine.ctx, *(out[k].ctypes.data_as(c_void_p) for k in ("group_a", "group_b", "weight", "edges", "over"))))
    return out


def _engine(engine):
    if engine is None:
        from .engine import Engine
        engine = Engine()
    return engine


def subcluster_link_table(mcl_subclusters, full_graph, weight_cutoff=0, engine=None):
    """(keys, table): `keys` = the dict's keys in iteration order, `table` = link_table over positions in `keys`."""
    keys = list(mcl_subclusters)
    groups = [mcl_subclusters[k]["mcl_subcluster"] for k in keys]
    return keys, link_table(_engine(engine), graph_arrays(full_graph), groups, cutoff=weight_cutoff)


def calc_connections_between_mcl_subclusters(mcl_subclusters, weight_cutoff=0, full_graph=None, engine=None):
    """karma.py:103-118: the list of [index_A, index_B], one entry per joining edge at which the running
    weight of the pair exceeds `weight_cutoff` (the reference appends inside its edge loop)."""
    if full_graph is None:
        raise NameError("name 'full_graph' is not defined (karma.py:114 reads a global the script never sets; pass full_graph=)")
    keys, t = subcluster_link_table(mcl_subclusters, full_graph, weight_cutoff, engine)
    combine = []
    for ga, gb, over in zip(t["group_a"].tolist(), t["group_b"].tolist(), t["over"].tolist()):
        combine.extend([keys[ga], keys[gb]] for _ in range(over))
    return combine


def calc_distance_between_subgraphs(graph, nodes_a, nodes_b, engine=None):
    """read_graph.py:359-373: summed weight of the edges between two node lists (0 when none)."""
    t = link_table(_engine(engine), graph_arrays(graph), [list(nodes_a), []], [[], list(nodes_b)], n_groups=2)
    return float(t["weight"][0]) if len(t["weight"]) else 0

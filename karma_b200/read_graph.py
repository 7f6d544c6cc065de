"""Read graph from salmon equivalence classes on the GPU.

Drop-in for the body of ``ReadGraph.from_equivalence_classes``
(/root/reference/karma/read_graph.py:61-148): the Python loops over every pair OCCURRENCE
(itertools.combinations + has_edge/get_edge_data/add_edge) become a sort / reduce-by-key on
the device (``kb_readgraph_build``); the graph object is then filled with exactly the node
order, edge order, adjacency order and float64 weights the reference produces.

    from karma_b200.read_graph import from_equivalence_classes
    graph = from_equivalence_classes(eq_file, sequences)            # networkx.Graph
    full_graph = ReadGraph(incoming_graph_data=graph)               # karma's own subclass
"""
from ctypes import byref, c_int64, c_void_p

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def parse(eq_file):
    """eq_classes.txt -> dict(names list[str], class_off int64[C+1], ids int32, counts int64[C], skip uint8[C])."""
    lib = _lib.load()
    h = c_void_p()
    n, c, ni, nb = c_int64(), c_int64(), c_int64(), c_int64()
    check(lib.kb_eq_open(str(eq_file).encode(), byref(h), byref(n), byref(c), byref(ni), byref(nb)))
    try:
        names = np.empty(nb.value, dtype=np.uint8)
        name_off = np.empty(n.value + 1, dtype=np.int64)
        class_off = np.empty(c.value + 1, dtype=np.int64)
        ids = np.empty(ni.value, dtype=np.int32)
        counts = np.empty(c.value, dtype=np.int64)
        skip = np.empty(c.value, dtype=np.uint8)
        check(lib.kb_eq_fill(h, names.ctypes.data_as(c_void_p), name_off.ctypes.data_as(c_void_p),
                             class_off.ctypes.data_as(c_void_p), ids.ctypes.data_as(c_void_p),
                             counts.ctypes.data_as(c_void_p), skip.ctypes.data_as(c_void_p)))
    finally:
        lib.kb_eq_close(h)
    raw = names.tobytes().decode("utf-8")
    if len(raw) == len(names):
        no = name_off.tolist()
        name_list = [raw[no[i]:no[i + 1]] for i in range(n.value)]
    else:                                   # multi-byte characters: slice the bytes
        b = names.tobytes()
        no = name_off.tolist()
        name_list = [b[no[i]:no[i + 1]].decode("utf-8") for i in range(n.value)]
    return {"names": name_list, "class_off": class_off, "ids": ids, "counts": counts, "skip": skip}


def build_edges(engine, parsed):
    """Device part.  Returns (totals uint64[n], a int32[E], b int32[E], weight float64[E]) with the
    edges in the reference's graph.edges() order, zero-shared edges already dropped."""
    n = len(parsed["names"])
    c = len(parsed["counts"])
    dev = engine.device
    d_off = torch.from_numpy(parsed["class_off"]).to(dev)
    d_ids = torch.from_numpy(parsed["ids"]).to(dev) if len(parsed["ids"]) else torch.zeros(1, dtype=torch.int32, device=dev)
    d_cnt = torch.from_numpy(parsed["counts"]).to(dev) if c else torch.zeros(1, dtype=torch.int64, device=dev)
    d_skip = torch.from_numpy(parsed["skip"]).to(dev) if c else torch.zeros(1, dtype=torch.uint8, device=dev)
    d_tot = torch.zeros(max(n, 1), dtype=torch.int64, device=dev)
    engine._bind_stream()
    ne = c_int64()
    check(engine.lib.kb_readgraph_build(engine.ctx, n, c, ptr(d_off), ptr(d_ids), ptr(d_cnt), ptr(d_skip), ptr(d_tot), byref(ne)))
    e = ne.value
    a = np.empty(e, dtype=np.int32)
    b = np.empty(e, dtype=np.int32)
    w = np.empty(e, dtype=np.float64)
    sh = np.empty(e, dtype=np.uint64)
    if e:
        check(engine.lib.kb_readgraph_fetch(engine.ctx, a.ctypes.data_as(c_void_p), b.ctypes.data_as(c_void_p),
                                            w.ctypes.data_as(c_void_p), sh.ctypes.data_as(c_void_p)))
    keep = sh != 0                          # read_graph.py:121-122: shared == 0 adds no edge
    return d_tot[:n].cpu().numpy().view(np.uint64), a[keep], b[keep], w[keep]


def from_equivalence_classes(equivalence_class_file, sequences_from_fasta, engine=None, graph_cls=None):
    """Same result as the reference method, as ``graph_cls`` (default networkx.Graph)."""
    import networkx as nx
    if engine is None:
        from .engine import Engine
        engine = Engine()
    parsed = parse(equivalence_class_file)
    names = parsed["names"]
    _, a, b, w = build_edges(engine, parsed)
    g = nx.Graph()
    g.add_nodes_from(names)                                             # read_graph.py:118-119
    g.add_weighted_edges_from(zip((names[i] for i in a.tolist()), (names[i] for i in b.tolist()), w.tolist()))
    have = set(g.nodes())
    for key in sequences_from_fasta.keys():                             # read_graph.py:135-146
        nm = key.lstrip(">")
        if nm not in have:
            g.add_node(nm)
            have.add(nm)
    assert len(g.nodes()) == len(sequences_from_fasta), \
        "The read graph has not enough nodes. Maybe Salmon could couldn't add all contigs to a equivalence class"
    return g if graph_cls is None else graph_cls(incoming_graph_data=g)

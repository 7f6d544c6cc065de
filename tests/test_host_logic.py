"""Host-side logic of the drop-in class and the column dictionary (no GPU)."""
import os
import sys
import types

import numpy as np
import pytest

from karma_b200 import _lib
from karma_b200.engine import (decode_exotic_key, merge_columns, mode_column_names, mode_of, shard_bounds)
from karma_b200.kmer import KmerClustering
from oracle import kmer_oracle as ko


def test_reference_test_suite_passes_on_the_mirror():
    # /root/reference/tests/test_kmer.py:6-8
    assert KmerClustering.is_palindrome("ACGT") is False
    assert KmerClustering.is_palindrome("AAAA") is True


def test_mode_of_matches_reference_errors():
    assert mode_of("5p6") == _lib.KB_MODE_5P6 and mode_of(4) == _lib.KB_MODE_K(4) and mode_of(7) == 23
    with pytest.raises(TypeError):            # kmer.py:84 `len(sequence) - "4p5"`
        mode_of("4p5")
    assert mode_of(9) == _lib.KB_MODE_K(9) and mode_of(16) == _lib.KB_MODE_K(16)      # sorted k-mer keys (k = 8..16)
    with pytest.raises(_lib.KarmaB200Error):
        mode_of(17)                                # a k-mer no longer fits the 128-bit sort key
    with pytest.raises(_lib.KarmaB200Error):
        mode_of(0)


def test_column_names_match_oracle_order():
    assert mode_column_names(_lib.KB_MODE_5P6) == ko.columns_5p6_full()
    assert mode_column_names(_lib.KB_MODE_K(3)) == sorted(mode_column_names(_lib.KB_MODE_K(3)))
    assert len(mode_column_names(_lib.KB_MODE_DENSE_5_6)) == 5120


def _pack_key(s):
    key = 0
    for t, ch in enumerate(s):
        key |= (ord(ch) + 1) << (9 * (6 - t))
    return key


def test_exotic_key_order_is_python_string_order():
    words = ["ACGTN", "ACGTNA", "NACGT", "acgtA", "A\rCGT", "TTTTN", "ACGT\xff", "ACGTa", "ACGTNN", "\x00ACGT"]
    keys = [_pack_key(w) for w in words]
    assert [decode_exotic_key(k) for k in keys] == words
    assert [words[i] for i in np.argsort(np.array(keys, dtype=np.uint64))] == sorted(words)


def test_merge_columns_reproduces_sorted_union(golden):
    case = next(c for c in golden if c["name"] == "KA1")
    names = mode_column_names(_lib.KB_MODE_5P6)
    acgt = [c for c in case["columns"] if set(c) <= set("ACGT")]
    exo = [c for c in case["columns"] if not set(c) <= set("ACGT")]
    present = np.array([n in set(acgt) for n in names])
    cols, colmap, keycol = merge_columns(names, present, exo[::-1])
    assert cols == case["columns"]
    assert all(cols[colmap[i]] == names[i] for i in np.flatnonzero(present)) and (colmap[~present] == -1).all()
    assert [cols[j] for j in keycol] == exo[::-1]


def test_shard_bounds_cover_and_align():
    for n, w in [(10, 3), (8, 2), (7, 8), (50000, 8), (1, 4)]:
        seen = []
        for r in range(w):
            lo, hi, per = shard_bounds(n, w, r)
            assert lo == min(r * per, n) and hi - lo <= per
            seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_mask_list_same_as_reference_semantics():
    seqs = {">a x": "A", ">b": "C", ">c": "G", ">d": "T", ">e": "A"}
    mask = np.array([1, -1, 0, 1, -1])
    labeled, unlabeled = KmerClustering._KmerClustering__mask_list(seqs, mask)
    # reference (kmer.py:30-44) written out
    exp_l, exp_u = [], []
    for i in set(mask):
        pos = [list(seqs.keys())[c].lstrip(">") for c, k in enumerate(mask) if k == i]
        (exp_u if i == -1 else exp_l).append(pos)
    assert labeled == exp_l and unlabeled == exp_u


def test_cluster_file_resume_and_header_fix(tmp_path):
    out = str(tmp_path)
    with open(os.path.join(out, "cluster.txt"), "w") as f:
        f.write("u1 extra\tu2\n")
        f.write("a\tb c\n")
        f.write("d\n")
    k = KmerClustering({}, out, "5p6", 1)
    k.run(2, 10, 0, 42, 2)                       # kmer.py:275-277: skips profile/UMAP/HDBSCAN
    assert k.unlabeled_cluster == [["u1", "u2"]] and k.clusters == [["a", "b"], ["d"]]
    assert k.output_file == f"{out}/cluster.txt" and k.output_eval == f"{out}/eval.txt"


def test_run_flow_with_stubbed_umap_hdbscan(tmp_path, monkeypatch):
    """run() hands the GPU profile + kNN to UMAP (precomputed_knn) and HDBSCAN and writes
    cluster.txt / eval.txt in the reference format; the GPU step itself is stubbed here."""
    seen = {}

    class UMAP:
        def __init__(self, **kw):
            seen["umap"] = kw

        def fit_transform(self, x):
            seen["x"] = x
            return x[:, :2]

    class HDBSCAN:
        def __init__(self, **kw):
            seen["hdb"] = kw

        def fit(self, emb):
            self.labels_ = np.array([0, 0, -1, 1])
            self.probabilities_ = np.array([1.0, 0.5, 0.0, 1.0])
            return self

    monkeypatch.setitem(sys.modules, "umap", types.SimpleNamespace(UMAP=UMAP))
    monkeypatch.setitem(sys.modules, "hdbscan", types.SimpleNamespace(HDBSCAN=HDBSCAN))
    seqs = {">c0 len=5": "ACGTA", ">c1": "ACGTT", ">c2": "GGGGG", ">c3": "TTTTT"}
    k = KmerClustering(seqs, str(tmp_path), "5p6", 2)

    def fake_profile(n_neighbors=None):
        k.knn_indices = np.zeros((4, n_neighbors), dtype=np.int32)
        k.knn_dists = np.zeros((4, n_neighbors), dtype=np.float32)
        return np.eye(4)
    monkeypatch.setattr(k, "_KmerClustering__calc_kmer_profile", fake_profile)
    k.run(neighbors=2, components=10, dist=0, r_state=42, min_cluster_size=2)
    assert seen["umap"]["n_neighbors"] == 2 and seen["umap"]["random_state"] == 42
    idx, dst = seen["umap"]["precomputed_knn"][:2]
    assert idx.shape == (4, 2) and dst.dtype == np.float32 and k.umap_route == "precomputed_knn"
    assert seen["hdb"] == {"min_cluster_size": 2}
    assert k.unlabeled_cluster == [["c2"]] and sorted(k.clusters) == [["c0", "c1"], ["c3"]]
    lines = open(k.output_file).read().split("\n")
    assert lines[0] == "c2"
    ev = open(k.output_eval).read().split("\n")
    assert ev[0].split("\t") == ["kmer_size", "n_neighbors", "n_components", "min_dist", "random_state",
                                 "min_cluster_size", "unlabeled", "no_groups", "mean_probability"]
    assert ev[1].split("\t")[:7] == ["5p6", "2", "10", "0", "42", "2", "1"]


def _fake_umap_039(seen):
    """A module with the signatures of umap-learn 0.3.9 (the version conda/meta.yaml:15 pins): UMAP has no
    precomputed_knn; umap.umap_ offers the stages UMAP.fit runs."""
    class UMAP:
        def __init__(self, n_neighbors=15, n_components=2, metric="euclidean", n_epochs=None, learning_rate=1.0,
                     init="spectral", min_dist=0.1, spread=1.0, set_op_mix_ratio=1.0, local_connectivity=1.0,
                     repulsion_strength=1.0, negative_sample_rate=5, transform_queue_size=4.0, a=None, b=None,
                     random_state=None, metric_kwds=None, angular_rp_forest=False, target_n_neighbors=-1,
                     target_metric="categorical", target_metric_kwds=None, target_weight=0.5, transform_seed=42,
                     verbose=False):
            seen["stock_ctor"] = dict(n_neighbors=n_neighbors, n_components=n_components, min_dist=min_dist, random_state=random_state)

        def fit_transform(self, x):
            seen["stock_fit"] = x
            return np.zeros((x.shape[0], 2))

    def fuzzy_simplicial_set(X, n_neighbors, random_state, metric, metric_kwds={}, knn_indices=None, knn_dists=None,
                             angular=False, set_op_mix_ratio=1.0, local_connectivity=1.0, verbose=False):
        seen["fss"] = dict(X=X, n_neighbors=n_neighbors, metric=metric, knn_indices=knn_indices, knn_dists=knn_dists,
                           random_state=random_state)
        return "GRAPH"

    def simplicial_set_embedding(data, graph, n_components, initial_alpha, a, b, gamma, negative_sample_rate, n_epochs,
                                 init, random_state, metric, metric_kwds, verbose):
        seen["sse"] = dict(data=data, graph=graph, n_components=n_components, initial_alpha=initial_alpha, a=a, b=b, gamma=gamma,
                           negative_sample_rate=negative_sample_rate, n_epochs=n_epochs, init=init, metric=metric)
        return np.full((data.shape[0], n_components), 7.0)

    def find_ab_params(spread, min_dist):
        seen["ab"] = (spread, min_dist)
        return 1.5, 0.9

    return types.SimpleNamespace(UMAP=UMAP, umap_=types.SimpleNamespace(
        fuzzy_simplicial_set=fuzzy_simplicial_set, simplicial_set_embedding=simplicial_set_embedding, find_ab_params=find_ab_params))


def test_umap_039_handoff_through_fuzzy_simplicial_set():
    """umap-learn 0.3.9 has no precomputed_knn argument: the GPU graph goes into fuzzy_simplicial_set and the
    embedding stage is run as UMAP.fit runs it -- the stock neighbour search is never invoked."""
    from karma_b200.kmer import umap_embedding
    seen = {}
    fake = _fake_umap_039(seen)
    rng = np.random.default_rng(3)
    prof = rng.random((30, 12))
    idx = np.argsort(rng.random((30, 30)), axis=1)[:, :5].astype(np.int32)
    dst = np.sort(rng.random((30, 5)).astype(np.float32), axis=1)
    args = {"n_neighbors": 5, "n_components": 3, "min_dist": 0.0, "random_state": 42}
    emb, route = umap_embedding(prof, idx, dst, args, umap_module=fake)
    assert route == "fuzzy_simplicial_set" and "stock_fit" not in seen and "stock_ctor" not in seen
    assert np.array_equal(seen["fss"]["knn_indices"], idx) and seen["fss"]["knn_indices"].dtype == np.int64
    assert np.array_equal(seen["fss"]["knn_dists"], dst) and seen["fss"]["n_neighbors"] == 5 and seen["fss"]["metric"] == "euclidean"
    assert seen["fss"]["X"].dtype == np.float32 and seen["fss"]["X"].shape == (30, 12)            # UMAP.fit's float32 cast
    assert seen["ab"] == (1.0, 0.0)
    assert seen["sse"]["graph"] == "GRAPH" and seen["sse"]["n_components"] == 3 and (seen["sse"]["a"], seen["sse"]["b"]) == (1.5, 0.9)
    assert seen["sse"]["initial_alpha"] == 1.0 and seen["sse"]["gamma"] == 1.0 and seen["sse"]["negative_sample_rate"] == 5
    assert seen["sse"]["n_epochs"] == 0 and seen["sse"]["init"] == "spectral"
    assert emb.shape == (30, 3) and (emb == 7.0).all()
    # the same draws from the seed as UMAP.fit: one RandomState shared by both stages
    assert seen["fss"]["random_state"] is not None
    # no usable graph (n_neighbors >= number of contigs): the stock call of kmer.py:285-290
    seen.clear()
    emb, route = umap_embedding(prof[:4], None, None, {"n_neighbors": 5, "n_components": 2, "min_dist": 0.0, "random_state": 1},
                                umap_module=fake)
    assert route == "stock" and seen["stock_ctor"]["n_neighbors"] == 5 and seen["stock_fit"].shape == (4, 12)


def test_umap_05_rejecting_precomputed_knn_falls_to_the_next_route():
    """umap-learn 0.5.0-0.5.3 reject a graph without an NNDescent index when the parameters are validated:
    that is caught at validation, never out of fit_transform, and the stock call follows (0.5's staged functions
    have a different signature)."""
    from karma_b200.kmer import umap_embedding
    calls = []

    class UMAP:
        def __init__(self, n_neighbors=15, n_components=2, min_dist=0.1, random_state=None, precomputed_knn=(None, None, None)):
            self.precomputed_knn = precomputed_knn
            calls.append(("ctor", precomputed_knn[0] is not None))

        def _validate_parameters(self):
            if self.precomputed_knn[0] is not None:
                raise ValueError("precomputed_knn[2] (knn_search_index) must be an NNDescent object.")

        def fit_transform(self, x):
            calls.append(("fit", self.precomputed_knn[0] is not None))
            if self.precomputed_knn[0] is not None:
                raise TypeError("must not be reached with a rejected graph")
            return np.zeros((x.shape[0], 2))

    fake = types.SimpleNamespace(UMAP=UMAP)
    prof = np.random.default_rng(0).random((10, 4))
    idx = np.zeros((10, 3), dtype=np.int32); dst = np.zeros((10, 3), dtype=np.float32)
    emb, route = umap_embedding(prof, idx, dst, {"n_neighbors": 3, "n_components": 2, "min_dist": 0.0, "random_state": 1}, umap_module=fake)
    assert route == "stock" and calls[-1] == ("fit", False) and ("fit", True) not in calls


def test_run_skips_the_gpu_graph_when_umap_would_truncate_n_neighbors(tmp_path, monkeypatch):
    """n_neighbors >= number of contigs (UMAP truncates it itself) or a non-integer value: profile only, stock call."""
    seen = {}

    class UMAP:
        def __init__(self, **kw):
            seen["umap"] = kw

        def fit_transform(self, x):
            return x[:, :2]

    class HDBSCAN:
        def __init__(self, **kw):
            pass

        def fit(self, emb):
            self.labels_ = np.array([0, 0, -1])
            self.probabilities_ = np.ones(3)
            return self
    monkeypatch.setitem(sys.modules, "umap", types.SimpleNamespace(UMAP=UMAP))
    monkeypatch.setitem(sys.modules, "hdbscan", types.SimpleNamespace(HDBSCAN=HDBSCAN))
    k = KmerClustering({">a": "ACGTA", ">b": "ACGTT", ">c": "GGGGG"}, str(tmp_path), "5p6", 1)

    def fake_profile(n_neighbors=None):
        seen["n_neighbors"] = n_neighbors
        return np.eye(3)
    monkeypatch.setattr(k, "_KmerClustering__calc_kmer_profile", fake_profile)
    k.run(neighbors=30, components=2, dist=0, r_state=1, min_cluster_size=2)
    assert seen["n_neighbors"] is None and "precomputed_knn" not in seen["umap"] and k.umap_route == "stock"


def test_knn_oracle_exact_vs_fp64():
    from oracle import knn_oracle
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 6, (40, 30)).astype(np.uint32)
    counts[7] = counts[3]                                  # exact duplicate -> tie at distance 0
    key_len = rng.integers(5, 30, 40).astype(np.int32)
    key_len[7] = key_len[3]
    prof = counts / key_len[:, None].astype(np.float64)
    i_e, d_e = knn_oracle.knn_exact(counts, key_len, 5)
    i_f, d_f = knn_oracle.knn_fp64(prof, 5)
    assert (i_e[:, 0] == np.arange(40)).all() and (i_f[:, 0] == np.arange(40)).all()
    rep = knn_oracle.check_knn(i_e, np.sqrt(d_e), knn_oracle.d2_fp64(prof))
    assert knn_oracle.parity_ok(rep), rep
    assert i_e[3, 1] == 7 and i_e[7, 1] == 3 and d_e[3, 1] == 0.0
    # a wrong answer must be caught
    bad = i_f.copy()
    bad[0, 1] = i_f[0, -1]
    bad[0, -1] = (set(range(40)) - set(i_f[0].tolist())).pop()
    rep = knn_oracle.check_knn(bad, np.sqrt(knn_oracle.d2_fp64(prof)[np.arange(40)[:, None], bad]), knn_oracle.d2_fp64(prof))
    assert not knn_oracle.parity_ok(rep)


def test_host_chunks_cover_every_row_once():
    """Chunk plan of the host-to-host pass (engine.host_chunks): contiguous, complete, balanced by bases."""
    import numpy as np
    from karma_b200.engine import host_chunks
    rng = np.random.default_rng(3)
    for n, chunks in ((0, 4), (1, 4), (1500, 4), (9000, 1), (9000, 4), (50000, 4), (50000, 7)):
        lens = rng.integers(200, 15000, size=n)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        plan = host_chunks(off, chunks)
        assert plan[0][0] == 0 and plan[-1][1] == n and len(plan) <= max(1, chunks)
        for (lo, hi, b0, b1), nxt in zip(plan, plan[1:] + [None]):
            assert lo <= hi and b0 == off[lo] and b1 == off[hi]
            if nxt is not None:
                assert nxt[0] == hi and hi > lo
        if n >= 2048 and chunks > 1:
            assert len(plan) == chunks
            sizes = [b1 - b0 for _, _, b0, b1 in plan]
            assert max(sizes) - min(sizes) <= 2 * 15000
        if n < 2048:
            assert len(plan) == 1

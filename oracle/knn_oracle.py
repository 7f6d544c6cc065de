"""kNN oracle (TEST INFRASTRUCTURE ONLY) -- parity UNPINNED by the reference.

The neighbour search of the reference lives in umap-learn (pinned 0.3.9 in
/root/reference/conda/meta.yaml:15, unpinned in requirements.txt:1), called at
/root/reference/karma/kmer.py:285-290 with the default euclidean metric; its
source is not under /root/reference and is not installed here (no network), so
no reference-owned vector pins this stage.  Truth is therefore defined by this
file:

* ``knn_fp64``   brute force over the fp64 profile (rows = counts/len(key)),
                 squared euclidean, self included (UMAP's n_neighbors counts the
                 point itself; it comes back as column 0).
* ``knn_exact``  exact-rational distances from the integer counts:
                 d2_ij = sum_c (c_ic*l_j - c_jc*l_i)^2 / (l_i*l_j)^2, compared by
                 cross-multiplication in Python ints (defines truth under ties).
* ``check_knn``  the stated-ties parity rule (1e-5 relative on d2).
* ``knn_umap_small_data``  restatement of umap-learn 0.3.9's neighbour search for fewer than 4096
                 rows, the only branch of that package that is exact: ``fit`` casts the data to
                 float32 (check_array), builds ``sklearn.metrics.pairwise_distances(X, metric=
                 "euclidean")`` and takes, per row, ``argsort(kind="quicksort")[:n_neighbors]``
                 with the distances read back from the matrix [published algorithm, from memory of
                 umap/umap_.py 0.3.x: ``UMAP.fit`` small-data branch, ``nearest_neighbors`` with
                 metric="precomputed", ``fast_knn_indices``].  scikit-learn IS in this image, so this
                 leg runs the third-party function the reference would call; tests/golden/
                 knn_umap_small_golden.npz holds its output (oracle/make_golden_knn.py).  It is a
                 weaker pin than a reference-owned vector: stated, not hidden.
"""
import numpy as np


def d2_fp64(profile, rows=None):
    """Squared euclidean distances of ``rows`` (default all) to all rows,
    direct sum of squared differences in fp64 (no norm expansion)."""
    p = np.asarray(profile, dtype=np.float64)
    idx = np.arange(p.shape[0]) if rows is None else np.asarray(rows)
    out = np.empty((len(idx), p.shape[0]), dtype=np.float64)
    for a, i in enumerate(idx):
        diff = p - p[i]
        out[a] = np.einsum("ij,ij->i", diff, diff)
    return out


def knn_fp64(profile, k, rows=None, d2=None):
    """(idx int64 (R,k), d2 float64 (R,k)): self first, then (distance, index).
    ``d2``: the rows' distances if the caller already has them (d2_fp64(profile, rows))."""
    p = np.asarray(profile, dtype=np.float64)
    ridx = np.arange(p.shape[0]) if rows is None else np.asarray(rows)
    if d2 is None:
        d2 = d2_fp64(p, ridx)
    n = p.shape[0]
    out_i = np.empty((len(ridx), k), dtype=np.int64)
    out_d = np.empty((len(ridx), k), dtype=np.float64)
    for a, i in enumerate(ridx):
        row = d2[a].copy()
        row[i] = -1.0                       # force self to the front
        order = np.lexsort((np.arange(n), row))[:k]
        out_i[a] = order
        out_d[a] = d2[a][order]
    return out_i, out_d


def knn_exact(counts, key_len, k, rows=None):
    """Exact-rational kNN from integer counts (Python ints; small inputs only).
    Order: self, then exact (distance, index)."""
    from fractions import Fraction
    c = [[int(v) for v in r] for r in np.asarray(counts)]
    l = [int(v) for v in key_len]
    n = len(c)
    ridx = range(n) if rows is None else list(rows)
    out_i, out_d = [], []
    for i in ridx:
        ci, li = c[i], l[i]
        keyed = []
        for j in range(n):
            cj, lj = c[j], l[j]
            num = sum((a * lj - b * li) ** 2 for a, b in zip(ci, cj))
            keyed.append((Fraction(num, (li * lj) ** 2), j))
        keyed.sort(key=lambda t: (t[1] != i, t[0], t[1]))
        out_i.append([j for _, j in keyed[:k]])
        out_d.append([float(d) for d, _ in keyed[:k]])
    return np.array(out_i, dtype=np.int64), np.array(out_d, dtype=np.float64)


def knn_umap_small_data(profile, k):
    """(idx int64 (N,k), dist float32 (N,k)) as umap-learn 0.3.9 computes them for N < 4096."""
    from sklearn.metrics import pairwise_distances
    x = np.ascontiguousarray(np.asarray(profile), dtype=np.float32)
    dmat = pairwise_distances(x, metric="euclidean")
    idx = np.argsort(dmat, axis=1, kind="quicksort")[:, :k]
    return idx.astype(np.int64), dmat[np.arange(dmat.shape[0])[:, None], idx].astype(np.float32)


def agreement(idx_a, idx_b):
    """Fraction of rows with the same neighbour SET, and with the same neighbour ORDER."""
    a, b = np.asarray(idx_a), np.asarray(idx_b)
    same_set = np.array([set(x.tolist()) == set(y.tolist()) for x, y in zip(a, b)])
    same_order = (a == b).all(axis=1)
    return float(same_set.mean()), float(same_order.mean())


def check_knn(idx, dist, truth_d2, rows=None, rtol=1e-5, sqrt_dist=True):
    """The stated-ties parity rule.  ``truth_d2``: (R, N) fp64
    squared distances for the checked rows.  Returns a dict of violation counts
    (all zero == parity) plus the fraction of index-identical rows against the
    canonical (self, distance, index) order."""
    idx = np.asarray(idx)
    dist = np.asarray(dist, dtype=np.float64)
    r, k = idx.shape
    n = truth_d2.shape[1]
    ridx = np.arange(r) if rows is None else np.asarray(rows)
    bad_distinct = bad_in = bad_missing = bad_order = bad_dist = bad_self = 0
    identical = 0
    for a in range(r):
        t = truth_d2[a]
        i = ridx[a]
        got = idx[a]
        if len(set(got.tolist())) != k:
            bad_distinct += 1
        if got[0] != i:
            bad_self += 1
        tt = t.copy()
        tt[i] = -1.0
        order = np.lexsort((np.arange(n), tt))
        tau = t[order[k - 1]] if k - 1 < n else np.inf
        tg = t[got]
        tol = rtol * max(tau, 0.0) + 1e-300
        if np.any(tg > tau + tol):
            bad_in += 1
        must = np.flatnonzero(t < tau - tol)
        if len(np.setdiff1d(must, got)):
            bad_missing += 1
        if np.any(np.diff(tg[1:]) < -rtol * np.maximum(tg[1:-1], tg[2:]) - 1e-300):
            bad_order += 1
        ref = np.sqrt(tg) if sqrt_dist else tg
        if not np.allclose(dist[a], ref, rtol=rtol, atol=1e-12):
            bad_dist += 1
        if np.array_equal(got, order[:k]):
            identical += 1
    return {"rows": r, "not_distinct": bad_distinct, "self_not_first": bad_self,
            "beyond_tau": bad_in, "missing_closer": bad_missing,
            "order": bad_order, "dist": bad_dist,
            "index_identical_frac": identical / max(r, 1)}


def parity_ok(report):
    return all(report[k] == 0 for k in
               ("not_distinct", "self_not_first", "beyond_tau", "missing_closer", "order", "dist"))

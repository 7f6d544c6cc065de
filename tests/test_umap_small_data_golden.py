"""kNN lists of umap-learn 0.3.9's exact small-data branch (restated over scikit-learn's own
pairwise_distances, oracle/knn_oracle.knn_umap_small_data; frozen by oracle/make_golden_knn.py)
against the fp64 truth of the oracle (CPU) and against the GPU path (-m gpu)."""
import hashlib
import os

import numpy as np
import pytest

from karma_b200 import synth
from oracle import kmer_oracle as ko
from oracle import knn_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("S1", 1500), ("S2", 1200), ("S0", 600)]
DIST_ATOL = 2e-6          # float32 cast of the profile (UMAP's check_array) against fp64: distances are O(1-10)


@pytest.fixture(scope="module")
def knn_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "knn_umap_small_golden.npz"))


def _case(knn_golden, kind, n):
    asm = synth.make(kind, n)
    assert hashlib.sha1(asm.bases.tobytes()).digest() == knn_golden["%s_%d_sha1" % (kind, n)].tobytes(), "synthetic generator changed"
    return asm


def _compare(knn_golden, kind, n, k, idx, dist, truth_d2, rows):
    """idx/dist: the lists under test for `rows`; truth_d2: fp64 distances of `rows` to all contigs."""
    gi = knn_golden["%s_%d_k%d_idx" % (kind, n, k)].astype(np.int64)[rows]
    gd = knn_golden["%s_%d_k%d_dist" % (kind, n, k)].astype(np.float64)[rows]
    # the golden lists are valid neighbour lists of the fp64 truth once float32-sized ties are allowed
    rep = knn_oracle.check_knn(gi, gd, truth_d2, rows=rows, rtol=1e-4)
    assert rep["not_distinct"] == rep["beyond_tau"] == rep["missing_closer"] == rep["order"] == 0, rep
    # and the lists under test carry the same distances, position by position
    assert np.abs(np.sort(np.asarray(dist, dtype=np.float64), axis=1) - np.sort(gd, axis=1)).max() <= DIST_ATOL
    return knn_oracle.agreement(idx, gi)


@pytest.mark.parametrize("kind,n", [("S1", 1500), ("S0", 600)])
def test_fp64_oracle_agrees_with_umap_small_data_golden(knn_golden, kind, n):
    asm = _case(knn_golden, kind, n)
    _, prof = ko.profile_np(asm.as_dict(), "5p6")
    rows = np.arange(0, n, 6)                                  # a sixth of the rows keeps the CPU suite short
    truth = knn_oracle.d2_fp64(prof, rows)
    for k in (2, 15):
        idx, d2 = knn_oracle.knn_fp64(prof, k, rows, d2=truth)
        same_set, _ = _compare(knn_golden, kind, n, k, idx, np.sqrt(d2), truth, rows)
        assert same_set >= 0.85, (kind, k, same_set)


def test_umap_small_data_restatement_reproduces_golden(knn_golden):
    """scikit-learn is part of the image: the restatement itself is re-run against the frozen lists."""
    asm = _case(knn_golden, "S0", 600)
    _, prof = ko.profile_np(asm.as_dict(), "5p6")
    idx, dist = knn_oracle.knn_umap_small_data(prof, 15)
    gd = knn_golden["S0_600_k15_dist"]
    assert np.abs(np.sort(dist, axis=1).astype(np.float64) - np.sort(gd, axis=1)).max() <= DIST_ATOL
    assert knn_oracle.agreement(idx, knn_golden["S0_600_k15_idx"])[0] >= 0.95      # BLAS threading may flip exact ties


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", CASES)
def test_gpu_knn_agrees_with_umap_small_data_golden(engine, knn_golden, kind, n):
    from karma_b200 import _lib
    from karma_b200.engine import profile_and_knn
    asm = _case(knn_golden, kind, n)
    _, prof = ko.profile_np(asm.as_dict(), "5p6")
    rows = np.arange(0, n, 2)
    truth = knn_oracle.d2_fp64(prof, rows)
    for k in (2, 15):
        res = profile_and_knn(engine, asm.bases, asm.offsets, asm.key_len, "5p6", n_neighbors=k, impl=_lib.KB_KNN_TC)
        same_set, same_order = _compare(knn_golden, kind, n, k, res["knn_idx"][rows], res["knn_dist"][rows], truth, rows)
        print("%s n=%d k=%d: neighbour sets equal to umap-learn's exact branch in %.1f %% of rows, same order in %.1f %%"
              % (kind, n, k, 100 * same_set, 100 * same_order))
        assert same_set >= 0.85, (kind, k, same_set)

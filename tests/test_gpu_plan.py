"""PassPlan: the pre-planned, graph-captured form of the hot path (K1+K3 -> K4 -> K5) and its host-to-host pass must
return exactly what the eager general path (`profile_and_knn`, itself checked against the oracle in test_gpu_knn.py /
test_gpu_count.py) returns, and must report -- not hide -- inputs that break its optimistic assumptions."""
import numpy as np
import pytest
import torch

from karma_b200 import _lib, synth
from karma_b200.engine import PassPlan, profile_and_knn
from oracle import kmer_oracle as ko
from oracle import knn_oracle

pytestmark = pytest.mark.gpu


def _plan(engine, asm, kmer_size, k, graph=True):
    plan = PassPlan(engine, asm.n, int(asm.offsets[-1]), kmer_size, n_neighbors=k, impl=_lib.KB_KNN_TC, graph=graph)
    plan.load(asm.bases, asm.offsets, asm.key_len)
    plan.capture(warmup=1)
    return plan


@pytest.mark.parametrize("graph", [True, False])
@pytest.mark.parametrize("kmer_size,k", [("5p6", 2), ("5p6", 15), ("5+6", 6), (4, 3)])
def test_plan_equals_eager_pass_and_oracle(engine, kmer_size, k, graph):
    asm = synth.make("S1", 2600, seed=11)
    ref = profile_and_knn(engine, asm.bases, asm.offsets, asm.key_len, kmer_size, n_neighbors=k, impl=_lib.KB_KNN_TC)
    plan = _plan(engine, asm, kmer_size, k, graph)
    for _ in range(3):                                   # replays must not depend on what the pass before left behind
        chk = plan.check(plan.run())
        assert chk["ok"] and chk["uncertified"] == 0, chk
        assert plan.profile.cpu().numpy().tobytes() == ref["profile"].tobytes()
        assert np.array_equal(plan.idx.cpu().numpy(), ref["knn_idx"])
        assert np.array_equal(plan.dist.cpu().numpy(), ref["knn_dist"])
    if kmer_size == "5p6":
        cols, prof = ko.profile_np(asm.as_dict(), "5p6")
        assert plan.columns() == cols and plan.profile.cpu().numpy().tobytes() == prof.tobytes()
        rows = np.arange(0, asm.n, 13)
        rep = knn_oracle.check_knn(plan.idx.cpu().numpy()[rows], plan.dist.cpu().numpy()[rows], knn_oracle.d2_fp64(prof, rows), rows=rows)
        assert knn_oracle.parity_ok(rep), rep


@pytest.mark.parametrize("chunks", [1, 4])
def test_host_to_host_pass(engine, chunks):
    asm = synth.make("S1", 9000, seed=5)
    ref = profile_and_knn(engine, asm.bases, asm.offsets, asm.key_len, "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC)
    plan = PassPlan(engine, asm.n, int(asm.offsets[-1]), "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC)
    plan.bind_host(asm.bases, asm.offsets, asm.key_len, chunks=chunks)
    assert len(plan.host["chunks"]) == chunks and plan.host["chunks"][0][0] == 0 and plan.host["chunks"][-1][1] == asm.n
    for _ in range(3):
        plan.host["h_profile"].fill_(-1.0)
        plan.host["h_idx"].fill_(-7)
        plan.d_bases.zero_()                             # nothing may survive on the device from the pass before
        res = plan.run_host()
        assert res["ok"] and res["uncertified"] == 0
        assert res["profile"].tobytes() == ref["profile"].tobytes()
        assert np.array_equal(res["knn_idx"], ref["knn_idx"]) and np.array_equal(res["knn_dist"], ref["knn_dist"])
    # other contigs of the same lengths: rewrite the pinned input in place, no re-planning
    other = synth.make("S1", 9000, seed=6)
    perm = other.bases[:int(asm.offsets[-1])] if other.offsets[-1] >= asm.offsets[-1] else np.resize(other.bases, int(asm.offsets[-1]))
    plan.host["h_bases"][:len(perm)].copy_(torch.from_numpy(np.ascontiguousarray(perm)))
    graph_before = plan.host["graph"]
    res = plan.run_host()
    ref2 = profile_and_knn(engine, perm, asm.offsets, asm.key_len, "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC)
    assert plan.host["graph"] is graph_before
    assert res["ok"] and res["profile"].tobytes() == ref2["profile"].tobytes() and np.array_equal(res["knn_idx"], ref2["knn_idx"])
    # other lengths: bind again (new chunk bounds, new graph)
    asm3 = synth.make("S1", 9000, seed=8)
    if asm3.offsets[-1] <= asm.offsets[-1]:
        plan.bind_host(asm3.bases, asm3.offsets, asm3.key_len, chunks=chunks)
        res = plan.run_host()
        ref3 = profile_and_knn(engine, asm3.bases, asm3.offsets, asm3.key_len, "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC)
        assert res["ok"] and res["profile"].tobytes() == ref3["profile"].tobytes() and np.array_equal(res["knn_idx"], ref3["knn_idx"])


def test_plan_reports_inputs_it_cannot_serve(engine):
    """Non-ACGT bytes, a contig shorter than k and a missing column all fail the validation words (the caller then takes
    the general path); nothing is silently wrong."""
    asm = synth.make("S1", 1500, seed=3)
    plan = _plan(engine, asm, "5p6", 2)
    assert plan.check(plan.run())["ok"]
    bases = asm.bases.copy()
    bases[int(asm.offsets[700]) + 40] = ord("N")
    plan.load(bases, asm.offsets, asm.key_len)
    chk = plan.check(plan.run())
    assert not chk["ok"] and chk["exotic"]
    # a contig shorter than k: all-zero row (kmer.py:250-258)
    off = asm.offsets.copy()
    short = synth.Assembly(np.concatenate([asm.bases[:int(off[1499])], asm.bases[int(off[1499]):int(off[1499]) + 3]]),
                           np.concatenate([off[:1500], [off[1499] + 3]]), asm.gene, asm.iso)
    p2 = _plan(engine, short, "5p6", 2)
    chk = p2.check(p2.run())
    assert not chk["ok"] and (chk["flags_or"] & 4)
    # homopolymers only: most columns never occur, kmer.py's dictionary would be smaller
    n = 600
    hb = np.frombuffer((b"A" * 300) * n, dtype=np.uint8).copy()
    ho = np.arange(n + 1, dtype=np.int64) * 300
    p3 = PassPlan(engine, n, int(ho[-1]), "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC)
    p3.load(hb, ho, np.full(n, 9, dtype=np.int32))
    p3.capture(warmup=1)
    chk = p3.check(p3.run())
    assert not chk["ok"] and not chk["complete"]

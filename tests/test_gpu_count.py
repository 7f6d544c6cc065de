"""K1/K1x/K2/K3 parity on a B200: counts bit-exact against the oracle, the fp64 profile
bit-exact against the golden vectors produced by the real kmer.py."""
import numpy as np
import pytest
import torch

from karma_b200 import synth
from karma_b200.engine import mode_of
from oracle import kmer_oracle as ko

pytestmark = pytest.mark.gpu


def _count(engine, bases, offsets, mode_name):
    mode = mode_of(mode_name)
    key_len = np.ones(len(offsets) - 1, dtype=np.int32)
    d_bases, d_offsets, _ = engine.upload(bases, offsets, key_len)
    counts, exotic, presence = engine.count(d_bases, d_offsets, len(offsets) - 1, mode)
    torch.cuda.synchronize()
    pres = presence.cpu().numpy() != 0
    ex = exotic.cpu().numpy().view(np.uint32)
    assert pres[-1] == (ex.sum() > 0), "presence[D] must flag non-ACGT windows"
    return counts.cpu().numpy().view(np.uint32), ex, pres[:-1]


@pytest.mark.parametrize("mode", ["5p6", "5+6", "4+5", 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("kind", ["S0", "S1"])
def test_counts_bit_exact(engine, mode, kind):
    asm = synth.make(kind, 257)
    got, exotic, presence = _count(engine, asm.bases, asm.offsets, mode)
    want, wexo = ko.counts_mode(asm.bases, asm.offsets, mode)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(exotic, wexo.astype(np.uint32))
    assert np.array_equal(presence, want.any(0))


def _pack(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    return np.frombuffer("".join(seqs).encode("latin-1"), dtype=np.uint8), offsets


@pytest.mark.parametrize("mode", ["5p6", "5+6", 4, 7])
def test_counts_edge_cases(engine, mode):
    rng = np.random.default_rng(5)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    seqs = ["", "A", "ACGT", "ACGTA", "ACGTAC", "ACGTACG", "A" * 700, "ACGT" * 50, rnd(15), rnd(16), rnd(17),
            rnd(31), rnd(33), rnd(2047), rnd(2048), rnd(2049), rnd(4097), "", rnd(9), "CGTTGC" * 7, "GAGGAG",
            "NNNNNNNNNN", "ACGTNACGTACGTTGCAACGTnACGT", rnd(100) + "N" + rnd(100), "acgtacgtacgt", rnd(5000)]
    bases, offsets = _pack(seqs)
    got, exotic, presence = _count(engine, bases, offsets, mode)
    want, wexo = ko.counts_mode(bases, offsets, mode)
    assert np.array_equal(got, want)
    assert np.array_equal(exotic, wexo.astype(np.uint32))
    assert np.array_equal(presence, want.any(0))


@pytest.mark.parametrize("mode", ["5p6", "5+6", 7])
def test_long_contig_split_path(engine, mode):
    """Contigs above the 64 kb threshold are tiled over CTAs and merged."""
    rng = np.random.default_rng(9)
    lens = [300, 70000, 65536, 65537, 1200, 200001, 16385 * 4 + 3, 50]
    seqs = ["".join("ACGT"[i] for i in rng.integers(0, 4, n)) for n in lens]
    seqs[5] = seqs[5][:1000] + "N" + seqs[5][1001:150000] + "a" + seqs[5][150001:]
    bases, offsets = _pack(seqs)
    got, exotic, presence = _count(engine, bases, offsets, mode)
    n_long, ex_total = engine.count_stats()
    want, wexo = ko.counts_mode(bases, offsets, mode)
    L = np.array(lens)
    assert n_long in (int((L > 65536).sum()), int((L > 16384).sum())) and n_long >= 4   # CTA-per-contig modes split above 16 kb when contigs are few
    assert np.array_equal(got, want)
    assert np.array_equal(exotic, wexo.astype(np.uint32)) and ex_total == int(wexo.sum())
    assert np.array_equal(presence, want.any(0))


def test_profile_bit_exact_against_reference_golden(engine, golden):
    """Through the reference-facing class: columns, float64 matrix and exit(1) behaviour
    must equal what /root/reference/karma/kmer.py produced (tests/golden)."""
    from karma_b200.kmer import KmerClustering
    for c in golden:
        seqs = dict(zip(c["keys"], c["seqs"]))
        k = KmerClustering(seqs, "/tmp", c["kmer_size"], 2)
        k._engine = engine
        if "exit" in c:
            with pytest.raises(SystemExit) as e:
                k._KmerClustering__calc_kmer_profile()
            assert e.value.code == c["exit"]
            continue
        mat = k._KmerClustering__calc_kmer_profile()
        cols = sorted(k.kmers, key=k.kmers.get)
        assert cols == c["columns"], c["name"]
        assert mat.dtype == np.float64 and list(mat.shape) == c["shape"], c["name"]
        assert np.ascontiguousarray(mat).tobytes().hex() == c["matrix_hex"], c["name"]


@pytest.mark.parametrize("kmer_size", ["5p6", 4, 6])
def test_profile_bit_exact_against_oracle_with_exotic_bytes(engine, kmer_size):
    from karma_b200.kmer import KmerClustering
    rng = np.random.default_rng(21)
    seqs = {}
    for i in range(40):
        alpha = "ACGT" if i % 3 else "ACGTNacgtRY\r"
        L = int(rng.integers(8, 600))
        seqs[">c%d_%s" % (i, "x" * int(rng.integers(0, 20)))] = "".join(alpha[j] for j in rng.integers(0, len(alpha), L))
    k = KmerClustering(seqs, "/tmp", kmer_size, 2)
    k._engine = engine
    mat = k._KmerClustering__calc_kmer_profile()
    cols, want = ko.profile_np(seqs, kmer_size)
    assert sorted(k.kmers, key=k.kmers.get) == cols
    assert mat.tobytes() == want.tobytes()


def test_full_size_properties_50k(engine):
    """BASELINE config 2 size (50k contigs): size-independent properties of the count
    matrix -- every row of the 5-mer block sums to L-4, the 6-mer block to L-5, a
    checksum of checksums against numpy, and linearity (counts of a concatenated pair
    differ from the sum of the parts only by the junction windows)."""
    asm = synth.s1_families(50000)
    L = np.diff(asm.offsets)
    got, exotic, presence = _count(engine, asm.bases, asm.offsets, "5+6")
    assert exotic.sum() == 0
    assert np.array_equal(got[:, :1024].sum(1, dtype=np.int64), L - 4)
    assert np.array_equal(got[:, 1024:].sum(1, dtype=np.int64), L - 5)
    sub = np.arange(0, 50000, 97)
    want, _ = ko.counts_mode(*_slice(asm, sub), "5+6")
    assert np.array_equal(got[sub], want)
    # marginalising the 6-mer block over its last base gives the 5-mer block minus the last window
    marg = got[:, 1024:].reshape(50000, 1024, 4).sum(2)
    diff = got[:, :1024].astype(np.int64) - marg
    assert (diff >= 0).all() and np.array_equal(diff.sum(1), np.ones(50000, dtype=np.int64))
    g5p6, _, _ = _count(engine, asm.bases, asm.offsets, "5p6")
    full = ko.columns_5p6_full()
    is5 = np.array([len(c) == 5 for c in full])
    assert np.array_equal(g5p6[:, is5], got[:, :1024])


def _slice(asm, rows):
    seqs = [asm.bases[asm.offsets[r]:asm.offsets[r + 1]] for r in rows]
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    return np.concatenate(seqs), off


@pytest.mark.parametrize("mode", ["5p6", "5+6", "4+5", 1, 3, 5, 6, 7])
def test_fused_count_profile_equals_count_then_normalise(engine, mode):
    """kb_count_profile (K1+K3 in one kernel, no u32 rows in HBM) against kb_count + kb_normalise: profile, operand and
    row records bit for bit, incl. empty / short / 70 kb / 200 kb contigs (no split path in the fused kernel), padding
    rows, the exotic tallies and the presence summary."""
    rng = np.random.default_rng(13)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    seqs = [rnd(int(x)) for x in rng.integers(200, 6000, 300)] + ["", "ACG", rnd(70000), rnd(200001), "A" * 5000, rnd(4096), rnd(4097),
                                                                  rnd(300) + "N" + rnd(200), "acgt" + rnd(50)]
    bases, offsets = _pack(seqs)
    n = len(seqs)
    key_len = rng.integers(3, 60, n).astype(np.int32)
    m = mode_of(mode)
    d_bases, d_offsets, d_len = engine.upload(bases, offsets, key_len)
    counts, exotic, presence = engine.count(d_bases, d_offsets, n, m)
    cols = counts.shape[1]
    flags_a = torch.zeros(1, dtype=torch.int32, device="cuda")
    prof_a, op_a, meta_a = engine.normalise(counts, cols, d_len, rows_alloc=n + 3, flags_or=flags_a)
    prof_b = torch.empty_like(prof_a); op_b = torch.full_like(op_a, 7.0); meta_b = torch.full_like(meta_a, -5)
    exo_b = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    pres_b = torch.zeros(cols + 1, dtype=torch.int32, device="cuda")
    flags_b = torch.zeros(1, dtype=torch.int32, device="cuda")
    engine.count_profile(d_bases, d_offsets, d_len, 0, n, m, prof_b, op_b, meta_b, exo_b, pres_b, flags_b, rows_alloc=n + 3)
    torch.cuda.synchronize()
    assert torch.equal(prof_a.view(torch.int64), prof_b.view(torch.int64)), "profile rows differ"
    assert torch.equal(op_a.view(torch.int16), op_b.view(torch.int16)), "operand rows differ"
    assert torch.equal(meta_a, meta_b), "row records differ"
    assert torch.equal(exotic, exo_b) and int(flags_a) == int(flags_b)
    pa, pb = presence.cpu().numpy(), pres_b.cpu().numpy()
    assert (pa[-1] != 0) == bool(pb[-1] & 1)
    complete = bool(pb[-1] & 2) or bool((pb[:-1] != 0).all())
    assert complete == bool((pa[:-1] != 0).all())
    if not (pb[-1] & 2):
        assert np.array_equal(pa[:-1] != 0, pb[:-1] != 0)


@pytest.mark.parametrize("k", [8, 9, 12, 16])
def test_integer_k_beyond_7_sorted_counting(engine, k):
    """kmer.py:83-85 takes any integer -k.  Beyond 7 the windows are counted through sorted 128-bit keys: columns =
    the observed k-mers in sorted() order, any byte a k-mer character -- bit-exact against the pure-Python port."""
    from karma_b200.engine import profile_and_knn
    rng = np.random.default_rng(40 + k)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    fam = rnd(400)
    seqs = {">c%d" % i: rnd(int(x)) for i, x in enumerate(rng.integers(k, 700, 30))}
    seqs[">rep a"] = "ACGT" * 40
    seqs[">withN"] = rnd(60) + "N" + rnd(50) + "nn" + rnd(30)
    seqs[">f1"] = fam
    seqs[">f2 x"] = fam[:200] + "T" + fam[201:]
    seqs[">exact_k"] = rnd(k)
    bases, offsets, key_len = ko.pack(seqs)
    res = profile_and_knn(engine, bases, offsets, key_len, k, n_neighbors=3)
    cols, want = ko.profile_port(seqs, k)
    assert res["columns"] == cols
    assert res["profile"].shape == want.shape and res["profile"].tobytes() == want.tobytes()
    from oracle import knn_oracle
    rep = knn_oracle.check_knn(res["knn_idx"], res["knn_dist"], knn_oracle.d2_fp64(want))
    assert knn_oracle.parity_ok(rep), rep
    names = list(seqs)
    assert res["knn_idx"][names.index(">f1"), 1] == names.index(">f2 x")


def test_integer_k_8_full_column_space(engine):
    """k = 8 on a few hundred random contigs: tens of thousands of observed columns, checked against the vectorised oracle."""
    from karma_b200.engine import profile_and_knn
    asm = synth.s0_iid(300, seed=8)
    res = profile_and_knn(engine, asm.bases, asm.offsets, asm.key_len, 8, n_neighbors=None)
    counts, _ = ko.counts_mode(asm.bases, asm.offsets, 8)
    keep = counts.any(0)
    want = counts[:, keep] / asm.key_len[:, None].astype(np.float64)
    assert res["profile"].shape == want.shape and res["profile"].tobytes() == want.tobytes()
    assert res["columns"] == [ko.code_to_kmer(int(c), 8) for c in np.flatnonzero(keep)]

"""The oracle is only trusted once it reproduces the REAL reference: every case in
tests/golden/kmer_profile_golden.json was produced by /root/reference/karma/kmer.py
(oracle/make_golden.py) and must be matched bit for bit."""
import numpy as np
import pytest

from oracle import kmer_oracle as ko


def _case_dict(c):
    return dict(zip(c["keys"], c["seqs"]))


@pytest.mark.parametrize("impl", ["port", "np"])
def test_oracle_matches_reference_golden(golden, impl):
    f = ko.profile_port if impl == "port" else ko.profile_np
    assert len(golden) >= 15
    for c in golden:
        seqs = _case_dict(c)
        if "exit" in c:
            with pytest.raises(SystemExit) as e:
                f(seqs, c["kmer_size"])
            assert e.value.code == c["exit"]
            continue
        cols, mat = f(seqs, c["kmer_size"])
        assert cols == c["columns"], c["name"]
        assert list(mat.shape) == c["shape"], c["name"]
        assert mat.dtype == np.float64
        assert mat.tobytes().hex() == c["matrix_hex"], c["name"]


def test_known_answers_from_survey(golden):
    by = {c["name"]: c for c in golden}
    ka1 = by["KA1"]
    assert ka1["shape"] == [3, 50]
    assert ka1["columns"][:5] == ["AAAAA", "AAAAAA", "AAAAT", "AAATT", "AACGT"]
    assert ka1["columns"][-4:] == ["acgtA", "cgtAC", "gtACG", "tACGT"]
    pal = [c for c in ka1["columns"] if len(c) == 6]
    assert pal == ["AAAAAA", "CGTTGC", "GAGGAG", "GCAACG", "GGGGGG", "TTGGTT", "TTTTTT"]
    m = np.frombuffer(bytes.fromhex(ka1["matrix_hex"]), dtype=np.float64).reshape(3, 50)
    col = {k: i for i, k in enumerate(ka1["columns"])}
    assert m[1, col["AAAAA"]] == 6 / 11 and m[1, col["AAAAAA"]] == 5 / 11
    assert m[1, col["TTTTT"]] == 6 / 11 and m[1, col["TTTTTT"]] == 5 / 11
    assert m[2, col["GGGGG"]] == 2 / 3
    assert np.count_nonzero(m[0]) == 19 and set(m[0][m[0] != 0]) == {1 / 3}
    ka2 = by["KA2"]
    assert ka2["columns"] == ["ACGTA"] and bytes.fromhex(ka2["matrix_hex"]) == np.array([[0.5]]).tobytes()
    assert by["KA3"]["columns"] == ["ACGT", "CGTA", "GTAC", "TACG", "TTAC", "TTTA", "TTTT"]
    assert by["KA4_short_contig"]["exit"] == 1
    assert by["KA5_cr"]["columns"] == ["A\rCGT", "ACGTA", "CGTA\r", "GTA\rC", "TA\rCG"]


def test_full_column_set_ka6():
    cols = ko.columns_5p6_full()
    assert len(cols) == 1088
    assert cols[:6] == ["AAAAA", "AAAAAA", "AAAAC", "AAAAG", "AAAAT", "AAACA"]
    assert cols[-2:] == ["TTTTT", "TTTTTT"]
    for i, c in enumerate(cols):
        if len(c) == 6:
            assert c == c[::-1] and cols[i - 1] == c[:5]
    # closed form of the survey: col(pal6 r) = p_r + r + 1
    col = {k: i for i, k in enumerate(cols)}
    for r in range(64):
        x1, x2, x3 = r >> 4, (r >> 2) & 3, r & 3
        assert col[ko.code_to_kmer(ko.pal6_rank_to_code(r), 6)] == 256 * x1 + 65 * x2 + 20 * x3 + r + 1


def test_counts_mode_consistent_with_profile():
    from karma_b200 import synth
    asm = synth.s1_families(60, seed=3)
    cols, prof = ko.profile_np(asm.as_dict(), "5p6")
    counts, exotic = ko.counts_mode(asm.bases, asm.offsets, "5p6")
    assert exotic.sum() == 0
    full = ko.columns_5p6_full()
    keep = [i for i, c in enumerate(full) if c in set(cols)]
    assert [full[i] for i in keep] == cols
    assert np.array_equal(counts[:, keep].astype(np.float64) / asm.key_len[:, None].astype(np.float64), prof)
    # dense modes: the 5-mer block of "5+6" is the plain 5-mer count, rows sum to L-4 / L-5
    d56, _ = ko.counts_mode(asm.bases, asm.offsets, "5+6")
    L = np.diff(asm.offsets)
    assert np.array_equal(d56[:, :1024].sum(1), L - 4) and np.array_equal(d56[:, 1024:].sum(1), L - 5)
    k7, _ = ko.counts_mode(asm.bases, asm.offsets, 7)
    assert k7.shape == (60, 16384) and np.array_equal(k7.sum(1), L - 6)


def test_synth_is_deterministic_and_trinity_like():
    from karma_b200 import synth
    a, b = synth.s1_families(500, seed=11), synth.s1_families(500, seed=11)
    assert np.array_equal(a.bases, b.bases) and np.array_equal(a.offsets, b.offsets)
    L = np.diff(synth.s0_iid(2000).offsets)
    assert L.min() >= 200 and L.max() <= 15000 and 500 < np.median(L) < 900
    assert set(np.unique(a.bases)) <= set(b"ACGT")
    d = a.as_dict()
    assert len(d) == 500 and all(len(k) == kl for k, kl in zip(d, a.key_len))
    s3 = synth.s3_long(400)
    assert np.diff(s3.offsets).max() > 30000

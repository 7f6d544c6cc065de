"""Native FASTA packer == read_fasta_file of the reference (karma.py:40-61), host only."""
from collections import OrderedDict

import numpy as np
import pytest

from karma_b200 import fasta


def reference_read_fasta(text_path):
    """Restatement of /root/reference/karma/karma.py:40-61 (text mode, universal newlines)."""
    sequences = OrderedDict()
    with open(text_path, "r") as reader:
        seq_name = reader.readline().rstrip("\n").split(" ")[0]
        sequence = ""
        for line in reader:
            if line.startswith(">"):
                sequences[seq_name] = sequence
                seq_name = line.rstrip("\n").split(" ")[0]
                sequence = ""
            else:
                sequence += line.rstrip("\n")
        sequences[seq_name] = sequence
    return sequences


CASES = {
    "plain": ">a desc here\nACGT\nACG\n>b\nTTTT\n",
    "no_trailing_newline": ">a\nACGT\n>b len=3\nTTT",
    "crlf": ">a x\r\nACGT\r\nAC\r\n>b\r\nGG\r\n",
    "lone_cr": ">a\rACGT\rAC\r>b\rGG",
    "blank_lines": ">a\n\nACGT\n\n\n>b\n\n",
    "empty_sequence": ">a\n>b\nAC\n>c\n",
    "tabs_and_spaces": ">a\tb c\nAC GT\t\n",
    "no_header": "ACGT\nAAAA\n>x\nCC\n",
    "empty_file": "",
    "duplicate_keys": ">a\nAC\n>b\nGG\n>a other\nTTTT\n",
    "gt_in_sequence_line": ">a\nAC>GT\n>b\nA\n",
    "header_only_gt": ">\nACGT\n> name\nGG\n",
}


@pytest.mark.parametrize("threads", [1, 5])
@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_reference_reader(tmp_path, monkeypatch, name, threads):
    monkeypatch.setenv("KB_HOST_THREADS", str(threads))      # 5 threads + 1-byte chunks: cuts inside records
    monkeypatch.setenv("KB_FASTA_MIN_CHUNK", "1")
    path = tmp_path / (name + ".fa")
    with open(path, "wb") as f:
        f.write(CASES[name].encode("ascii"))
    want = reference_read_fasta(path)
    got = fasta.read_fasta_file(path)
    assert list(got.items()) == list(want.items())
    pk = got.packed()
    if name in ("duplicate_keys", "header_only_gt"):   # repeated keys collapse in the dict
        assert pk is None                       # the mapping no longer equals the record list
    else:
        bases, offsets, key_len = pk
        assert [bases[offsets[i]:offsets[i + 1]].tobytes().decode() for i in range(len(want))] == list(want.values())
        assert key_len.tolist() == [len(k) for k in want]


def test_large_random_fasta_and_kmer_pack(tmp_path, monkeypatch):
    from karma_b200 import synth
    from karma_b200.kmer import KmerClustering
    asm = synth.s1_families(300, seed=2)
    d = asm.as_dict()
    path = tmp_path / "asm.fa"
    with open(path, "w") as f:
        for k, s in d.items():
            f.write(k + " len=%d path=[1:2]\n" % len(s))
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    got = fasta.read_fasta_file(path)
    assert list(got.items()) == list(d.items())
    monkeypatch.setenv("KB_FASTA_MIN_CHUNK", "20000")         # ~17 chunks over the 340 KB file
    monkeypatch.setenv("KB_HOST_THREADS", "32")
    assert list(fasta.read_fasta_file(path).items()) == list(d.items())
    bases, offsets, key_len = KmerClustering(got, "/tmp", "5p6", 1)._pack()
    assert np.array_equal(bases, asm.bases) and np.array_equal(offsets, asm.offsets) and np.array_equal(key_len, asm.key_len)
    got[">extra"] = "ACGT"                      # a modified mapping falls back to the generic packer
    b2, o2, k2 = KmerClustering(got, "/tmp", "5p6", 1)._pack()
    assert len(o2) == 302 and b2[-4:].tobytes() == b"ACGT"


def test_non_ascii_is_rejected(tmp_path):
    from karma_b200 import _lib
    path = tmp_path / "x.fa"
    with open(path, "wb") as f:
        f.write(b">a\nAC\xc3\xa9\n")
    with pytest.raises(_lib.KarmaB200Error):
        fasta.read_fasta_file(path)

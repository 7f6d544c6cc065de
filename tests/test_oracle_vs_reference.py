"""Live comparison against the real reference (authoring container only: the tree
/root/reference does not travel to the GPU box, where these tests skip)."""
import random

import pytest

from oracle import kmer_oracle as ko
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


def test_reference_own_test():
    # /root/reference/tests/test_kmer.py:6-8
    cls = ref_shim.load()
    assert cls.is_palindrome("ACGT") is False and cls.is_palindrome("AAAA") is True
    assert ko.is_palindrome("ACGT") is False and ko.is_palindrome("AAAA") is True


@pytest.mark.parametrize("alphabet,k", [("ACGT", "5p6"), ("ACGTN", "5p6"), ("ACGTacgtNRY", "5p6"), ("AC", "5p6"),
                                        ("ACGT", 3), ("ACGTN", 4), ("ACGT", 7)])
def test_randomised_against_reference(alphabet, k):
    rng = random.Random(hash((alphabet, str(k))) & 0xffff)
    for _ in range(3):
        seqs = {}
        for i in range(rng.randint(2, 9)):
            name = ">" + "".join(rng.choice("abcdefgh012") for _ in range(rng.randint(1, 30))) + str(i)
            seqs[name] = "".join(rng.choice(alphabet) for _ in range(rng.randint(8, 400)))
        cols, mat = ref_shim.reference_profile(dict(seqs), k)
        for f in (ko.profile_port, ko.profile_np):
            c2, m2 = f(seqs, k)
            assert c2 == cols and m2.tobytes() == mat.tobytes()


def test_synthetic_assembly_against_reference():
    from karma_b200 import synth
    asm = synth.s1_families(150, seed=5)
    d = asm.as_dict()
    cols, mat = ref_shim.reference_profile(d, "5p6", threads=4)
    c2, m2 = ko.profile_np(d, "5p6")
    assert c2 == cols and m2.tobytes() == mat.tobytes()

"""Freeze golden vectors from the REAL reference (run in the authoring container).

    python -m oracle.make_golden

Runs ``/root/reference/karma/kmer.py`` (through oracle/ref_shim.py) on the
known-answer inputs of SURVEY.md section 8c and on seeded random FASTA, and
writes ``tests/golden/kmer_profile_golden.json``: for every case the input
mapping, the k-mer size, the reference's column list and its float64 matrix as
hex bytes (so the comparison is bit-exact), or ``"exit": 1`` where the
reference calls exit(1) (kmer.py:248,258).
"""
import json
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden", "kmer_profile_golden.json")


def rand_case(rng, alphabet, n, lo, hi, kmer_size):
    seqs = {}
    for i in range(n):
        name = ">" + "".join(rng.choice("abcXYZ_0123456789") for _ in range(rng.randint(1, 30)))
        while name in seqs:
            name += "x"
        seqs[name] = "".join(rng.choice(alphabet) for _ in range(rng.randint(lo, hi)))
    return seqs, kmer_size


def main():
    cases = []
    # --- known-answer inputs
    cases.append(("KA1", {">c1": "ACGTACGTTGCAACGTNACGT",
                          ">contig_two": "AAAAAAAAAATTTTTTTTTT",
                          ">c3": "acgtACGTGGGGGGAGGAGTTGGTT"}, "5p6"))
    cases.append(("KA2", {">a": "ACGTA"}, "5p6"))
    cases.append(("KA3", {">a": "ACGTACGTAC", ">b": "TTTTACGT"}, 4))
    cases.append(("KA4_short_contig", {">a": "ACGTACGT", ">b": "ACG"}, "5p6"))
    cases.append(("KA5_cr", {">a": "ACGTA\rCGT"}, "5p6"))
    cases.append(("pal6_all", {">p": "".join(a + b + c + c + b + a for a in "ACGT" for b in "ACGT" for c in "ACGT")}, "5p6"))
    cases.append(("homopolymer", {">h": "A" * 700, ">g": "ACGT" * 100}, "5p6"))
    rng = random.Random(20261018)
    specs = [("ACGT", 12, 8, 400, "5p6"), ("ACGTN", 10, 8, 300, "5p6"),
             ("ACGTacgtNRY", 8, 8, 200, "5p6"), ("AC", 6, 8, 100, "5p6"),
             ("ACGT", 10, 8, 300, 3), ("ACGT", 10, 8, 300, 4), ("ACGT", 6, 20, 400, 7),
             ("ACGTN", 8, 8, 200, 4), ("ACGT", 40, 200, 1500, "5p6"),
             ("ACGTN", 5, 6, 40, 6), ("ACGT", 5, 5, 9, "5p6")]
    for i, (alpha, n, lo, hi, k) in enumerate(specs):
        seqs, k = rand_case(rng, alpha, n, lo, hi, k)
        cases.append((f"rand{i}_{alpha}_k{k}", seqs, k))

    out = []
    for name, seqs, k in cases:
        rec = {"name": name, "kmer_size": k, "keys": list(seqs.keys()),
               "seqs": list(seqs.values())}
        try:
            cols, mat = ref_shim.reference_profile(dict(seqs), k, threads=2)
            assert mat.dtype == np.float64
            rec.update(columns=cols, shape=list(mat.shape),
                       matrix_hex=np.ascontiguousarray(mat).tobytes().hex())
        except SystemExit as e:
            rec["exit"] = int(e.code)
        out.append(rec)
        print(name, rec.get("shape", "exit"))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump({"generator": "oracle/make_golden.py",
                   "reference": "/root/reference/karma/kmer.py (shimmed, see oracle/ref_shim.py)",
                   "cases": out}, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

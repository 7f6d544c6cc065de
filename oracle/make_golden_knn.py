"""Freeze the kNN lists that umap-learn 0.3.9's exact small-data branch produces (its algorithm
restated in oracle/knn_oracle.knn_umap_small_data on top of scikit-learn's own pairwise_distances)
for seeded synthetic assemblies.

    python -m oracle.make_golden_knn

Writes tests/golden/knn_umap_small_golden.npz: per case the SHA-1 of the generated bases (so a
change of the generator is noticed), the kNN indices (int32) and distances (float32) for
n_neighbors = 2 (karma's default) and 15 (UMAP's default) over the float32-cast `5p6` profile.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth  # noqa: E402
from oracle import kmer_oracle as ko  # noqa: E402
from oracle import knn_oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "knn_umap_small_golden.npz")
CASES = [("S1", 1500), ("S2", 1200), ("S0", 600)]


def main():
    out = {}
    for kind, n in CASES:
        asm = synth.make(kind, n)
        _, prof = ko.profile_np(asm.as_dict(), "5p6")
        tag = "%s_%d" % (kind, n)
        out[tag + "_sha1"] = np.frombuffer(hashlib.sha1(asm.bases.tobytes()).digest(), dtype=np.uint8)
        for k in (2, 15):
            idx, dist = knn_oracle.knn_umap_small_data(prof, k)
            out["%s_k%d_idx" % (tag, k)] = idx.astype(np.int32)
            out["%s_k%d_dist" % (tag, k)] = dist
            truth = knn_oracle.d2_fp64(prof)
            rep = knn_oracle.check_knn(idx, dist, truth, rtol=1e-4)
            print(tag, "k", k, "vs fp64 truth:", rep)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()

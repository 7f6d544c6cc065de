"""Tiny run of every hot kernel for compute-sanitizer (racecheck / memcheck): K1 unfused (+ long split), K1+K3 fused,
K3, K4 tcgen05 (2-CTA, piece table, insertion queues), K4 SIMT, K4x, K5, K6."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth, _lib
from karma_b200.engine import Engine, mode_of, profile_and_knn
which = sys.argv[1] if len(sys.argv) > 1 else "all"
eng = Engine(0)
asm = synth.s1_families(int(os.environ.get("KB_SAN_N", "700")), seed=3)
if which in ("all", "count"):
    d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
    for mode in ("5p6", "5+6"):
        counts, exo, pres = eng.count(d_b, d_o, asm.n, mode_of(mode))
        eng.normalise(counts, counts.shape[1], d_l)
    torch.cuda.synchronize()
    print("count ok")
if which in ("all", "knn"):
    for k, impl in ((3, _lib.KB_KNN_TC), (15, _lib.KB_KNN_TC), (3, _lib.KB_KNN_SIMT)):
        res = profile_and_knn(eng, asm.bases, asm.offsets, asm.key_len, "5p6", n_neighbors=k, impl=impl)
        assert (res["knn_idx"][:, 0] == np.arange(asm.n)).all()
    res = profile_and_knn(eng, asm.bases, asm.offsets, asm.key_len, "5p6", n_neighbors=70, impl=_lib.KB_KNN_TC)   # K6
    assert res["uncertified"] == asm.n
    print("knn ok")

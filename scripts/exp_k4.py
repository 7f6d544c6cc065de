"""Experiment: K4 (+K5) time for query shards of the 50k problem under schedule variants (env KB_KNN_SCHED, KB_KNN_SPLITS, KB_KNN_L2_MB)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth, _lib
from karma_b200.engine import Engine, mode_of
eng = Engine(0); eng.enable_timing(True)
kmer = sys.argv[1] if len(sys.argv) > 1 else "5p6"
ks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "2").split(",")]
shards = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "1,2,4,8").split(",")]
asm = synth.s1_families(50000)
m = mode_of(kmer)
d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
counts, _, _ = eng.count(d_b, d_o, asm.n, m)
cols = counts.shape[1]
_, operand, rowmeta = eng.normalise(counts, cols, d_l, want_profile=False)
tag = "sched=%s splits=%s l2=%s" % (os.environ.get("KB_KNN_SCHED", "-"), os.environ.get("KB_KNN_SPLITS", "-"), os.environ.get("KB_KNN_L2_MB", "-"))
for k in ks:
    for w in shards:
        nq = -(-asm.n // w)
        q0 = nq * (w // 2) if w > 1 else 0
        if q0 + nq > asm.n:
            q0 = asm.n - nq
        for _ in range(3):
            eng.knn_enqueue(operand, rowmeta, k, q_row0=q0, nq=nq, impl=_lib.KB_KNN_TC)
        torch.cuda.synchronize(); eng.stage_ms("knn_gemm"); eng.stage_ms("rerank")
        for _ in range(10):
            eng.knn_enqueue(operand, rowmeta, k, q_row0=q0, nq=nq, impl=_lib.KB_KNN_TC)
        torch.cuda.synchronize()
        g, _ = eng.stage_ms("knn_gemm"); r, _ = eng.stage_ms("rerank")
        tf = 2.0 * nq * asm.n * cols / (g / 1e3) / 1e12
        print("K4 %s cols=%d k=%d nq=%d: gemm %.4f ms = %.0f TFLOP/s, rerank %.4f ms" % (tag, cols, k, nq, g, tf, r))

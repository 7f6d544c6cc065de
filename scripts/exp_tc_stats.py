"""Experiment: insertion statistics of the K4 epilogue (needs a build with -DKB_TC_STATS)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import _lib, synth
from karma_b200.engine import Engine, mode_of
eng = Engine(0); eng.enable_timing(True)
lib = ctypes.CDLL(_lib.LIB_PATH)
asm = synth.s1_families(50000)
d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
counts, _, _ = eng.count(d_b, d_o, asm.n, mode_of("5p6"))
_, op, meta = eng.normalise(counts, 1088, d_l, want_profile=False)
out = (ctypes.c_ulonglong * 3)()
for k in (2, 15):
    for rep in range(2):
        lib.kb_debug_tc_stats(out, 1)
        eng.knn(op, meta, k, impl=_lib.KB_KNN_TC); torch.cuda.synchronize()
        lib.kb_debug_tc_stats(out, 0)
    ms, _ = eng.stage_ms("knn_gemm")
    print("k=%d splits=%s: K4 %.3f ms inserts/row %.1f cold chunk fraction %.3f (chunk-warps %d)" % (
        k, os.environ.get("KB_KNN_SPLITS", "auto"), ms, out[0] / asm.n, out[1] / max(out[2], 1), out[2]))

"""Experiment: K1 time under cost-model builds (KB_K1_EXPERIMENT)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth
from karma_b200.engine import Engine, mode_of
eng = Engine(0); eng.enable_timing(True)
asm = synth.s1_families(50000)
d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
for mode in ("5p6", "5+6", 5):
    m = mode_of(mode)
    for _ in range(3):
        eng.count(d_b, d_o, asm.n, m)
    torch.cuda.synchronize(); eng.stage_ms("count")
    for _ in range(10):
        eng.count(d_b, d_o, asm.n, m)
    torch.cuda.synchronize()
    ms, n = eng.stage_ms("count")
    print("K1 mode %s: %.4f ms (%d launches) exp=%s" % (mode, ms, n, os.environ.get("KB_NVCC_EXTRA", "")))

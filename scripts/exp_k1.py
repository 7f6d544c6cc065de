"""Experiment: K1 time per variant (env KB_K1_OCC, KB_K1_FIRST) at the full and the per-rank shard size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth
from karma_b200._lib import check, ptr
from karma_b200.engine import Engine, mode_of
eng = Engine(0); eng.enable_timing(True)
full = synth.s1_families(50000)
tag = "occ=%s first=%s" % (os.environ.get("KB_K1_OCC", "5"), os.environ.get("KB_K1_FIRST", "2048"))
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["5p6"]
for n in (50000, 25000, 12500, 6250):
    asm = full.slice(0, n)
    d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
    for mode in modes:
        m = mode_of(int(mode) if mode.isdigit() else mode)
        cols = eng.lib.kb_mode_columns(m)
        counts = torch.empty((n, cols), dtype=torch.int32, device="cuda")
        exo = torch.empty(n, dtype=torch.int32, device="cuda")
        pres = torch.zeros(cols + 1, dtype=torch.int32, device="cuda")
        for track in (1, 0):
            def run():
                check(eng.lib.kb_count(eng.ctx, m, ptr(d_b), ptr(d_o), n, ptr(counts), cols, ptr(exo), ptr(pres) if track else None))
            for _ in range(3):
                run()
            torch.cuda.synchronize(); eng.stage_ms("count"); eng.stage_ms("count_long")
            for _ in range(20):
                run()
            torch.cuda.synchronize()
            ms, nl = eng.stage_ms("count")
            ms2, _ = eng.stage_ms("count_long")
            gb = (float(asm.offsets[-1]) + 4.0 * n * cols) / 1e9
            print("K1 %s n=%d mode=%s track=%d: %.4f ms (+long %.4f) %.0f GB/s" % (tag, n, mode, track, ms, ms2, gb / (ms / 1e3)))

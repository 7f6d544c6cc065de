"""BASELINE configs[4]: k-size sweep (-k 4+5 / 5p6 / 5+6 / 7 = 16384 columns) at 200k contigs and the long-contig input
(S3: 1 % of the contigs 50-200 kb).  Device-resident passes (PassPlan, eager with the library's event pairs on): time of the
fused counting kernel against its HBM roofline and of K4 against the tensor peak, per shape.  One JSON line per case."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth, _lib
from karma_b200.engine import Engine, PassPlan, device_pass
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
eng = Engine(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cases = [("S1", "4+5", 2), ("S1", "5p6", 2), ("S1", "5+6", 15), ("S1", 7, 2), ("S3", "5p6", 2)]
if len(sys.argv) > 2:                                   # e.g. "S3" or "S1:7"
    want = sys.argv[2].split(",")
    cases = [c for c in cases if c[0] in want or ("%s:%s" % (c[0], c[1])) in want]
for kind, kmer, k in cases:
    asm = synth.make(kind, n)
    total = int(asm.offsets[-1])
    rec = {"synth": kind, "contigs": n, "kmer": str(kmer), "n_neighbors": k, "total_bases": total,
           "longest_contig": int(np.diff(asm.offsets).max())}
    plan = PassPlan(eng, asm.n, total, kmer, n_neighbors=k, impl=_lib.KB_KNN_TC, want_profile=True, graph=False)
    plan.load(asm.bases, asm.offsets, asm.key_len)
    eng.enable_timing(True)
    chk = plan.check(plan.run())
    for st in ("count", "knn_gemm", "rerank"):
        eng.stage_ms(st)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    ev0.record()
    for _ in range(steps):
        tok = plan.run()
    ev1.record()
    chk = plan.check(tok)
    torch.cuda.synchronize()
    c_ms, _ = eng.stage_ms("count"); g_ms, _ = eng.stage_ms("knn_gemm"); r_ms, _ = eng.stage_ms("rerank")
    eng.enable_timing(False)
    cols, dp = plan.cols, plan.dp
    bytes_k1 = total + n * (8.0 * cols + 2.0 * dp + 32.0)
    rec.update(columns=cols, optimistic_ok=chk["ok"], flags_or=chk["flags_or"], uncertified=chk["uncertified"],
               ms_per_pass=ev0.elapsed_time(ev1) / steps, count_ms=c_ms, knn_gemm_ms=g_ms, rerank_ms=r_ms,
               count_gbs=bytes_k1 / (c_ms / 1e3) / 1e9, count_frac_of_hbm_peak=bytes_k1 / (c_ms / 1e3) / 1e9 / peaks["hbm_gbs"],
               k4_tflops=2.0 * n * n * cols / (g_ms / 1e3) / 1e12,
               k4_frac_of_sustained_peak=2.0 * n * n * cols / (g_ms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"])
    if not chk["ok"]:
        # the general path (rows beyond the tensor range -> exact side path): one eager pass, wall clock
        plan.close(); del plan; torch.cuda.empty_cache()
        d_b, d_o, d_l = eng.upload(asm.bases, asm.offsets, asm.key_len)
        times = []
        for it in range(2):                              # the first pass also plans, allocates and uploads the piece table
            eng.enable_timing(True)
            for st in ("count", "count_long", "compact", "normalise", "knn_gemm", "knn_exact", "rerank"):
                eng.stage_ms(st)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            out = device_pass(eng, d_b, d_o, d_l, asm.n, kmer, n_neighbors=k, impl=_lib.KB_KNN_TC)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            stages = {}
            for st in ("count", "count_long", "compact", "normalise", "knn_gemm", "knn_exact", "rerank"):
                ms, cnt = eng.stage_ms(st)
                stages[st] = {"mean_ms": round(ms, 3), "launches": cnt}
            rec["general_path_stage_ms"] = stages
            eng.enable_timing(False)
        rec["general_path_ms"] = times[-1]
        rec["general_path_first_call_ms"] = times[0]
        rec["flagged_rows"] = int((out["rowmeta"][:asm.n, 3] & 3).ne(0).sum().item())
        rec["general_path_note"] = "optimistic pass + validation + redo with K1/K2/K3 unfused, flagged rows through the exact side path (K4x: integer Gram entries); stage times cover BOTH passes of the call"
        del out
    else:
        plan.close(); del plan
    torch.cuda.empty_cache()
    print(json.dumps(rec), flush=True)

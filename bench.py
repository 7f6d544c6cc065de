#!/usr/bin/env python
"""Benchmark of the karma k-mer front end (profile matrix + exact kNN graph).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): contigs/s for (profile + kNN).  Workload at every N is
BASELINE.json configs[1]: a 50 000-contig synthetic Trinity-like assembly (S1 gene
families, log-normal lengths), kmer.py's default -k 5p6 (1088 columns),
n_neighbors=2.  N>1 is strong scaling: the same 50k contigs, rows sharded for
counting, query rows sharded for the kNN after an all-gather of the operand.

A step = one pass of the hot path over the whole assembly.
  value : inputs already resident in HBM, CUDA-event time (max over ranks)
  e2e   : host (pinned) buffers in, host results out -- H2D, column dictionary,
          kernels, D2H of the float64 profile and the kNN lists all inside the
          timed region (wall clock, device-synchronised on both sides)
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contigs/sec (profile+kNN)"
UNIT = "contigs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--contigs", type=int, default=50000)
    ap.add_argument("--synth", default="S1")
    ap.add_argument("--kmer", default="5p6")
    ap.add_argument("--neighbors", type=int, default=2)
    ap.add_argument("--knn-impl", default="tc", choices=["tc", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shard-gen", action="store_true", help="each rank synthesises only its own row shard")
    ap.add_argument("--profile-host", action="store_true", help="cProfile the timed loop on rank 0 (diagnostics)")
    return ap.parse_args()


def kmer_arg(s):
    return int(s) if s.isdigit() else s


def workload_name(a):
    return "%d-contig synthetic Trinity-like assembly (%s), -k %s, n_neighbors=%d" % (
        a.contigs, a.synth, a.kmer, a.neighbors)


def ncu_traffic(kernel_regex):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the
    committed `ncu --set full` summary of the same bench command (profiles/): None if absent."""
    import glob
    import re
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "ncu_k4_*summary.csv"))):
        rd = wr = None
        name_ok = False
        for line in open(path):
            parts = line.rstrip("\n").split(",")
            if parts[0] == "Kernel Name" and re.search(kernel_regex, line):
                name_ok = True
            if parts[0] == "dram__bytes_read.sum":
                rd = (float(parts[-1]), parts[-2])
            if parts[0] == "dram__bytes_write.sum":
                wr = (float(parts[-1]), parts[-2])
        if rd and wr:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            best = {"bytes": rd[0] * scale.get(rd[1], 1.0) + wr[0] * scale.get(wr[1], 1.0), "source": os.path.relpath(path, ROOT)}
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops"], "tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for nme, v in zip(names, r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(nme)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------ CPU arm (oracle port)
def _port_chunk(args):
    """Worker: kmer.py's counting loops (oracle port) on a chunk of the sample."""
    from oracle import kmer_oracle
    keys, seqs, kmer_size = args
    counters = [kmer_oracle._contig_counter(s, kmer_size) for s in seqs]
    return counters


def cpu_profile_rate(asm, kmer_size, n_sample, procs):
    """contigs/s of the kmer.py algorithm (oracle port) on the first n_sample contigs,
    counting loops spread over `procs` processes (more than kmer.py itself does:
    its counting is single-threaded, kmer.py:56-92)."""
    from multiprocessing import get_context
    from oracle import kmer_oracle
    sub = asm.slice(0, n_sample)
    d = sub.as_dict()
    keys, seqs = list(d.keys()), list(d.values())
    t0 = time.perf_counter()
    if procs > 1:
        step = -(-len(seqs) // procs)
        chunks = [(keys[i:i + step], seqs[i:i + step], kmer_size) for i in range(0, len(seqs), step)]
        with get_context("fork").Pool(procs) as pool:
            parts = pool.map(_port_chunk, chunks)
        counters = [c for p in parts for c in p]
    else:
        counters = _port_chunk((keys, seqs, kmer_size))
    kset = set()
    for c in counters:
        kset.update(c.keys())
    cols = sorted(kset)
    col = {k: i for i, k in enumerate(cols)}
    out = np.zeros((len(seqs), len(cols)), dtype=np.float64)
    for r, (key, c) in enumerate(zip(keys, counters)):
        ln = len(key)
        for kmer, cnt in c.items():
            out[r, col[kmer]] = cnt / ln
    dt = time.perf_counter() - t0
    return n_sample / dt, dt


def cpu_knn_rate(profile32, k, n_query):
    """queries/s of an exact CPU neighbour search (what UMAP does below 4096 points:
    full pairwise distances + partial sort), fp32 as UMAP casts, BLAS on all threads,
    for the first n_query rows against ALL keys."""
    t0 = time.perf_counter()
    q = profile32[:n_query]
    sq = np.einsum("ij,ij->i", profile32, profile32)
    d2 = sq[:n_query, None] + sq[None, :] - 2.0 * (q @ profile32.T)
    idx = np.argpartition(d2, k, axis=1)[:, :k]
    dd = np.take_along_axis(d2, idx, 1)
    order = np.argsort(dd, axis=1)
    idx = np.take_along_axis(idx, order, 1)
    dt = time.perf_counter() - t0
    return n_query / dt, dt, idx


def cpu_arm(a, asm, steps, warmup):
    """contigs/s of the CPU path, extrapolated from a bounded sample of the same
    workload: profile is linear in contigs, exact kNN linear in query rows."""
    from oracle import kmer_oracle
    cores = os.cpu_count() or 1
    kmer_size = kmer_arg(a.kmer)
    n_sample = min(asm.n, 1500)
    n_query = min(asm.n, 1024)
    # keys for the neighbour search: the full profile from the vectorised oracle (untimed)
    counts, _ = kmer_oracle.counts_mode(asm.bases, asm.offsets, kmer_size)
    prof32 = (counts / asm.key_len[:, None].astype(np.float64)).astype(np.float32)
    del counts
    vals, times = [], []
    for it in range(warmup + steps):
        rp, tp = cpu_profile_rate(asm, kmer_size, n_sample, cores)
        rk, tk, _ = cpu_knn_rate(prof32, a.neighbors, n_query)
        v = 1.0 / (1.0 / rp + 1.0 / rk)
        if it >= warmup:
            vals.append(v); times.append(tp + tk)
    value = statistics.mean(vals)
    sample = ("per step: oracle port of kmer.py (pure-Python counting loops over %d processes) on %d contigs + exact "
              "fp32 brute-force kNN (numpy/BLAS, %d threads) of %d query rows vs all %d keys; contigs/s = "
              "1/(1/profile_rate + 1/knn_rate); umap-learn/NN-descent not installed" %
              (cores, n_sample, cores, n_query, asm.n))
    return value, cores, sample, statistics.mean(times) * 1e3


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from karma_b200 import synth
    asm = synth.make(a.synth, a.contigs)
    steps = max(1, min(a.steps, 3))
    warm = min(a.warmup, 1)
    value, cores, sample, ms = cpu_arm(a, asm, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "python int / f64 (profile), f32 (kNN)", "data": "synthetic",
            "config": {"workload": workload_name(a)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from karma_b200 import _lib, synth
    from karma_b200.engine import Engine, device_pass, mode_of, profile_and_knn, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000)] + sys.argv
        return subprocess.call(cmd)
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    eng = Engine(local)
    eng.enable_timing(True)
    kmer_size = kmer_arg(a.kmer)
    mode = mode_of(kmer_size)
    impl = _lib.KB_KNN_TC if a.knn_impl == "tc" else _lib.KB_KNN_SIMT
    k = a.neighbors

    n_total = a.contigs
    lo, hi, per = shard_bounds(n_total, world, rank)
    if world > 1 and (a.shard_gen or n_total >= 400000):
        # large assemblies: every rank synthesises only its own rows (seeded per rank)
        shard = synth.make(a.synth, hi - lo, seed=4321 + rank)
        asm = None
        tb = torch.tensor([int(shard.offsets[-1])], dtype=torch.int64, device="cuda")
        dist.all_reduce(tb)
        total_bases_all = int(tb.item())
    else:
        asm = synth.make(a.synth, a.contigs)
        shard = asm.slice(lo, hi) if world > 1 else asm
        total_bases_all = int(asm.offsets[-1])
    n = shard.n
    h_bases = torch.from_numpy(shard.bases.copy()).pin_memory()
    h_offsets = torch.from_numpy(shard.offsets.copy()).pin_memory()
    h_keylen = torch.from_numpy(shard.key_len.copy()).pin_memory()

    d_bases, d_offsets, d_keylen = eng.upload(h_bases, h_offsets, h_keylen)
    cols_full = eng.lib.kb_mode_columns(mode)
    b_counts = torch.empty((n, cols_full), dtype=torch.int32, device=eng.device)
    b_exotic = torch.empty(n, dtype=torch.int32, device=eng.device)
    b_presence = torch.empty(cols_full + 1, dtype=torch.int32, device=eng.device)
    state = {}

    bufs = {"counts": b_counts, "exotic": b_exotic, "presence": b_presence}

    def device_step():
        r = device_pass(eng, d_bases, d_offsets, d_keylen, n, kmer_size, n_neighbors=k, impl=impl, want_profile=True,
                        group=group, rank=rank, world=world, n_total=n_total, gather_lists=True, bufs=bufs)
        state.update(idx=r.get("all_idx", r["idx"]), dist=r.get("all_dist", r["dist"]), profile=r["profile"],
                     d_cols=r["d_cols"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for st in ("count", "normalise", "knn_gemm", "rerank"):
        eng.stage_ms(st)                                    # drop the warm-up launches
    barrier()
    prof = None
    if a.profile_host and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    ev0.record()
    for _ in range(a.steps):
        device_step()
    ev1.record()
    if prof is not None:
        prof.disable()
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(28)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.launches() - launches0
    # mean per-launch device time of each kernel over the timed region (CUDA-event pairs the
    # library records on the launching stream around every launch; read after the region)
    gemm_ms, gemm_n = eng.stage_ms("knn_gemm")
    count_ms, _ = eng.stage_ms("count")
    rerank_ms, _ = eng.stage_ms("rerank")
    norm_ms, _ = eng.stage_ms("normalise")
    t = torch.tensor([total_ms], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_total * a.steps / (total_ms / 1e3)
    d_cols = state["d_cols"]

    # ---- e2e: host buffers in, host results out
    e2e = None
    if not a.no_e2e:
        def e2e_step():
            return profile_and_knn(eng, h_bases, h_offsets, h_keylen, kmer_size, n_neighbors=k, impl=impl,
                                   group=group, rank=rank, world=world, row0=lo, n_total=n_total, reuse_host=True)
        for _ in range(min(a.warmup, 2)):
            res = e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            res = e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = h_bases.numel() + h_offsets.numel() * 8 + h_keylen.numel() * 4
        d2h = res["profile"].nbytes + res["knn_idx"].nbytes + res["knn_dist"].nbytes + n + cols_full * 4
        e2e = {"value": n_total * a.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt / a.steps * 1e3}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    flops = 2.0 * n * n_total * d_cols                     # this rank's query rows x all keys
    ach_tf = flops / (gemm_ms / 1e3) / 1e12
    count_bytes = float(shard.offsets[-1]) + 4.0 * n * cols_full
    ach_gbs = count_bytes / (count_ms / 1e3) / 1e9
    # the same counting kernel on north_star's dense 5120-column shape (supplementary: outside the timed region)
    dense = None
    if world == 1 and mode != _lib.KB_MODE_DENSE_5_6:
        eng.stage_ms("count")
        for _ in range(5):
            eng.count(d_bases, d_offsets, n, _lib.KB_MODE_DENSE_5_6)
        torch.cuda.synchronize()
        dms, dn = eng.stage_ms("count")
        dbytes = float(shard.offsets[-1]) + 4.0 * n * 5120
        dense = {"bound": "hbm", "kernel": "k1_count (5120 dense columns)", "achieved": dbytes / (dms / 1e3) / 1e9,
                 "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dbytes / (dms / 1e3) / 1e9 / pk["hbm_gbs"],
                 "bytes_per_launch": dbytes, "launches": dn}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32 counts, f16 x f16 -> f32 Gram (tcgen05), f64 profile + rerank", "data": "synthetic",
        "config": {"workload": workload_name(a), "columns": d_cols, "total_bases": total_bases_all,
                   "knn_impl": a.knn_impl,
                   "l2": "no explicit flush: one step streams %.2f GB of inputs+intermediates (> 126 MB L2)" %
                         ((total_bases_all + n_total * d_cols * (4 + 8 + 2)) / 1e9)},
        "clocks": clocks, "gpu_launches": int(launches),
        "stage_ms": {"count": count_ms, "normalise": norm_ms, "knn_gemm": gemm_ms, "rerank": rerank_ms},
        "roofline": {"bound": "tensor", "kernel": "k4_tc2 (distance GEMM, 2-CTA tcgen05 MMA + fused top-k)" if a.knn_impl == "tc" else "k4_simt",
                     "achieved": ach_tf, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach_tf / pk["tflops"],
                     "traffic": (ncu_traffic("k4_tc") or {}).get("bytes"), "traffic_source": (ncu_traffic("k4_tc") or {}).get("source"),
                     "peak_source": pk["source"] + ", bf16 burst; sustained %s" % pk["tflops_sustained"],
                     "flops_per_launch": flops},
        "roofline_count": {"bound": "hbm", "kernel": "k1_count", "achieved": ach_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": ach_gbs / pk["hbm_gbs"], "bytes_per_launch": count_bytes},
    }
    if dense:
        line["roofline_count_dense5120"] = dense
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not a.no_cpu_baseline and asm is not None:
        v, cores, sample, _ = cpu_arm(a, asm, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))

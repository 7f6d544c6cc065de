"""Multi-GPU parity: sharded run (torchrun, one rank per GPU) == single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/multi_gpu_check.py [--contigs 6000] [--neighbors 15]

Every rank counts its row shard, the column dictionary is agreed through the presence
all-reduce, the kNN operand is all-gathered and each rank searches its query rows.
Rank 0 then recomputes everything alone on its GPU and compares: the profile rows must
be bit-identical and the kNN lists identical (indices and distances).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import _lib, synth  # noqa: E402
from karma_b200.engine import Engine, PassPlan, profile_and_knn, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=6001)
    ap.add_argument("--neighbors", type=int, default=15)
    ap.add_argument("--synth", default="S1")
    ap.add_argument("--kmer", default="5p6")
    ap.add_argument("--exotic", action="store_true", help="inject N / lowercase bases into contigs of different ranks")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local)
    kmer = int(a.kmer) if a.kmer.isdigit() else a.kmer
    asm = synth.make(a.synth, a.contigs)
    if a.exotic:
        # windows with non-ACGT bytes get string-keyed columns (kmer.py has no alphabet); put different
        # ones on different ranks so that the key union and the column merge cross the rank boundary
        bases = asm.bases.copy()
        for i, (pos, ch) in zip((1, asm.n // 2 + 1, asm.n - 1), ((20, b"N"), (33, b"n"), (7, b"R"))):
            bases[asm.offsets[i] + pos] = ch[0]
            bases[asm.offsets[i] + pos + 9] = ord("a")
        asm = synth.Assembly(bases, asm.offsets, asm.gene, asm.iso)
    lo, hi, per = shard_bounds(asm.n, world, rank)
    shard = asm.slice(lo, hi)
    res = profile_and_knn(eng, shard.bases, shard.offsets, shard.key_len, kmer, n_neighbors=a.neighbors,
                          impl=_lib.KB_KNN_TC, group=dist.group.WORLD, rank=rank, world=world, row0=lo, n_total=asm.n)
    # the pre-planned pass: peer-memory exchange, three passes back to back (eager, eager, CUDA graph replay);
    # every rank must end up with ALL k-lists, and its own rows must equal the eager / NCCL path bit for bit
    plan_ok = True
    if not a.exotic and a.synth != "S3" and shard.n > 0:      # (exotic bytes / flagged rows: the plan reports them, the eager path serves them)
        plan = PassPlan(eng, shard.n, int(shard.offsets[-1]), kmer, n_neighbors=a.neighbors, impl=_lib.KB_KNN_TC,
                        group=dist.group.WORLD, rank=rank, world=world, n_total=asm.n, graph=not a.no_graph)
        plan.load(shard.bases, shard.offsets, shard.key_len)
        plan.capture(warmup=2)
        for it in range(3):
            tok = plan.run()
            chk = plan.check(tok)
            if chk["uncertified"]:
                plan.fixup()
            torch.cuda.synchronize()
            mine_i = plan.idx.cpu().numpy(); mine_d = plan.dist.cpu().numpy()
            same = np.array_equal(mine_i, res["knn_idx"]) and np.array_equal(mine_d, res["knn_dist"]) and chk["ok"]
            prof_same = plan.profile.cpu().numpy().tobytes() == res["profile"].tobytes()
            plan_ok &= same and prof_same
            print("rank %d plan pass %d: own rows == eager path %s, profile %s, check %r" % (rank, it, same, prof_same, chk))
        plan_all_idx = plan.all_idx.cpu().numpy()[:asm.n]
        plan_all_dist = plan.all_dist.cpu().numpy()[:asm.n]
        # the host-to-host form of the same pass (pinned inputs in, pinned results out, one graph): twice
        plan.bind_host(shard.bases, shard.offsets, shard.key_len, chunks=3)
        for it in range(2):
            hr = plan.run_host()
            same = hr["ok"] and np.array_equal(hr["knn_idx"], res["knn_idx"]) and np.array_equal(hr["knn_dist"], res["knn_dist"]) \
                and hr["profile"].tobytes() == res["profile"].tobytes()
            plan_ok &= bool(same)
            print("rank %d host pass %d: == eager path %s" % (rank, it, same))
        plan.close()
    else:
        plan_all_idx = plan_all_dist = None
    # gather the shards' results on rank 0 (object gather: this is a test, not the data path)
    gathered = [None] * world
    dist.gather_object((lo, hi, res["columns"], res["profile"], res["knn_idx"], res["knn_dist"], plan_ok, plan_all_idx, plan_all_dist),
                       gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        full = profile_and_knn(eng, asm.bases, asm.offsets, asm.key_len, kmer, n_neighbors=a.neighbors, impl=_lib.KB_KNN_TC)
        for gr, (glo, ghi, cols, prof, idx, dst, p_ok, p_idx, p_dst) in enumerate(gathered):
            if p_idx is not None:
                all_same = np.array_equal(p_idx, full["knn_idx"]) and np.array_equal(p_dst, full["knn_dist"])
                print("rank %d: planned passes ok %s, gathered lists of ALL rows == single GPU %s" % (gr, p_ok, all_same))
                ok &= bool(p_ok) and all_same
            same_cols = cols == full["columns"]
            same_prof = prof.tobytes() == full["profile"][glo:ghi].tobytes()
            same_idx = np.array_equal(idx, full["knn_idx"][glo:ghi])
            same_dst = np.array_equal(dst, full["knn_dist"][glo:ghi])
            print("rows [%d,%d): columns %s profile %s knn_idx %s knn_dist %s" % (glo, ghi, same_cols, same_prof, same_idx, same_dst))
            ok &= same_cols and same_prof and same_idx and same_dst
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED", "world", world, "contigs", asm.n, "k", a.neighbors)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

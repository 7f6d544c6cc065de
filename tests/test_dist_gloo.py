"""world_size-2 gloo run of the multi-rank host logic (CPU): row sharding, presence
all-reduce -> identical column dictionary on every rank, padded all-gather of the
kNN operand with index == global row."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from karma_b200 import synth
        from karma_b200._lib import KB_MODE_5P6
        from karma_b200.engine import all_gather_padded, merge_columns, mode_column_names, shard_bounds
        from oracle import kmer_oracle as ko
        asm = synth.s0_iid(n_total, seed=99)
        # make some 5-mers globally absent so that compaction matters, and absent on ONE rank only
        lo, hi, per = shard_bounds(n_total, world, rank)
        shard = asm.slice(lo, hi)
        counts, _ = ko.counts_mode(shard.bases, shard.offsets, "5p6")
        counts[:, 5] = 0                          # absent everywhere
        if rank == 0:
            counts[:, 9] = 0                      # present on rank 1 only
        presence = torch.from_numpy((counts != 0).any(0).astype(np.int32))
        dist.all_reduce(presence, op=dist.ReduceOp.MAX)
        names = mode_column_names(KB_MODE_5P6)
        cols, colmap, _ = merge_columns(names, presence.numpy() != 0, [])
        local = torch.from_numpy(counts.astype(np.int32))
        gathered = all_gather_padded(local, hi - lo, per, None, 0)
        lens = all_gather_padded(torch.from_numpy(shard.key_len.copy()), hi - lo, per, None, 1)
        q.put((rank, cols, colmap.tolist(), gathered.numpy(), lens.numpy(), lo, hi, per))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding():
    world, n_total = 2, 41                        # odd: the last shard is short
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, c0, m0, g0, l0, lo0, hi0, per), (r1, c1, m1, g1, l1, lo1, hi1, _) = res
    assert c0 == c1 and m0 == m1, "ranks disagree on the column dictionary"
    assert np.array_equal(g0, g1) and np.array_equal(l0, l1)
    assert (lo0, hi0, lo1, hi1) == (0, 21, 21, 41) and per == 21
    from karma_b200 import synth
    from oracle import kmer_oracle as ko
    asm = synth.s0_iid(n_total, seed=99)
    full, _ = ko.counts_mode(asm.bases, asm.offsets, "5p6")
    full[:, 5] = 0
    full[:21, 9] = 0
    assert np.array_equal(g0[:n_total], full.astype(np.int32)), "gathered row r must be global row r"
    assert (g0[n_total:] == 0).all() and np.array_equal(l0[:n_total], asm.key_len)
    names = ko.columns_5p6_full()
    assert names[5] not in c0 and names[9] in c0 and len(c0) == 1087

"""Diagnostics (torchrun, N ranks): where the time of one planned multi-GPU pass goes.  Eager passes with CUDA events between
the stages (ranks aligned by a barrier before every pass), next to the graph-replay time of the same plan."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from karma_b200 import synth, _lib
from karma_b200.engine import Engine, PassPlan, shard_bounds

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
asm = synth.make("S1", n_total)
lo, hi, per = shard_bounds(n_total, world, rank)
shard = asm.slice(lo, hi)
plan = PassPlan(eng, shard.n, int(shard.offsets[-1]), "5p6", n_neighbors=2, impl=_lib.KB_KNN_TC, group=dist.group.WORLD,
                rank=rank, world=world, n_total=n_total)
plan.load(shard.bases, shard.offsets, shard.key_len)
plan.capture(warmup=2)


def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


# graph replays back to back
for _ in range(5):
    plan.run()
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    plan.run()
e1.record()
sync()
graph_ms = e0.elapsed_time(e1) / 20
# single replays, ranks aligned before each
single = []
for _ in range(10):
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); plan.run(); b.record()
    torch.cuda.synchronize()
    single.append(a.elapsed_time(b))
# eager passes with marks
acc = {}
order = []
for it in range(12):
    sync()
    plan._marks = []
    plan.enqueue()
    torch.cuda.synchronize()
    m = plan._marks
    plan._marks = None
    if it < 2:
        continue
    t0 = m[0][1]
    for name, ev in m[1:]:
        acc.setdefault(name, []).append(t0.elapsed_time(ev))
        if name not in order:
            order.append(name)
out = "rank %d: graph back-to-back %.4f ms/pass; single aligned replay %.4f (min %.4f); eager marks (ms since start): " % (
    rank, graph_ms, float(np.mean(single)), float(np.min(single)))
out += ", ".join("%s %.4f" % (nm, float(np.mean(acc[nm]))) for nm in order)
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print("\n".join(gathered))
plan.close()
dist.destroy_process_group()

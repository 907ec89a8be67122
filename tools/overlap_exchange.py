"""Development helper (torchrun, N >= 2): the exchange step after the decode (quantize_mean_sharded_peer) against the exchange
overlapped with the sliced decode (quantize_mean_overlapped_peer) -- bit-equality and time per eager step."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import distributed as dmed

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d = 128, 1 << 24
g = torch.Generator(device="cuda").manual_seed(1 + rank)
X = torch.randn((n, d), generator=g, device="cuda")
mc = world >= 4
kw = dict(n_total=n * world, client0=rank * n, seed=5, multicast=mc)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ref = dmed.quantize_mean_sharded_peer(X, 1, **kw).clone()
t_plain = timed(lambda: dmed.quantize_mean_sharded_peer(X, 1, **kw))
if rank == 0:
    print(f"world={world} after the decode: {t_plain:.3f} ms per eager step", flush=True)
for S in (2, 4, 8):
    got = dmed.quantize_mean_overlapped_peer(X, 1, slices=S, **kw).clone()
    same = torch.equal(got.view(torch.int32), ref.view(torch.int32))
    t = timed(lambda: dmed.quantize_mean_overlapped_peer(X, 1, slices=S, **kw))
    if rank == 0:
        print(f"world={world} overlapped, {S} slices: {t:.3f} ms per eager step, bit-equal {same}", flush=True)
dist.destroy_process_group()

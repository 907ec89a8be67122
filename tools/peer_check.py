"""Development helper (torchrun, N >= 2 GPUs): the peer-memory sum of partial means (csrc/peer_reduce.cu) against NCCL's all-reduce:
bit comparison and CUDA-event times at the bench's vector length."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from dme_b200 import distributed as dmed

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

for d in (1 << 24, 1 << 20, 122626, 1000):
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    part = torch.randn(d, generator=g, device="cuda")
    for mc in (True, False):
        ref = part.clone(); dist.all_reduce(ref)
        try:
            pr = dmed.PeerReduce.get(d, None, mc)
        except Exception as ex:
            if rank == 0: print("PeerReduce unavailable:", repr(ex)[:300], flush=True)
            continue
        pr.buffer().copy_(part)
        out = pr.sum_().clone()
        same = bool(torch.equal(out, ref)) if world == 2 else None
        err = float((out - ref).abs().max())
        def step():
            pr.buffer().copy_(part); pr.sum_()
        def step_nccl():
            ref.copy_(part); dist.all_reduce(ref)
        t_peer, t_nccl = timeit(step), timeit(step_nccl)
        t_copy = timeit(lambda: ref.copy_(part))
        if rank == 0:
            print(f"d={d} world={world} variant={'multimem' if pr.mc else 'peer loads'}: bit-equal to NCCL {same}, max |diff| {err:.3g}; "
                  f"peer {t_peer - t_copy:.4f} ms vs NCCL {t_nccl - t_copy:.4f} ms (copy {t_copy:.4f} subtracted)", flush=True)
dist.destroy_process_group()

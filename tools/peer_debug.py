"""Development helper (torchrun, 2 GPUs): the peer-load variant of dme_peer_sum_slice on random data, with and without host-side
synchronisation around the call."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from dme_b200 import distributed as dmed
for d in (1000, 1 << 20):
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    part = torch.randn(d, generator=g, device="cuda")
    ref = part.clone(); dist.all_reduce(ref)
    for mc in (True, False):
        pr = dmed.PeerReduce.get(d, None, mc)
        for hostsync in (True, False, False):
            pr.buffer().copy_(part)
            if hostsync:
                torch.cuda.synchronize(); dist.barrier()
            out = pr.sum_().clone()
            torch.cuda.synchronize()
            bad = (out != ref)
            nb = int(bad.sum())
            first = int(bad.nonzero()[0]) if nb else -1
            last = int(bad.nonzero()[-1]) if nb else -1
            # is a wrong element equal to one of the un-reduced inputs?
            eq_part = int((out[bad] == part[bad]).sum()) if nb else 0
            print(f"[{rank}] d={d} mc={mc} hostsync={hostsync}: wrong {nb} first {first} last {last} equal-to-own-input {eq_part}", flush=True)
            if hostsync:
                dist.barrier()
dist.destroy_process_group()

"""Multi-GPU: clients are sharded across ranks (one process per GPU); each rank quantizes, decodes and sums its own
clients with the global divisor n_total, and ONE all-reduce (NCCL over NVLink on GPUs) adds the partial means
(clients are independent, the mean is the path's only exchange step).  Philox is keyed by the GLOBAL
client id, so the type vectors do not depend on the number of ranks."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_clients(n_total: int, rank: int, world: int):
    """Contiguous block of clients owned by `rank`: (client0, n_local).  Blocks differ by at most one client."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad rank/world/n_total")
    base, rem = divmod(n_total, world)
    n_local = base + (1 if rank < rem else 0)
    client0 = rank * base + min(rank, rem)
    return client0, n_local


def allreduce_partial_mean(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank partial means in place (each rank already divided by n_total)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def quantize_mean_sharded(x_local, bits_per_dimension=1, *, n_total: int, client0: int, seed: int = 0, mode="unbiased",
                          out=None, group=None, local_fn=None):
    """x_local: this rank's client rows.  Returns the global mean estimate on every rank.

    local_fn(x_local, bits, mode=, seed=, client0=, n_total=, out=) -> partial mean; defaults to the CUDA fused path
    (dme_b200.api.quantize_mean).  Tests on CPU-only boxes inject a stand-in to exercise the sharding/all-reduce logic."""
    if local_fn is None:
        from . import api
        local_fn = api.quantize_mean
    partial = local_fn(x_local, bits_per_dimension, mode=mode, seed=seed, client0=client0, n_total=n_total, out=out)
    return allreduce_partial_mean(partial, group)

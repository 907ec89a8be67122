"""CPU, world_size 2, gloo: the N>1 path's host logic -- client sharding by global id, partial means divided by the
global n, one all-reduce.  The per-rank compute is a stand-in (the oracle): the CUDA path needs a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, d, R, seed, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dme_b200 import _cabi
    from dme_b200.distributed import quantize_mean_sharded, shard_clients
    from oracle import oracle as orc
    L = _cabi.lib()
    rng = np.random.default_rng(123)
    X = rng.standard_normal((n_total, d)).astype(np.float32)                      # every rank builds the same data, uses its shard
    c0, nl = shard_clients(n_total, rank, world)

    def local_fn(x, bits, mode, seed, client0, n_total, out):
        m = orc.m_for(bits, x.shape[1])
        est = np.zeros(x.shape[1], np.float32)
        for j in range(x.shape[0]):
            q = orc.type_unbiased(x[j], m, L.dme_uniform_x(seed, client0 + j))["deq"]
            orc.lib().orc_mean_accumulate(est, q, est.size, n_total)
        return torch.from_numpy(est)

    mean = quantize_mean_sharded(X[c0:c0 + nl], R, n_total=n_total, client0=c0, seed=seed, local_fn=local_fn)
    np.save(os.path.join(out_dir, f"mean_{rank}.npy"), mean.numpy())
    dist.destroy_process_group()


def test_two_rank_sharded_mean(tmp_path):
    from dme_b200 import _cabi
    from oracle import oracle as orc
    n_total, d, R, seed, world = 7, 3000, 1, 99, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_total, d, R, seed, str(tmp_path)), nprocs=world, join=True)
    m0, m1 = np.load(tmp_path / "mean_0.npy"), np.load(tmp_path / "mean_1.npy")
    assert np.array_equal(m0, m1)                                                 # all-reduce: same result on every rank
    L = _cabi.lib()
    X = np.random.default_rng(123).standard_normal((n_total, d)).astype(np.float32)
    m = orc.m_for(R, d)
    ref = orc.mean_of([orc.type_unbiased(X[c], m, L.dme_uniform_x(seed, c))["deq"] for c in range(n_total)])
    # fp32 addition order differs between 1 and 2 ranks: tolerance n * 2^-24 * max|partial|
    assert np.max(np.abs(m0 - ref)) <= n_total * 2.0 ** -24 * max(1e-30, np.max(np.abs(ref))) * 4

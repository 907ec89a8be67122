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


def rotated_quantize_mean_sharded(x_local, bits_per_dimension=1, *, n_total: int, client0: int, seed: int = 0, mode="unbiased",
                                  rotation_seed: int = 123, out=None, group=None, check=False):
    """Config 3 across ranks: every rank rotates its clients with the SHARED diagonal (AS:127-144), quantizes and averages them
    in the rotated domain with the global divisor, ONE all-reduce adds the rotated partial means (dpad floats), and every rank
    applies the single inverse rotation (AS:151-156; linear, so it commutes with the sum)."""
    from . import api
    X, n, d, _ = api._rows(x_local)
    rot = api.rht(X, rotation_seed)
    if rot.dim() == 1:
        rot = rot.unsqueeze(0)
    partial = api.quantize_mean(rot, bits_per_dimension, mode=mode, seed=seed, client0=client0, n_total=n_total, check=check)
    allreduce_partial_mean(partial, group)
    back = api.irht(partial, rotation_seed)[:d]
    if out is not None:
        out.copy_(back)
        return out
    return back.contiguous()


class PeerReduce:
    """The sum of the ranks' partial means by OUR kernels over NVLink peer memory (csrc/peer_reduce.cu) instead of an NCCL
    all-reduce: `buffer(d)` is a symmetric tensor (torch.distributed._symmetric_memory) to use as `out=` of the local fused path;
    `sum_(t)` = barrier, every rank reduces and broadcasts its slice (in-switch multimem reduction when the box has a multicast
    object, peer loads in rank order otherwise), barrier.  One instance per process and vector length."""
    _cache: dict = {}

    def __init__(self, d: int, group=None, multicast=True):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.d = int(d)
        self.dpad = (self.d + 3) // 4 * 4
        self.buf = symm.empty(self.dpad, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        self.mc = int(self.hdl.multicast_ptr) if (multicast and self.hdl.has_multicast_support and int(self.hdl.multicast_ptr) != 0) else 0
        self._C = C

    @classmethod
    def get(cls, d: int, group=None, multicast=True) -> "PeerReduce":
        key = (int(d), id(group), bool(multicast), torch.cuda.current_device())
        if key not in cls._cache:
            cls._cache[key] = PeerReduce(d, group, multicast)
        return cls._cache[key]

    def buffer(self) -> torch.Tensor:
        return self.buf[: self.d]

    def sum_(self) -> torch.Tensor:
        from . import _cabi, api
        C = self._C
        self.hdl.barrier(channel=0)                         # every rank's partial mean is in its buffer
        api._check(_cabi.lib().dme_peer_sum_slice(C.c_void_p(int(self.hdl.buffer_ptrs_dev)), C.c_void_p(self.mc) if self.mc else None,
                                                  int(self.hdl.offset), self.rank, self.world, self.dpad, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.hdl.barrier(channel=1)                         # every slice has been broadcast
        return self.buffer()

    def sum_range_(self, lo: int, count: int) -> None:
        """The same for coordinates [lo, lo + count) only (lo a multiple of 4), on the current stream: a finished slice of the mean
        can be exchanged while the decoder works on the next one (quantize_mean_overlapped_peer)."""
        from . import _cabi, api
        C = self._C
        cnt = min((int(count) + 3) // 4 * 4, self.dpad - int(lo))
        self.hdl.barrier(channel=0)
        api._check(_cabi.lib().dme_peer_sum_slice(C.c_void_p(int(self.hdl.buffer_ptrs_dev)), C.c_void_p(self.mc) if self.mc else None,       # the library adds offset_bytes to either mapping
                                                  int(self.hdl.offset) + 4 * int(lo), self.rank, self.world, cnt,
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.hdl.barrier(channel=1)


def quantize_mean_sharded_peer(x_local, bits_per_dimension=1, *, n_total: int, client0: int, seed: int = 0, mode="unbiased", out=None,
                               group=None, multicast=True, check=False):
    """quantize_mean_sharded with the exchange step over peer memory: the local fused path writes this rank's partial mean straight
    into the symmetric buffer, `PeerReduce.sum_` adds the ranks' buffers in place."""
    from . import api
    X, n, d, _ = api._rows(x_local)
    pr = PeerReduce.get(d, group, multicast)
    api.quantize_mean(X, bits_per_dimension, mode=mode, seed=seed, client0=client0, n_total=n_total, out=pr.buffer(), check=check)
    res = pr.sum_()
    if out is not None:
        out.copy_(res)
        return out
    return res


def quantize_mean_overlapped_peer(x_local, bits_per_dimension=1, *, n_total: int, client0: int, seed: int = 0, mode="unbiased", group=None,
                                  multicast=True, slices: int = 2, check=False):
    """quantize_mean_sharded_peer with the exchange step overlapped: the tile-major decoder writes the symmetric buffer in `slices` runs
    of tiles, each finished run is summed over the ranks (PeerReduce.sum_range_, on a side stream) while the next one is decoded.
    Bit-equal to quantize_mean_sharded_peer.  Measured on 2 B200s at d = 2^24, n = 128 per GPU (tools/overlap_exchange.py): 5.097 ms with
    2, 4 or 8 slices against 5.072 ms for the exchange after the decode -- two device-side barriers per slice and the decoder's tables
    rebuilt per CTA cost what the overlap hides of a 0.13 ms exchange; on 4 B200s (in-switch variant) 5.205-5.213 ms against 5.192 ms;
    bench.py uses the plain call."""
    from . import api
    X, n, d, _ = api._rows(x_local)
    pr = PeerReduce.get(d, group, multicast)
    buf = pr.buffer()
    main = torch.cuda.current_stream()
    comm = _Comm.stream(X.device)

    def exchange(view):
        lo = (view.data_ptr() - buf.data_ptr()) // 4
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            pr.sum_range_(lo, view.numel())

    api.quantize_mean_sliced(X, bits_per_dimension, slices=slices, on_slice=exchange, mode=mode, seed=seed, client0=client0, n_total=n_total,
                             out=buf, check=check)
    main.wait_stream(comm)
    return buf


class _Comm:
    """Side stream on which the per-slice all-reduces of quantize_mean_overlapped are enqueued (one per device)."""
    _streams: dict = {}

    @classmethod
    def stream(cls, dev):
        if dev.index not in cls._streams:
            cls._streams[dev.index] = torch.cuda.Stream(device=dev)
        return cls._streams[dev.index]


def quantize_mean_overlapped(x_local, bits_per_dimension=1, *, n_total: int, client0: int, seed: int = 0, mode="unbiased",
                             out=None, group=None, slices: int = 4, check=True):
    """quantize_mean_sharded with the exchange step overlapped: the decoder is tile-major, so the mean is produced in
    `slices` runs of tiles and each finished slice is all-reduced (NCCL, on a side stream) while the next one is decoded.
    Same result as quantize_mean_sharded (the all-reduce of a slice does not depend on how the vector was cut).
    Measured on 2 B200s at d = 2^24, n = 128 per GPU: 5.24 ms (2 slices) .. 5.30 ms (8) vs 5.22 ms un-overlapped -- the
    0.4 ms decode is too short to hide a 0.2 ms all-reduce behind once NCCL's kernels share the SMs; bench.py uses the
    plain call.  Kept for shapes where the decode dominates (many clients per GPU)."""
    from . import api
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return api.quantize_mean(x_local, bits_per_dimension, mode=mode, seed=seed, client0=client0, n_total=n_total, out=out, check=check)
    dev = x_local.device
    main = torch.cuda.current_stream()
    comm = _Comm.stream(dev)

    def reduce_slice(view):
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ev)
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)

    res = api.quantize_mean_sliced(x_local, bits_per_dimension, slices=slices, on_slice=reduce_slice, mode=mode, seed=seed,
                                   client0=client0, n_total=n_total, out=out, check=check)
    main.wait_stream(comm)
    return res

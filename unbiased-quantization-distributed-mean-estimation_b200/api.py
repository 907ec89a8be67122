"""Batched host API over the C ABI (include/dme_b200.h).  PyTorch supplies device memory and streams only.

Client vectors are rows of X[n, d] (a single vector is accepted as shape (d,)).  `bits_per_dimension`
keeps the reference's meaning: m = int(table[R] * d) (AS:614-623); `m=` overrides it.
"""
from __future__ import annotations

import ctypes as C
import math
import struct
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi

TILE = 1024          # coordinates per tile of the packed code (DME_TILE)
MSG_HEADER = struct.Struct("<8sIIIQQQQfI")      # PackedCodes.to_messages
MODE = {"unbiased": 0, "biased": 1, 0: 0, 1: 1}

# AS:614-620 (the API of the reference: R -> m/d; kept verbatim because it IS the interface)
RATE_TABLE = {
    0.5: 0.08282, 1: 0.21403, 1.5: 0.39443, 2: 0.63752, 2.5: 0.96656, 3: 1.41725, 3.5: 2.04187,
    4: 2.91504, 4.5: 4.14217, 5: 5.87195, 5.5: 8.31416, 6: 11.76507, 6.5: 16.64332, 7: 23.54075,
    7.5: 33.29414, 8: 47.0868, 8.5: 66.59204, 9: 94.17625, 9.5: 133.18596, 10: 188.35383,
}


class DmeError(RuntimeError):
    pass


class _RetryBiased(DmeError):
    """dme_status: a tie-heavy row overflowed the biased mode's fast selection; the call is repeated on the radix path."""


def _biased_retry(fn):
    """Run fn(); when the status word asks for it, run it again with the radix selection (dme_set_biased_path)."""
    try:
        return fn()
    except _RetryBiased:
        L = _cabi.lib()
        _check(L.dme_set_biased_path(1))
        try:
            return fn()
        finally:
            _check(L.dme_set_biased_path(0))


def set_biased_path(path) -> None:
    """Test hook (dme_set_biased_path): "linear" (default) = one linear histogram of the residuals + a candidate list of the
    threshold bin; "radix" = MSB-first radix select over all coordinates (any number of ties)."""
    _check(_cabi.lib().dme_set_biased_path({"linear": 0, "radix": 1}.get(path, path)))


def m_for_rate(bits_per_dimension, d: int) -> int:
    """AS:622-623.  Unknown rates raise KeyError exactly like the reference's dict lookup."""
    return int(RATE_TABLE[bits_per_dimension] * d)


def _check(rc: int):
    if rc == 0:
        return
    msg = _cabi.last_error()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    if rc == -4:
        raise OverflowError(msg)
    if rc == -5:
        raise _RetryBiased(msg)
    raise DmeError(f"libdme_b200 error {rc}: {msg}")


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise DmeError("no CUDA device: this package has no CPU fallback (the oracle under oracle/ is test-only)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _rows(x, device=None):
    """-> (X2 [n, ld] fp32 contiguous on CUDA with ld % 4 == 0 and 16-byte aligned rows, n, d, was_1d)."""
    dev = _device(device)
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    elif not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    was_1d = x.dim() == 1
    if was_1d:
        x = x.unsqueeze(0)
    if x.dim() != 2:
        raise ValueError("expected a vector (d,) or a matrix of client rows (n, d)")
    x = x.to(device=dev, dtype=torch.float32)
    n, d = x.shape
    if d == 0 or n == 0:
        raise ValueError("empty input")
    ok = x.stride(1) == 1 and (n == 1 or x.stride(0) % 4 == 0) and x.data_ptr() % 16 == 0 and (n == 1 or x.stride(0) >= d)
    if not ok:
        ld = (d + 3) // 4 * 4
        buf = torch.zeros((n, ld), dtype=torch.float32, device=dev)
        buf[:, :d] = x
        x = buf[:, :d]
    return x, n, d, was_1d


def _ld(x) -> int:
    n, d = x.shape
    return x.stride(0) if n > 1 else (d + 3) // 4 * 4


class Workspace:
    """Caller-owned scratch for the type-quantizer entry points (grown on demand), one per (device, CUDA stream): the kernels
    of a call keep their ticket counter, look-back records and status word there, so two calls enqueued on different streams
    must not share it."""
    _cache = {}

    def __init__(self, device):
        self.device = device
        self.buf = None

    @classmethod
    def get(cls, device) -> "Workspace":
        key = (device.type, device.index, int(torch.cuda.current_stream(device).cuda_stream))
        if key not in cls._cache:
            cls._cache[key] = Workspace(device)
        return cls._cache[key]

    def ensure(self, n: int, d: int):
        need = int(_cabi.lib().dme_workspace_bytes(n, d))
        if self.buf is None or self.buf.numel() < need + 256:
            self.buf = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        off = (-self.buf.data_ptr()) % 256
        return C.c_void_p(self.buf.data_ptr() + off), need

    def status(self):
        off = (-self.buf.data_ptr()) % 256
        _check(_cabi.lib().dme_status(C.c_void_p(self.buf.data_ptr() + off), C.c_void_p(_stream())))


UNBIASED_PATHS = {"tiles": 0, "literal": 1}


def set_unbiased_path(path) -> None:
    """Test hook (dme_set_unbiased_path): "tiles" = l1_kernel + quantize_warp_kernel (the product path, default),
    "literal" = AS:625-637 as written, one CTA per row (slow; an independent implementation the GPU tests check
    against the same oracle)."""
    _check(_cabi.lib().dme_set_unbiased_path(UNBIASED_PATHS.get(path, path)))


def profile_kernels(fn, *, warm=1):
    """Run fn() with per-kernel CUDA events on the current stream -> [(kernel name, ms), ...] in launch order
    (dme_profile_enable / dme_profile_read; bench.py's roofline leg -- not for use inside a timed region)."""
    L = _cabi.lib()
    for _ in range(warm):
        fn()
    _check(L.dme_profile_enable(1, C.c_void_p(_stream())))
    try:
        fn()
        buf = (C.c_float * 96)()
        k = L.dme_profile_read(buf, 96)
        return [((L.dme_profile_name(i) or b"").decode(), float(buf[i])) for i in range(k)]
    finally:
        L.dme_profile_enable(0, None)


def client_uniforms(seed: int, client0: int, n: int) -> np.ndarray:
    """X_c for clients client0 .. client0+n-1: the same Philox draw the kernels make (AS:634)."""
    L = _cabi.lib()
    return np.array([L.dme_uniform_x(seed, client0 + c) for c in range(n)], dtype=np.float32)


def _resolve_m(d, bits_per_dimension, m):
    return int(m) if m is not None else m_for_rate(bits_per_dimension, d)


def _check_out(out, d, dev):
    """A caller-supplied mean buffer: fp32, contiguous, on the device of the rows, at least d elements, 16-byte aligned."""
    if not (isinstance(out, torch.Tensor) and out.dtype == torch.float32 and out.is_contiguous() and out.device == dev and
            out.numel() >= d and out.data_ptr() % 16 == 0):
        raise ValueError(f"out must be a contiguous 16-byte aligned float32 tensor on {dev} with at least {d} elements")


def _opt_vec(v, n, dev):
    if v is None:
        return None
    if not isinstance(v, torch.Tensor):
        v = np.array([float(e) for e in np.asarray(v, dtype=object).reshape(-1)], dtype=np.float32) if not isinstance(v, np.ndarray) else v.astype(np.float32)
    t = torch.as_tensor(v, dtype=torch.float32).to(dev).reshape(-1).contiguous()
    if t.numel() != n:
        raise ValueError(f"expected {n} injected values, got {t.numel()}")
    return t


def l1_norms(x):
    """AS:624: per-row sum |x| (fp64 accumulate, fp32 result)."""
    X, n, d, _ = _rows(x)
    ws, wsb = Workspace.get(X.device).ensure(n, d)
    out = torch.empty(n, dtype=torch.float32, device=X.device)
    _check(_cabi.lib().dme_l1_norms(_ptr(X), n, d, _ld(X), _ptr(out), ws, wsb, C.c_void_p(_stream())))
    return out


def type_quantize(x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, x_inject=None, l1_inject=None,
                  want=("deq",), check=True):
    """Type_unbiased_quantize / Type_biased_quantize (AS:609-641, AS:669-687) on every row.

    want: any of "deq" (what the reference returns), "k" (int32 magnitudes), "sgn" (uint8 sign bits), "l1".
    Returns a dict of CUDA tensors shaped like the input."""
    X, n, d, was_1d = _rows(x)
    mm = _resolve_m(d, bits_per_dimension, m)
    dev = X.device
    ws, wsb = Workspace.get(dev).ensure(n, d)
    ldo = (d + 3) // 4 * 4
    k = torch.empty((n, ldo), dtype=torch.int32, device=dev) if "k" in want else None
    s = torch.empty((n, ldo), dtype=torch.uint8, device=dev) if "sgn" in want else None
    q = torch.empty((n, ldo), dtype=torch.float32, device=dev) if "deq" in want else None
    l1 = torch.empty(n, dtype=torch.float32, device=dev)
    xi, li = _opt_vec(x_inject, n, dev), _opt_vec(l1_inject, n, dev)
    def call():
        _check(_cabi.lib().dme_type_quantize(_ptr(X), n, d, _ld(X), mm, MODE[mode], _ptr(xi), _ptr(li), seed, client0,
                                             _ptr(k), _ptr(s), _ptr(q), ldo, _ptr(l1), ws, wsb, C.c_void_p(_stream())))
        if check and (k is not None or mode == "biased"):        # biased: the fast selection reports tie-heavy rows here
            Workspace.get(dev).status()
    _biased_retry(call)
    out = {"m": mm, "l1": l1}
    for name, t in (("k", k), ("sgn", s), ("deq", q)):
        if t is not None:
            t = t[:, :d]
            out[name] = t[0] if was_1d else t
    return out


@dataclass
class PackedCodes:
    """Packed code DMEP1 of n client rows (include/dme_b200.h): arena + tile directory + fp32 L1 norms."""
    codes: torch.Tensor     # uint8 arena
    dir: torch.Tensor       # int64 view of uint64 entries: (offset/16) << 8 | width
    l1: torch.Tensor        # float32 [n]
    n: int
    d: int
    m: int
    mode: int

    def tile(self, c: int, t: int):
        """(width, uint32 words) of one tile, on the host (tests / serialisation)."""
        T = (self.d + TILE - 1) // TILE
        e = int(self.dir[c * T + t].item()) & 0xFFFFFFFFFFFFFFFF
        w, off = e & 0xFF, (e >> 8) * 16
        words = self.codes[off: off + 128 * w].cpu().numpy().view(np.uint32)
        return w, words

    def payload_bytes(self) -> int:
        w = (self.dir & 0xFF).sum().item()
        return int(w) * 128

    def bits_per_coordinate(self) -> float:
        """Rate of the code as sent by a client (payload + one width byte per tile + header), in bits per coordinate: to be
        read against the table's R (AS:614-620) -- the fixed-width fields overshoot it."""
        T = (self.d + TILE - 1) // TILE
        return 8.0 * (self.payload_bytes() + self.n * (T + MSG_HEADER.size)) / float(self.n * self.d)

    # ---- wire format: one self-contained message per client (what a client uploads, TU:216-229 ships a dense fp32 vector)
    def to_messages(self, seed: int = 0, client0: int = 0):
        """-> list of n `bytes`: header {magic "DMEP1", version, mode, tile, d, m, client id, seed, L1 fp32, tile count}, one
        width byte per tile, then the tiles' words (little-endian uint32, 128 * width bytes each) in tile order."""
        T = (self.d + TILE - 1) // TILE
        dirh = self.dir.cpu().numpy().view(np.uint64).reshape(self.n, T)
        arena = self.codes.cpu().numpy()
        l1 = self.l1.cpu().numpy()
        msgs = []
        for c in range(self.n):
            w = (dirh[c] & np.uint64(0xFF)).astype(np.uint8)
            off = (dirh[c] >> np.uint64(8)).astype(np.int64) * 16
            parts = [MSG_HEADER.pack(b"DMEP1\0\0\0", 1, self.mode, TILE, self.d, self.m, client0 + c, seed, float(l1[c]), T), w.tobytes()]
            parts += [arena[off[t]: off[t] + 128 * int(w[t])].tobytes() for t in range(T)]
            msgs.append(b"".join(parts))
        return msgs

    @classmethod
    def from_messages(cls, msgs, device=None) -> "PackedCodes":
        """Server side: n client messages -> device arena + directory + norms, ready for `decode_mean`."""
        dev = _device(device)
        heads = [MSG_HEADER.unpack_from(m, 0) for m in msgs]
        magic, ver, mode, tile, d, mm, _, _, _, T = heads[0]
        if magic != b"DMEP1\0\0\0" or ver != 1 or tile != TILE:
            raise ValueError("not a DMEP1 message of this library's tile size")
        n = len(msgs)
        dirh = np.zeros((n, T), np.uint64)
        l1 = np.zeros(n, np.float32)
        chunks, pos = [], 0
        for c, (m, h) in enumerate(zip(msgs, heads)):
            if h[:6] != heads[0][:6] or h[9] != T:
                raise ValueError(f"message {c} does not match the first one (mode / d / m)")
            l1[c] = h[8]
            w = np.frombuffer(m, np.uint8, T, MSG_HEADER.size).astype(np.int64)
            if not np.all(np.isin(w, (2, 4, 8, 16, 32))) or len(m) != MSG_HEADER.size + T + 128 * int(w.sum()):
                raise ValueError(f"message {c} is malformed")
            starts = pos + np.concatenate([[0], np.cumsum(128 * w)[:-1]])
            dirh[c] = ((starts // 16).astype(np.uint64) << np.uint64(8)) | w.astype(np.uint64)
            chunks.append(np.frombuffer(m, np.uint8, 128 * int(w.sum()), MSG_HEADER.size + T))
            pos += 128 * int(w.sum())
        arena = np.concatenate(chunks + [np.zeros(16, np.uint8)])
        return cls(torch.from_numpy(arena).to(dev), torch.from_numpy(dirh.reshape(-1).view(np.int64)).to(dev), torch.from_numpy(l1).to(dev),
                   n, int(d), int(mm), int(mode))


def type_encode(x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, x_inject=None, l1_inject=None,
                codes_bytes=None, check=True) -> PackedCodes:
    """Quantize every row and emit the packed code (sign/magnitude fields, per-tile width)."""
    X, n, d, _ = _rows(x)
    mm = _resolve_m(d, bits_per_dimension, m)
    dev = X.device
    L = _cabi.lib()
    ws, wsb = Workspace.get(dev).ensure(n, d)
    cb = int(codes_bytes) if codes_bytes is not None else int(L.dme_codes_bytes(n, d, mm, 1))
    xi, li = _opt_vec(x_inject, n, dev), _opt_vec(l1_inject, n, dev)
    while True:
        codes = torch.empty(cb, dtype=torch.uint8, device=dev)
        dr = torch.empty(int(L.dme_dir_entries(n, d)), dtype=torch.int64, device=dev)
        l1 = torch.empty(n, dtype=torch.float32, device=dev)
        def call():
            _check(L.dme_type_encode(_ptr(X), n, d, _ld(X), mm, MODE[mode], _ptr(xi), _ptr(li), seed, client0, _ptr(codes), cb,
                                     _ptr(dr), _ptr(l1), ws, wsb, C.c_void_p(_stream())))
            if check:
                Workspace.get(dev).status()
        try:
            _biased_retry(call)
            break
        except MemoryError:
            worst = int(L.dme_codes_bytes(n, d, mm, 0))
            if cb >= worst:
                raise
            cb = worst                       # heavy-tailed input: retry with the worst-case arena
    return PackedCodes(codes, dr, l1, n, d, mm, MODE[mode])


def decode_mean(pc: PackedCodes, *, n_total=None, out=None, accumulate=False, tiles=None, weights=None):
    """Server side: dequantise (AS:640 / AS:687) and average, `est += q / n` in client order (ND:133-147).

    weights [n] (positive integers, e.g. flwr's num_examples): the weighted mean sum_c w_c q_c / sum_c w_c of FedAvg's
    aggregate (TU:260-269; flwr 1.11.1 aggregate()) -- client c's norm is scaled by w_c (exact in fp32 for w_c < 2^24 and a
    norm with free low bits, one extra rounding otherwise) and the divisor becomes sum w.

    tiles=(tile0, count) decodes only coordinates [tile0 * TILE, (tile0 + count) * TILE) of `out` (slices of the mean
    complete in order, so a sharded run can all-reduce one slice while the next is decoded)."""
    dev = pc.codes.device
    if out is None:
        out = torch.empty(pc.d, dtype=torch.float32, device=dev)
        accumulate = False
    else:
        _check_out(out, pc.d, dev)
    nt = pc.n if n_total is None else int(n_total)
    l1 = pc.l1
    if weights is not None:
        w = torch.as_tensor(weights).to(dev).reshape(-1)
        if w.numel() != pc.n or bool((w <= 0).any()):
            raise ValueError("weights: one positive entry per client")
        l1 = (pc.l1 * w.to(torch.float32)).contiguous()
        nt = int(w.sum().item()) if n_total is None else int(n_total)
    T = (pc.d + TILE - 1) // TILE
    t0, cnt = (0, T) if tiles is None else (int(tiles[0]), int(tiles[1]))
    _check(_cabi.lib().dme_decode_mean_tiles(_ptr(pc.codes), _ptr(pc.dir), _ptr(l1), pc.n, pc.d, pc.m, pc.mode, nt, _ptr(out),
                                             int(bool(accumulate)), t0, cnt, C.c_void_p(_stream())))
    return out


def codes_to_host(pc: PackedCodes, pin=True):
    """The device code as the server receives it: host copies (pinned by default) of arena, directory and norms."""
    def h(t):
        o = torch.empty(t.shape, dtype=t.dtype, pin_memory=bool(pin))
        o.copy_(t)
        return o
    return {"codes": h(pc.codes), "dir": h(pc.dir), "l1": h(pc.l1), "n": pc.n, "d": pc.d, "m": pc.m, "mode": pc.mode}


def decode_mean_host(codes_host, *, out_host=None, n_total=None, weights=None, reduce_fn=None):
    """Server-side end to end with HOST buffers: the clients' packed codes (`codes_to_host` / `PackedCodes.from_messages` on the
    host side of a deployment) go to the GPU -- 0.25-0.3 bytes per coordinate at R = 1 instead of the 4 of a dense upload --, one
    decode-mean launch, D2H copy of the mean.  reduce_fn(mean_device) runs before the copy back (all-reduce of a sharded run)."""
    dev = _device()
    pc = PackedCodes(codes_host["codes"].to(dev, non_blocking=True), codes_host["dir"].to(dev, non_blocking=True),
                     codes_host["l1"].to(dev, non_blocking=True), codes_host["n"], codes_host["d"], codes_host["m"], codes_host["mode"])
    mean = decode_mean(pc, n_total=n_total, weights=weights)
    if reduce_fn is not None:
        reduce_fn(mean)
    if out_host is None:
        out_host = torch.empty(pc.d, dtype=torch.float32, pin_memory=True)
    out_host.copy_(mean, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out_host


def quantize_mean_sliced(x, bits_per_dimension=1, *, slices=4, on_slice=None, mode="unbiased", m=None, seed=0, client0=0,
                         n_total=None, x_inject=None, out=None, check=True):
    """quantize_mean with the decode split into `slices` runs of tiles; on_slice(out[lo:hi]) is called right after the
    decode of each slice has been enqueued (the caller's hook for a per-slice all-reduce on another stream).
    Same bits as quantize_mean: every coordinate still sees its clients in order."""
    X, n, d, _ = _rows(x)
    mm = _resolve_m(d, bits_per_dimension, m)
    dev = X.device
    L = _cabi.lib()
    ws, wsb = Workspace.get(dev).ensure(n, d)
    plan = _MeanPlan.get(n, d, mm, dev)
    if out is None:
        out = torch.empty(d, dtype=torch.float32, device=dev)
    else:
        _check_out(out, d, dev)
    xi = _opt_vec(x_inject, n, dev)
    nt = n if n_total is None else int(n_total)
    while True:
        try:
            _check(L.dme_type_encode(_ptr(X), n, d, _ld(X), mm, MODE[mode], _ptr(xi), None, seed, client0, _ptr(plan.codes), plan.cb,
                                     _ptr(plan.dir), _ptr(plan.l1), ws, wsb, C.c_void_p(_stream())))
            if check:
                Workspace.get(dev).status()
            break
        except MemoryError:
            if plan.cb >= int(L.dme_codes_bytes(n, d, mm, 0)):
                raise
            plan.grow_worst_case(n, d, mm, dev)
    T = (d + TILE - 1) // TILE
    S = max(1, min(int(slices), T))
    per = (T + S - 1) // S
    for t0 in range(0, T, per):
        cnt = min(per, T - t0)
        _check(L.dme_decode_mean_tiles(_ptr(plan.codes), _ptr(plan.dir), _ptr(plan.l1), n, d, mm, MODE[mode], nt, _ptr(out), 0, t0, cnt,
                                       C.c_void_p(_stream())))
        if on_slice is not None:
            on_slice(out[t0 * TILE: min(d, (t0 + cnt) * TILE)])
    return out


class _MeanPlan:
    """Reusable buffers of the fused quantize->decode->mean call for one (n, d, m) shape."""
    _cache = {}

    def __init__(self, n, d, m, dev):
        L = _cabi.lib()
        self.cb = int(L.dme_codes_bytes(n, d, m, 1))
        self.codes = torch.empty(self.cb, dtype=torch.uint8, device=dev)
        self.dir = torch.empty(int(L.dme_dir_entries(n, d)), dtype=torch.int64, device=dev)
        self.l1 = torch.empty(n, dtype=torch.float32, device=dev)

    @classmethod
    def get(cls, n, d, m, dev):
        st = int(torch.cuda.current_stream(dev).cuda_stream)
        key = (n, d, m, dev.index, st)
        if key not in cls._cache:
            for k in [k for k in cls._cache if k[3] == dev.index and k[4] == st]:
                del cls._cache[k]           # one plan per (device, stream): the arena can be GiBs; other streams keep theirs
            cls._cache[key] = _MeanPlan(n, d, m, dev)
        return cls._cache[key]

    def grow_worst_case(self, n, d, m, dev):
        self.cb = int(_cabi.lib().dme_codes_bytes(n, d, m, 0))
        self.codes = torch.empty(self.cb, dtype=torch.uint8, device=dev)


def quantize_mean(x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, n_total=None, x_inject=None,
                  out=None, accumulate=False, check=True, plan_rows=None):
    """The north-star path: quantize all rows, pack, decode and average -> mean[d] (fp32, CUDA).

    n_total: divisor of the mean (defaults to the number of rows; pass the global client count when the rows are
    one GPU's shard and the partial results are summed with one all-reduce, see `distributed.py`)."""
    X, n, d, _ = _rows(x)
    mm = _resolve_m(d, bits_per_dimension, m)
    dev = X.device
    L = _cabi.lib()
    ws, wsb = Workspace.get(dev).ensure(n, d)
    pn = n if plan_rows is None else max(n, int(plan_rows))      # buffers sized for pn rows serve any n <= pn
    plan = _MeanPlan.get(pn, d, mm, dev)
    if out is None:
        out = torch.empty(d, dtype=torch.float32, device=dev)
        accumulate = False
    else:
        _check_out(out, d, dev)
    xi = _opt_vec(x_inject, n, dev)
    nt = n if n_total is None else int(n_total)
    def run():
        if accumulate and check:
            # adding to `out` cannot be undone: encode first, read the status word, decode only a complete code (an arena
            # overflow drops tiles).  check=False keeps the single fused call and does NOT detect dropped tiles.
            _check(L.dme_type_encode(_ptr(X), n, d, _ld(X), mm, MODE[mode], _ptr(xi), None, seed, client0, _ptr(plan.codes), plan.cb,
                                     _ptr(plan.dir), _ptr(plan.l1), ws, wsb, C.c_void_p(_stream())))
            Workspace.get(dev).status()
            _check(L.dme_decode_mean(_ptr(plan.codes), _ptr(plan.dir), _ptr(plan.l1), n, d, mm, MODE[mode], nt, _ptr(out), 1,
                                     C.c_void_p(_stream())))
            return out
        _check(L.dme_quantize_mean(_ptr(X), n, d, _ld(X), mm, MODE[mode], _ptr(xi), seed, client0, nt, _ptr(out),
                                   int(bool(accumulate)), _ptr(plan.codes), plan.cb, _ptr(plan.dir), _ptr(plan.l1), ws, wsb,
                                   C.c_void_p(_stream())))
        if check:
            Workspace.get(dev).status()
        return out

    while True:
        try:
            return _biased_retry(run)
        except MemoryError:
            if plan.cb >= int(L.dme_codes_bytes(pn, d, mm, 0)):
                raise
            plan.grow_worst_case(pn, d, mm, dev)


class MeanGraph:
    """The north-star path of ONE shape as a captured CUDA graph: `g = MeanGraph(x, 1); mean = g()` replays
    {uniforms of this round, workspace reset, l1_kernel, quantize_warp_kernel, decode_mean} with one launch.  For short rows and
    few clients (the Flower hook's d = 122 626, n = 100: 30 us of kernels) the eager call is bound by its five launches and the
    host code around them; the replay is not.  `x` is read in place on every replay (refill it between rounds), the result lands
    in `out`.  Randomness: X_c comes from the device-resident round seed (dme_fill_uniforms), which every replay increments, so
    replay k of MeanGraph(seed=s) is bit-identical to quantize_mean(..., seed=s + k); `g(seed=...)` resets it.  A graph cannot
    read the status word back: call `g.status()` when convenient (an exhausted arena is reported there; build the graph with
    worst_case=True to rule it out)."""

    def __init__(self, x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, n_total=None, out=None, accumulate=False,
                 worst_case=False):
        X, n, d, _ = _rows(x)
        self.X, self.n, self.d = X, n, d
        # rows that are not 16-byte aligned (d % 4 != 0 in a dense matrix) were staged into a padded copy: the graph re-stages
        # them on every replay, so that `x` is still read "in place"
        self._src = x if (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 2 and x.data_ptr() != X.data_ptr()) else None
        if self._src is None and not (isinstance(x, torch.Tensor) and x.is_cuda and x.data_ptr() == X.data_ptr()):
            raise ValueError("MeanGraph needs a CUDA fp32 tensor (a vector or a matrix of client rows) that stays alive")
        self.mm = _resolve_m(d, bits_per_dimension, m)
        dev = X.device
        L = _cabi.lib()
        self.out = torch.empty(d, dtype=torch.float32, device=dev) if out is None else out
        _check_out(self.out, d, dev)
        nt = n if n_total is None else int(n_total)
        self._seed_host = torch.empty(1, dtype=torch.int64).pin_memory()
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._xu = torch.empty(n, dtype=torch.float32, device=dev)
        self._cb = int(L.dme_codes_bytes(n, d, self.mm, 0 if worst_case else 1))
        self._codes = torch.empty(self._cb, dtype=torch.uint8, device=dev)
        self._dir = torch.empty(int(L.dme_dir_entries(n, d)), dtype=torch.int64, device=dev)
        self._l1 = torch.empty(n, dtype=torch.float32, device=dev)
        wsb = int(L.dme_workspace_bytes(n, d))
        self._ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
        self._wsp = self._ws.data_ptr() + (-self._ws.data_ptr()) % 256
        self._wsb = wsb
        # one eager call first: per-device kernel attributes are set outside the capture
        quantize_mean(X, bits_per_dimension, mode=mode, m=m, seed=seed, client0=client0, n_total=n_total, out=self.out if not accumulate else None,
                      check=True)
        self.set_seed(seed)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.dme_launch_count()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            st = C.c_void_p(_stream())
            if self._src is not None:
                X.copy_(self._src)
            _check(L.dme_fill_uniforms(_ptr(self._xu), n, _ptr(self._seed_dev), client0, 1, st))
            _check(L.dme_quantize_mean(_ptr(X), n, d, _ld(X), self.mm, MODE[mode], _ptr(self._xu), 0, client0, nt, _ptr(self.out),
                                       int(bool(accumulate)), _ptr(self._codes), self._cb, _ptr(self._dir), _ptr(self._l1),
                                       C.c_void_p(self._wsp), self._wsb, st))
        self.launches = int(L.dme_launch_count() - n0)        # kernels per replay

    def set_seed(self, seed: int) -> None:
        self._seed_host[0] = int(seed)
        self._seed_dev.copy_(self._seed_host, non_blocking=True)

    def __call__(self, seed=None) -> torch.Tensor:
        if seed is not None:
            self.set_seed(seed)
        self.graph.replay()
        _cabi.lib().dme_add_launches(self.launches)
        return self.out

    def status(self) -> None:
        _check(_cabi.lib().dme_status(C.c_void_p(self._wsp), C.c_void_p(_stream())))


class _HostPipe:
    """Two device staging buffers + a copy stream for quantize_mean_host (kept per device and row length)."""
    _cache: dict = {}

    def __init__(self, rows, d, dev):
        self.bufs = [torch.empty((rows, d), dtype=torch.float32, device=dev) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]     # recorded when the compute stream is done with buffer b
        self.mean = torch.empty(d, dtype=torch.float32, device=dev)

    @classmethod
    def get(cls, rows, d, dev):
        key = (dev.index, rows, d)
        pipe = cls._cache.get(key)
        if pipe is None:
            cls._cache.clear()                                    # one plan at a time: the buffers can be GiBs
            pipe = cls._cache[key] = cls(rows, d, dev)
        return pipe


def quantize_mean_host(x_host, bits_per_dimension=1, *, out_host=None, chunk_clients=16, reduce_fn=None, **kw):
    """End-to-end call with HOST buffers: H2D copy of the client rows, fused path, D2H copy of the mean.

    The rows go to the GPU in chunks of `chunk_clients` on a copy stream, double-buffered, while the previous chunk is
    quantized, decoded and ADDED to the running mean on the caller's stream (decode_mean's accumulate mode adds the
    clients in order, so the result is bit-identical to one call over all rows).  Pin x_host for full PCIe speed.
    reduce_fn(mean_device), if given, runs before the D2H copy (the all-reduce of a sharded run)."""
    dev = _device()
    xh = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x_host, dtype=np.float32))
    if xh.dim() == 1:
        xh = xh.reshape(1, -1)
    if xh.dtype != torch.float32 or not xh.is_contiguous():
        xh = xh.to(torch.float32).contiguous()
    n, d = xh.shape
    g = max(1, min(int(chunk_clients), n))
    pipe = _HostPipe.get(g, d, dev)
    main = torch.cuda.current_stream()
    client0 = int(kw.pop("client0", 0))
    n_total = kw.pop("n_total", None)
    n_total = n if n_total is None else int(n_total)
    x_inject = kw.pop("x_inject", None)
    kw.pop("out", None); kw.pop("accumulate", None); kw.pop("plan_rows", None)
    check = kw.pop("check", True)
    starts = list(range(0, n, g))

    def stage(i):                       # H2D copy of chunk i into staging buffer i & 1, on the copy stream
        c0 = starts[i]
        k = min(g, n - c0)
        with torch.cuda.stream(pipe.copy_stream):
            if i >= 2:
                pipe.copy_stream.wait_event(pipe.freed[i & 1])
            pipe.bufs[i & 1][:k].copy_(xh[c0:c0 + k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(pipe.copy_stream)
        return ev

    for attempt in range(2):
        pipe.copy_stream.wait_stream(main)
        copied = stage(0)
        try:
            for i, c0 in enumerate(starts):
                k = min(g, n - c0)
                nxt = stage(i + 1) if i + 1 < len(starts) else None      # in flight while chunk i is quantized
                main.wait_event(copied)
                xi = None if x_inject is None else x_inject[c0:c0 + k]
                # a status read is a sync point of the caller's stream only: the copy stream keeps running
                quantize_mean(pipe.bufs[i & 1][:k], bits_per_dimension, client0=client0 + c0, n_total=n_total, x_inject=xi,
                              out=pipe.mean, accumulate=(i > 0), check=check, plan_rows=g, **kw)
                pipe.freed[i & 1].record(main)
                copied = nxt
            break
        except MemoryError:
            # a later chunk overflowed the code arena after earlier ones were already added: start over with the
            # worst-case arena (heavy-tailed rows only)
            main.synchronize(); pipe.copy_stream.synchronize()
            if attempt == 1:
                raise
            mm = _resolve_m(d, bits_per_dimension, kw.get("m"))
            _MeanPlan.get(g, d, mm, dev).grow_worst_case(g, d, mm, dev)
    if reduce_fn is not None:
        reduce_fn(pipe.mean)
    if out_host is None:
        out_host = torch.empty(d, dtype=torch.float32, pin_memory=True)
    out_host.copy_(pipe.mean, non_blocking=True)
    main.synchronize()
    return out_host


def mean_accumulate(q, *, n_total=None, out=None, accumulate=False):
    """ND:133-147 for dequantised rows: out (+)= sum_c q[c] / n_total, clients in order."""
    Q, n, d, _ = _rows(q)
    if out is None:
        out = torch.empty(d, dtype=torch.float32, device=Q.device)
        accumulate = False
    _check(_cabi.lib().dme_mean_accumulate(_ptr(Q), n, d, _ld(Q), n if n_total is None else int(n_total), _ptr(out),
                                           int(bool(accumulate)), C.c_void_p(_stream())))
    return out


# ------------------------------------------------------------------ rotations
def _pow2_ceil(d: int) -> int:
    return 1 << max(0, math.ceil(math.log2(d))) if d > 1 else 1


def _diag(diag_inject, dpad, dev):
    if diag_inject is None:
        return None
    t = torch.as_tensor(diag_inject, dtype=torch.float32).to(dev).reshape(-1).contiguous()
    if t.numel() != dpad:
        raise ValueError(f"diag_inject must have {dpad} entries")
    return t


def hadamard(v):
    """Hadamard.hadamard (AS:100-115): normalised natural-order WHT of every row; a copy is returned."""
    V, n, d, was_1d = _rows(v)
    if d & (d - 1):
        raise Exception("input numel must be a power of 2")      # AS:103-104, same type and message
    V = V.clone().contiguous()
    _check(_cabi.lib().dme_hadamard(_ptr(V), n, d, _ld(V), C.c_void_p(_stream())))
    return V[0] if was_1d else V


def rademacher(dpad: int, seed: int):
    """The +-1 diagonal the kernels derive from Philox(seed, coordinate) (stands in for AS:117-120)."""
    out = torch.empty(dpad, dtype=torch.float32, device=_device())
    _check(_cabi.lib().dme_rademacher(_ptr(out), dpad, seed, C.c_void_p(_stream())))
    return out


def rht(x, seed=0, *, diag_inject=None, per_row_seed=False):
    """HadamardSender.randomized_hadamard_transform (AS:127-144): pad to a power of two, H(D x)/sqrt(dpad)."""
    X, n, d, was_1d = _rows(x)
    dpad = _pow2_ceil(d)
    out = torch.empty((n, dpad), dtype=torch.float32, device=X.device)
    dg = _diag(diag_inject, dpad, X.device)
    _check(_cabi.lib().dme_rht(_ptr(X), n, d, _ld(X), _ptr(out), dpad, dpad, seed, int(bool(per_row_seed)), _ptr(dg), C.c_void_p(_stream())))
    return out[0] if was_1d else out


def irht(v, seed=0, *, diag_inject=None, per_row_seed=False):
    """HadamardReceiver.randomized_inverse_hadamard_transform (AS:151-156); returns a new tensor."""
    V, n, d, was_1d = _rows(v)
    if d & (d - 1):
        raise Exception("input numel must be a power of 2")
    V = V.clone().contiguous()
    dg = _diag(diag_inject, d, V.device)
    _check(_cabi.lib().dme_irht(_ptr(V), n, d, _ld(V), seed, int(bool(per_row_seed)), _ptr(dg), C.c_void_p(_stream())))
    return V[0] if was_1d else V


def rotated_type_quantize(x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, rotation_seed=123, x_inject=None,
                          diag_inject=None):
    """The rotated type quantizer of BASELINE config 3, per vector.  The reference has no such function: it is
    the composition of its pieces, randomized_hadamard_transform (AS:127-144, one shared rotation_seed as AS:802) ->
    Type_{un}biased_quantize on the dpad rotated coordinates (AS:609-641 / AS:669-687, m = int(table[R] * dpad)) ->
    randomized_inverse_hadamard_transform (AS:151-156) -> [:d].  Returns the dequantised rows (n, d)."""
    X, n, d, was_1d = _rows(x)
    rot = rht(X, rotation_seed, diag_inject=diag_inject)
    if rot.dim() == 1:
        rot = rot.unsqueeze(0)
    q = type_quantize(rot, bits_per_dimension, mode=mode, m=m, seed=seed, client0=client0, x_inject=x_inject, want=("deq",))["deq"]
    if q.dim() == 1:
        q = q.unsqueeze(0)
    back = irht(q, rotation_seed, diag_inject=diag_inject)
    if back.dim() == 1:
        back = back.unsqueeze(0)
    back = back[:, :d]
    return back[0] if was_1d else back


def rotated_quantize_mean(x, bits_per_dimension=1, *, mode="unbiased", m=None, seed=0, client0=0, n_total=None, rotation_seed=123,
                          x_inject=None, diag_inject=None, out=None, check=True):
    """Config 3 as a server would run it: every client rotates with the shared diagonal and sends its packed type code; the
    server decodes and averages IN THE ROTATED DOMAIN and rotates the mean back once (the inverse rotation is linear, so this
    is the mean of the per-client compositions up to the fp32 rounding of the last transform: one inverse transform instead of
    n).  -> mean[d] (fp32, CUDA).  With rows sharded over GPUs, all-reduce the rotated mean before the inverse transform
    (`rotated=True` returns it)."""
    X, n, d, _ = _rows(x)
    rot = rht(X, rotation_seed, diag_inject=diag_inject)
    if rot.dim() == 1:
        rot = rot.unsqueeze(0)
    mean_rot = quantize_mean(rot, bits_per_dimension, mode=mode, m=m, seed=seed, client0=client0, n_total=n_total, x_inject=x_inject, check=check)
    back = irht(mean_rot, rotation_seed, diag_inject=diag_inject)[:d]
    if out is not None:
        out.copy_(back)
        return out
    return back.contiguous()


def pair_transform(v):
    """`fast_walsh_hadamard_transform` as the reference executes it (AS:37-59)."""
    V, n, d, was_1d = _rows(v)
    if d & (d - 1):
        raise Exception("input numel must be a power of 2")
    V = V.clone().contiguous()
    _check(_cabi.lib().dme_pair_transform(_ptr(V), n, d, _ld(V), C.c_void_p(_stream())))
    return V[0] if was_1d else V


# ------------------------------------------------------------------ comparison quantizers
def drive_padded_len(d: int) -> int:
    tot, s0 = 0, 0
    while s0 < d:
        tot += _pow2_ceil(min(2048, d - s0))
        s0 += 2048
    return tot


def drive(x, *, seed=0, dsign_inject=None, compat="reference"):
    """DRIVE_quantize_Hadamard (AS:707-752).  compat="reference" reproduces the reference's transform,
    compat="correct" uses a true Walsh-Hadamard transform."""
    X, n, d, was_1d = _rows(x)
    out = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=X.device)
    ds = None
    if dsign_inject is not None:
        ds = torch.as_tensor(dsign_inject, dtype=torch.float32).to(X.device).reshape(-1).contiguous()
        if ds.numel() != n * drive_padded_len(d):
            raise ValueError("dsign_inject has the wrong length")
    _check(_cabi.lib().dme_drive(_ptr(X), n, d, _ld(X), _ptr(out), out.stride(0), seed, _ptr(ds), {"reference": 0, "correct": 1}[compat],
                                 C.c_void_p(_stream())))
    out = out[:, :d]
    return out[0] if was_1d else out


def _eden_rates(nbits):
    """AS:382-389: integer rates and rates below 1 bit use ONE table (ceil(nbits)); fractional rates above 1 mix the tables of
    floor(nbits) and ceil(nbits) with probability nbits - floor(nbits) for the high one.  -> (low, high, p_high, pdrop)"""
    nb = float(nbits)
    if nb <= 0:
        raise KeyError(nbits)
    if nb == round(nb) or nb < 1:
        b = int(np.ceil(nb))
        if b not in (1, 2):
            raise KeyError(nbits)       # AS:301-320 defines centroids for 1 and 2 bits only
        return b, b, 0.0, (1.0 - nb if nb < 1 else 0.0)
    lo, hi = int(np.floor(nb)), int(np.ceil(nb))
    if lo not in (1, 2) or hi not in (1, 2):
        raise KeyError(nbits)
    return lo, hi, nb - lo, 0.0


def eden_encode(x, nbits=1, *, seed=0, diag_inject=None, norm_inject=None, mask_inject=None):
    """EdenSender.compress (AS:370-390) -> dict(bins uint8 [n,dpad], scale [n], rot [n,dpad]).  nbits in {1, 2}, a rate below 1
    (quantized at 1 bit; the receiver drops coordinates, AS:413-421) or a fractional rate in (1, 2) (AS:352-368: per-coordinate
    Bernoulli mask between the 1- and 2-bit tables; mask_inject uint8 [n, dpad] or Philox keyed by seed * 7 + 13, AS:389).
    Row c is rotated with the diagonal of seed + c (the reference draws a fresh seed per call, AS:800) unless a diagonal is
    injected."""
    lo, hi, p_high, pdrop = _eden_rates(nbits)
    X, n, d, was_1d = _rows(x)
    dpad = _pow2_ceil(d)
    dev = X.device
    rot = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    bins = torch.empty((n, dpad), dtype=torch.uint8, device=dev)
    scale = torch.empty(n, dtype=torch.float32, device=dev)
    dg = _diag(diag_inject, dpad, dev)
    ni = _opt_vec(norm_inject, n, dev)
    mk = None
    if mask_inject is not None:
        mk = torch.as_tensor(mask_inject).to(dev, torch.uint8).reshape(n, dpad).contiguous()
    if lo == hi:
        _check(_cabi.lib().dme_eden_encode(_ptr(X), n, d, _ld(X), dpad, lo, seed, 1, _ptr(dg), _ptr(ni), _ptr(rot), _ptr(bins), _ptr(scale),
                                           C.c_void_p(_stream())))
    else:
        _check(_cabi.lib().dme_eden_encode_frac(_ptr(X), n, d, _ld(X), dpad, lo, hi, float(p_high), _ptr(mk), seed, 1, _ptr(dg), _ptr(ni),
                                                _ptr(rot), _ptr(bins), _ptr(scale), C.c_void_p(_stream())))
    return {"bins": bins, "scale": scale, "rot": rot, "d": d, "dpad": dpad, "nbits": nbits, "seed": seed, "was_1d": was_1d, "mask": mk}


def eden_decode(enc, *, diag_inject=None, drop_inject=None, pdrop=0.0):
    """EdenReceiver.decompress (AS:398-426).  Rates below 1 bit and pdrop > 0 zero round(dpad * p) coordinates per row (AS:413-421:
    a random permutation's head, drawn with torch on the device; drop_inject uint8 [n, dpad] replaces the draw) and divide the
    rest by 1 - p."""
    bins, scale = enc["bins"], enc["scale"]
    n, dpad = bins.shape
    d = enc["d"]
    dev = bins.device
    lo, hi, p_high, p0 = _eden_rates(enc["nbits"])
    p = p0 + (1.0 - p0) * float(pdrop) if pdrop > 0 else p0                               # AS:408-411
    work = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    out = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=dev)
    dg = _diag(diag_inject, dpad, dev)
    if lo == hi and p == 0:
        _check(_cabi.lib().dme_eden_decode(_ptr(bins), _ptr(scale), n, d, dpad, lo, enc["seed"], 1, _ptr(dg), _ptr(work), _ptr(out),
                                           out.stride(0), C.c_void_p(_stream())))
    else:
        drop = None
        if p > 0:
            if drop_inject is not None:
                drop = torch.as_tensor(drop_inject).to(dev, torch.uint8).reshape(n, dpad).contiguous()
            else:
                k = int(round(dpad * p))                                                   # AS:416
                drop = torch.zeros((n, dpad), dtype=torch.uint8, device=dev)
                for c in range(n):
                    drop[c, torch.randperm(dpad, device=dev)[:k]] = 1
        _check(_cabi.lib().dme_eden_decode_frac(_ptr(bins), _ptr(scale), n, d, dpad, lo, hi, float(p_high), _ptr(enc.get("mask")), _ptr(drop),
                                                float(1.0 - p), enc["seed"], 1, _ptr(dg), _ptr(work), _ptr(out), out.stride(0),
                                                C.c_void_p(_stream())))
    out = out[:, :d]
    return out[0] if enc.get("was_1d") else out


def eden(x, nbits=1, *, seed=0, diag_inject=None):
    return eden_decode(eden_encode(x, nbits, seed=seed, diag_inject=diag_inject), diag_inject=diag_inject)


def kashin_padded_dim(dim: int, pad_threshold: float = 0.85) -> int:
    """KashinSender.kashin_padded_dim (AS:203-211)."""
    if dim & (dim - 1):
        p = 1 << int(math.ceil(math.log2(dim)))
        return 2 * p if dim / p > pad_threshold else p
    return 2 * dim


def kashin(x, bits_per_dimension=1, *, seed=0, rotation_seed=123, eta=0.9, delta=1.0, pad_threshold=0.85, niters=3, err=1e-6,
           diag_inject=None, m0_inject=None, u_inject=None, want_parts=False):
    """Kashin_quantize on every row (AS:834-854): Kashin frame coefficients (KashinSender.kashin_coefficients AS:213-239: niters
    rounds of randomized Hadamard transform -> clamp to +-M -> accumulate -> residual, M *= eta) -> min/max stochastic
    quantization of the coefficients (StochasticQuantizationSender AS:67-83, "standard" step, nlevels = 2^bits) -> receiver
    (AS:90, AS:262-267): min + bins * step, inverse transform, [:dim].
    The transforms are the FWHT kernels (all rows in one launch, shared diagonal of rotation_seed); the clamp / residual / rounding
    steps are elementwise device ops.  No host synchronisation: the reference's early exit `if err < 1e-6: break` (AS:236-238)
    is a per-row device flag that freezes the row's coefficients.
    seed: an int (row c uses seed + c) or one seed per row.
    m0_inject [n]: the initial M (= ||x||_2 / sqrt(delta * pdim), an fp32 reduction without a defined order in the reference);
    u_inject [n, pdim]: the uniforms of the Bernoulli draw of AS:81; diag_inject [pdim]: the rotation diagonal."""
    X, n, d, was_1d = _rows(x)
    dev = X.device
    X = X[:, :d]
    pdim = kashin_padded_dim(d, pad_threshold)
    coeff = torch.zeros((n, pdim), dtype=torch.float32, device=dev)
    resid = X.clone()
    if m0_inject is not None:
        M = _opt_vec(m0_inject, n, dev).reshape(n, 1)
    else:
        # tensor / tensor: torch's CUDA division by a SCALAR multiplies by its reciprocal, which is not the IEEE quotient
        M = torch.linalg.vector_norm(X, dim=1, keepdim=True) / torch.full((n, 1), float(np.float32(np.sqrt(delta * pdim))), device=dev)   # AS:221
    done = torch.zeros((n, 1), dtype=torch.bool, device=dev)
    padded = torch.zeros((n, pdim), dtype=torch.float32, device=dev)
    for i in range(int(niters)):
        padded.zero_()
        padded[:, :d] = resid                                                                   # AS:223-224
        b = rht(padded, rotation_seed, diag_inject=diag_inject).reshape(n, pdim)                # AS:225
        b_hat = torch.maximum(torch.minimum(b, M), -M)                                          # AS:228 (clamp(b, -M, M))
        coeff = torch.where(done, coeff, coeff + b_hat)                                         # AS:229
        if i < niters - 1:
            back = irht(b_hat, rotation_seed, diag_inject=diag_inject).reshape(n, pdim)         # AS:232
            resid = torch.where(done, resid, resid - back[:, :d])                               # AS:233
            M = torch.where(done, M, M * np.float32(eta))                                       # AS:234
        rec = irht(coeff, rotation_seed, diag_inject=diag_inject).reshape(n, pdim)[:, :d]
        e = torch.linalg.vector_norm(X - rec, dim=1, keepdim=True) / torch.linalg.vector_norm(resid, dim=1, keepdim=True)   # AS:236
        done = done | (e < err)                                                                 # AS:237-238
    # StochasticQuantizationSender.compress (AS:67-83) / Receiver.decompress (AS:90), per row
    nlevels = np.float32(2 ** bits_per_dimension)
    vmin = coeff.min(dim=1, keepdim=True).values
    vmax = coeff.max(dim=1, keepdim=True).values
    step = (vmax - vmin) / torch.full((n, 1), float(nlevels - np.float32(1)), device=dev)       # AS:72 (IEEE quotient, see above)
    r = (coeff - vmin) / step                                                                   # AS:80
    fl = torch.floor(r)
    if u_inject is not None:
        u = torch.as_tensor(u_inject, dtype=torch.float32).to(dev).reshape(n, pdim)
    else:
        # AS:67, AS:841: the Bernoulli stream of a vector is a function of its `seed` alone -- rows with equal seeds share their
        # uniforms (the reference draws the seed from 100 values, so that happens and correlates the rounding errors)
        seeds = [int(seed) + c for c in range(n)] if np.isscalar(seed) else [int(v) for v in seed]
        if len(seeds) != n:
            raise ValueError(f"expected {n} seeds")
        u = torch.empty((n, pdim), dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev)
        for sd in sorted(set(seeds)):
            g.manual_seed(sd & 0x7FFFFFFFFFFFFFFF)
            row = torch.rand(pdim, dtype=torch.float32, device=dev, generator=g)
            for c in [c for c in range(n) if seeds[c] == sd]:
                u[c] = row
    bins = fl + (u < (r - fl)).to(torch.float32)                                                # AS:81
    deq = vmin + bins * step                                                                    # AS:90
    out = irht(deq, rotation_seed, diag_inject=diag_inject).reshape(n, pdim)[:, :d]             # AS:267
    out = out[0] if was_1d else out
    if want_parts:
        return {"out": out, "coeff": coeff, "bins": bins, "min": vmin.reshape(-1), "step": step.reshape(-1), "pdim": pdim}
    return out


def quicfl_decode(Xq, h, d, recv_table, scale, *, exact_mask=None, exact_vals=None, rotation_seed=123, diag_inject=None):
    """QuicFLReceiver.decompress (AS:526-535) for rows Xq[n, dpad]."""
    dev = _device()
    Xq = torch.as_tensor(Xq).to(dev, torch.int32)
    was_1d = Xq.dim() == 1
    if was_1d:
        Xq = Xq.unsqueeze(0)
    Xq = Xq.contiguous()
    n, dpad = Xq.shape
    h = torch.as_tensor(h).to(dev, torch.int32).reshape(n, dpad).contiguous()
    tab = torch.as_tensor(recv_table, dtype=torch.float32).to(dev).contiguous()
    h_len = tab.shape[-1]
    sc = _opt_vec(scale, n, dev)
    em = ev = eo = None
    if exact_mask is not None:
        em = torch.as_tensor(exact_mask).to(dev, torch.uint8).reshape(n, dpad).contiguous()
        cnt = em.sum(dim=1, dtype=torch.int64)
        eo = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        eo[1:] = torch.cumsum(cnt, 0)
        ev = torch.as_tensor(exact_vals, dtype=torch.float32).to(dev).reshape(-1).contiguous()
    work = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    out = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=dev)
    dg = _diag(diag_inject, dpad, dev)
    _check(_cabi.lib().dme_quicfl_decode(_ptr(Xq), _ptr(h), n, d, dpad, h_len, _ptr(tab), tab.numel(), _ptr(em), _ptr(ev), _ptr(eo),
                                         _ptr(sc), rotation_seed, _ptr(dg), _ptr(work), _ptr(out), out.stride(0), C.c_void_p(_stream())))
    out = out[:, :d]
    return out[0] if was_1d else out


_quic_dev: dict = {}


def _quicfl_tables(nbits, dev, prefix=None):
    """Receiver and derived sender tables of one rate on the device (quicfl_tables.py), cached."""
    from . import quicfl_tables as quicfl
    key = (int(nbits), prefix, dev.index)
    if key not in _quic_dev:
        t = quicfl.tables_for(nbits, prefix)
        _quic_dev[key] = dict(t, recv_dev=torch.from_numpy(np.ascontiguousarray(t["recv"])).to(dev),
                              send_X_dev=torch.from_numpy(np.ascontiguousarray(t["send_X"])).to(dev),
                              send_p_dev=torch.from_numpy(np.ascontiguousarray(t["send_p"])).to(dev))
    return _quic_dev[key]


def quicfl_encode(x, nbits=1, *, seed=0, client0=0, rotation_seed=123, diag_inject=None, tables=None):
    """QuicFLSender.compress (AS:455-503) for the rows of x -> dict {X, h (int32 [n, dpad]), exact_mask (uint8), exact_dense (fp32: the
    exactly sent values at their positions), scale [n], nbits, d, dpad}.  The sender tables are derived from the receiver table
    (dme_b200/quicfl_tables.py; the reference does not ship them); `tables` = a reference-style tables directory, default the packaged copy."""
    X, n, d, was_1d = _rows(x)
    dev = X.device
    from . import quicfl_tables as quicfl
    t = _quicfl_tables(nbits, dev, tables)
    dpad = max(_pow2_ceil(d), 4)
    rot = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    Xq = torch.empty((n, dpad), dtype=torch.int32, device=dev)
    h = torch.empty((n, dpad), dtype=torch.int32, device=dev)
    em = torch.empty((n, dpad), dtype=torch.uint8, device=dev)
    ed = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    sc = torch.empty(n, dtype=torch.float32, device=dev)
    dg = _diag(diag_inject, dpad, dev)
    _check(_cabi.lib().dme_quicfl_encode(_ptr(X), n, d, _ld(X), dpad, t["h_len"], t["x_len"], float(np.float32(t["delta"])),
                                         float(np.float32(quicfl.EXACT_THRESHOLD)), _ptr(t["send_X_dev"]), _ptr(t["send_p_dev"]), seed, client0,
                                         rotation_seed, _ptr(dg), _ptr(rot), _ptr(Xq), _ptr(h), _ptr(em), _ptr(ed), _ptr(sc), C.c_void_p(_stream())))
    return {"X": Xq, "h": h, "exact_mask": em, "exact_dense": ed, "scale": sc, "nbits": int(nbits), "d": d, "dpad": dpad, "was_1d": was_1d,
            "rotation_seed": rotation_seed, "tables": tables}


def quicfl_decode_dense(enc, *, diag_inject=None):
    """QuicFLReceiver.decompress (AS:526-535) of quicfl_encode's output (exact values in place, no compaction, no host sync)."""
    dev = enc["X"].device
    t = _quicfl_tables(enc["nbits"], dev, enc.get("tables"))
    n, dpad, d = enc["X"].shape[0], enc["dpad"], enc["d"]
    work = torch.empty((n, dpad), dtype=torch.float32, device=dev)
    out = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=dev)
    dg = _diag(diag_inject, dpad, dev)
    tab = t["recv_dev"]
    _check(_cabi.lib().dme_quicfl_decode(_ptr(enc["X"]), _ptr(enc["h"]), n, d, dpad, t["h_len"], _ptr(tab), tab.numel(), _ptr(enc["exact_mask"]),
                                         _ptr(enc["exact_dense"]), None, _ptr(enc["scale"]), enc["rotation_seed"], _ptr(dg), _ptr(work), _ptr(out),
                                         out.stride(0), C.c_void_p(_stream())))
    out = out[:, :d]
    return out[0] if enc["was_1d"] else out


def quicfl(x, nbits=1, *, seed=0, client0=0, rotation_seed=123, tables=None):
    """QUICFL_quantize (AS:814-832): sender + receiver for every row of x."""
    return quicfl_decode_dense(quicfl_encode(x, nbits, seed=seed, client0=client0, rotation_seed=rotation_seed, tables=tables))


def scalar_quantize(x, bits_per_dimension=1, *, seed=0, client0=0, u_inject=None):
    """Scalar_quantize (AS:755-790)."""
    X, n, d, was_1d = _rows(x)
    out = torch.empty((n, (d + 3) // 4 * 4), dtype=torch.float32, device=X.device)
    u = None
    if u_inject is not None:
        u = torch.as_tensor(u_inject, dtype=torch.float32).to(X.device).reshape(n, d).contiguous()
    _check(_cabi.lib().dme_scalar_quantize(_ptr(X), n, d, _ld(X), float(2 ** bits_per_dimension - 1), seed, client0, _ptr(u), _ptr(out),
                                           out.stride(0), C.c_void_p(_stream())))
    out = out[:, :d]
    return out[0] if was_1d else out

"""ctypes/numpy front-end of the CPU oracle (oracle/dme_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (and fails loudly without its CUDA library).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdme_oracle.so")

# R -> m/d table of the reference API (AS:614-620); duplicated here on purpose: the oracle must not
# import the product package.
RATE_TABLE = {
    0.5: 0.08282, 1: 0.21403, 1.5: 0.39443, 2: 0.63752, 2.5: 0.96656, 3: 1.41725, 3.5: 2.04187,
    4: 2.91504, 4.5: 4.14217, 5: 5.87195, 5.5: 8.31416, 6: 11.76507, 6.5: 16.64332, 7: 23.54075,
    7.5: 33.29414, 8: 47.0868, 8.5: 66.59204, 9: 94.17625, 9.5: 133.18596, 10: 188.35383,
}
TILE = 1024


def m_for(bits_per_dimension, d: int) -> int:
    """AS:622-623: m = int(table[R] * d)."""
    return int(RATE_TABLE[bits_per_dimension] * d)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dme_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_l1_f64.restype = C.c_double
        L.orc_l1_f64.argtypes = [_f32p, C.c_int64]
        L.orc_type_unbiased.restype = C.c_float
        L.orc_type_unbiased.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_type_biased.restype = C.c_float
        L.orc_type_biased.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_hadamard.restype = C.c_int
        L.orc_hadamard.argtypes = [_f32p, C.c_int64]
        L.orc_rht.restype = C.c_int
        L.orc_rht.argtypes = [_f32p, C.c_int64, _f32p, _f32p, C.c_int64]
        L.orc_irht.restype = C.c_int
        L.orc_irht.argtypes = [_f32p, C.c_int64, _f32p]
        L.orc_pair_transform.restype = None
        L.orc_pair_transform.argtypes = [_f32p, C.c_int64]
        L.orc_drive.restype = None
        L.orc_drive.argtypes = [_f32p, C.c_int64, _f32p, C.c_int, _f32p]
        L.orc_eden_encode.restype = C.c_float
        L.orc_eden_encode.argtypes = [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int, C.c_float, _f32p, _i32p]
        L.orc_eden_decode.restype = None
        L.orc_eden_decode.argtypes = [_i32p, C.c_int64, C.c_int64, _f32p, C.c_int, C.c_float, _f32p, _f32p]
        L.orc_quicfl_decode.restype = None
        L.orc_quicfl_decode.argtypes = [_i32p, _i32p, C.c_int64, C.c_int64, C.c_int, _f32p, C.c_void_p, C.c_void_p,
                                        C.c_float, _f32p, _f32p, _f32p]
        L.orc_scalar.restype = None
        L.orc_scalar.argtypes = [_f32p, C.c_int64, C.c_float, _f32p, _f32p]
        L.orc_mean_accumulate.restype = None
        L.orc_mean_accumulate.argtypes = [_f32p, _f32p, C.c_int64, C.c_int64]
        L.orc_tile_width.restype = C.c_int
        L.orc_tile_width.argtypes = [_i64p, C.c_int64]
        L.orc_pack_tile.restype = None
        L.orc_pack_tile.argtypes = [_i64p, _u8p, C.c_int64, C.c_int, _u32p]
        L.orc_unpack_tile.restype = None
        L.orc_unpack_tile.argtypes = [_u32p, C.c_int, C.c_int64, _i64p, _u8p]
        L.orc_deq_value.restype = C.c_float
        L.orc_deq_value.argtypes = [C.c_float, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.orc_quantize_mean_unbiased.restype = C.c_int
        L.orc_quantize_mean_unbiased.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _f32p, _f32p, C.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def l1(x) -> float:
    x = _f32(x)
    return float(lib().orc_l1_f64(x, x.size))


def type_unbiased(x, m: int, X: float, l1_inject=None):
    """-> dict(k int64[d] >=0, sgn uint8[d] (sign bit of x), deq float32[d], L1 float)."""
    x = _f32(x)
    d = x.size
    k = np.empty(d, np.int64); s = np.empty(d, np.uint8); q = np.empty(d, np.float32)
    L1 = lib().orc_type_unbiased(x, d, int(m), np.float32(X), np.float32(np.nan if l1_inject is None else l1_inject),
                                 _ptr(k), _ptr(s), _ptr(q))
    return {"k": k, "sgn": s, "deq": q, "L1": np.float32(L1)}


def type_biased(x, m: int, l1_inject=None):
    x = _f32(x)
    d = x.size
    k = np.empty(d, np.int64); s = np.empty(d, np.uint8); q = np.empty(d, np.float32)
    delta = C.c_int64(0)
    L1 = lib().orc_type_biased(x, d, int(m), np.float32(np.nan if l1_inject is None else l1_inject),
                               _ptr(k), _ptr(s), _ptr(q), C.addressof(delta))
    return {"k": k, "sgn": s, "deq": q, "L1": np.float32(L1), "Delta": int(delta.value)}


def hadamard(v):
    v = _f32(v).copy()
    if lib().orc_hadamard(v, v.size) != 0:
        raise Exception("input numel must be a power of 2")
    return v


def pad_pow2(d: int) -> int:
    p = 1
    while p < d:
        p <<= 1
    return p


def rht(x, diag):
    x = _f32(x); diag = _f32(diag)
    out = np.empty(diag.size, np.float32)
    if lib().orc_rht(x, x.size, diag, out, diag.size) != 0:
        raise Exception("input numel must be a power of 2")
    return out


def irht(v, diag):
    v = _f32(v).copy(); diag = _f32(diag)
    if lib().orc_irht(v, v.size, diag) != 0:
        raise Exception("input numel must be a power of 2")
    return v


def pair_transform(v):
    v = _f32(v).copy()
    lib().orc_pair_transform(v, v.size)
    return v


def drive_padded_len(d: int) -> int:
    tot, s0 = 0, 0
    while s0 < d:
        ln = min(2048, d - s0)
        tot += pad_pow2(ln)
        s0 += 2048
    return tot


def drive(x, dsign, compat: int = 0):
    x = _f32(x); dsign = _f32(dsign)
    assert dsign.size == drive_padded_len(x.size)
    out = np.empty(x.size, np.float32)
    lib().orc_drive(x, x.size, dsign, int(compat), out)
    return out


def eden_encode(x, diag, nbits: int, norm_inject=None):
    x = _f32(x); diag = _f32(diag)
    dpad = diag.size
    v = np.empty(dpad, np.float32); bins = np.empty(dpad, np.int32)
    scale = lib().orc_eden_encode(x, x.size, diag, dpad, int(nbits),
                                  np.float32(np.nan if norm_inject is None else norm_inject), v, bins)
    return {"bins": bins, "scale": np.float32(scale), "rot": v}


def eden_decode(bins, d: int, diag, nbits: int, scale):
    diag = _f32(diag)
    bins = np.ascontiguousarray(bins, np.int32)
    work = np.empty(diag.size, np.float32); out = np.empty(d, np.float32)
    lib().orc_eden_decode(bins, d, diag.size, diag, int(nbits), np.float32(scale), work, out)
    return out


def eden(x, diag, nbits: int):
    e = eden_encode(x, diag, nbits)
    return eden_decode(e["bins"], np.asarray(x).size, diag, nbits, e["scale"])


EDEN_CENT = {1: np.array([-0.7978845608028654, 0.7978845608028654], np.float32),
             2: np.array([-1.5104176087114887, -0.4527800398860679, 0.4527800398860679, 1.5104176087114887], np.float32)}   # AS:303-304


def _eden_bnd(nbits):
    c = EDEN_CENT[nbits]
    return ((c[:-1] + c[1:]) / np.float32(2)).astype(np.float32)                       # AS:311-312


def eden_encode_frac(x, diag, nbits_low: int, nbits_high: int, mask_high, norm_inject=None):
    """EdenSender.compress for a fractional rate (AS:352-368, AS:385-389): bucketize with both tables, take the high-rate bin
    where mask_high is set; scale = ||v||^2 / <centroids, v>.  The dot product accumulates in fp64 like the C port."""
    v = rht(x, diag)                                                                     # AS:378-380
    nrm = np.float32(np.sqrt(np.sum(v.astype(np.float64) ** 2))) if norm_inject is None else np.float32(norm_inject)
    z = (v * np.float32(np.float64(v.size) ** 0.5)) / nrm                                # AS:354
    mask = np.asarray(mask_high).astype(bool)
    bl = np.searchsorted(_eden_bnd(nbits_low), z, side="left").astype(np.int32)          # torch.bucketize, right=False
    bh = np.searchsorted(_eden_bnd(nbits_high), z, side="left").astype(np.int32)
    bins = np.where(mask, bh, bl).astype(np.int32)                                       # AS:363
    cent = np.where(mask, EDEN_CENT[nbits_high][bh], EDEN_CENT[nbits_low][bl]).astype(np.float32)    # AS:364
    scale = np.float32((nrm * nrm) / np.float32(np.sum(cent.astype(np.float64) * v.astype(np.float64))))   # AS:366
    return {"bins": bins, "scale": scale, "rot": v}


def eden_decode_frac(bins, d: int, diag, nbits_low: int, nbits_high: int, mask_high, scale, drop=None, pdrop=0.0):
    """EdenReceiver.decompress (AS:398-426) for a fractional rate or a rate below one bit (drop: the coordinates of AS:416-417)."""
    mask = np.asarray(mask_high).astype(bool)
    bins = np.asarray(bins)
    vec = np.where(mask, EDEN_CENT[nbits_high][np.minimum(bins, EDEN_CENT[nbits_high].size - 1)],
                   EDEN_CENT[nbits_low][np.minimum(bins, EDEN_CENT[nbits_low].size - 1)]).astype(np.float32)   # AS:401-411
    if drop is not None:
        vec = np.where(np.asarray(drop).astype(bool), np.float32(0), vec / np.float32(1.0 - pdrop)).astype(np.float32)   # AS:417-421
    out = irht(vec, diag)                                                                # AS:425
    return (np.float32(scale) * out)[:d].astype(np.float32)                              # AS:426


def kashin_padded_dim(dim: int, pad_threshold: float = 0.85) -> int:
    """AS:203-211."""
    if dim & (dim - 1):
        p = pad_pow2(dim)
        return 2 * p if dim / p > pad_threshold else p
    return 2 * dim


def kashin(x, diag, bits: int, u, m0=None, eta=0.9, delta=1.0, niters=3, err=1e-6):
    """Kashin_quantize (AS:834-854) = kashin_coefficients (AS:213-239) + StochasticQuantizationSender/Receiver (AS:67-90) + the
    inverse transform (AS:262-267), fp32 elementwise like torch.  u: the uniforms behind torch.bernoulli (AS:81): bit = [u < p];
    m0: the initial M (AS:221: an fp32 torch.norm in the reference; fp64-accumulated here when not injected)."""
    x = _f32(x); diag = _f32(diag)
    dim, pdim = x.size, diag.size
    f32 = np.float32
    coeff = np.zeros(pdim, f32)
    resid = x.copy()
    M = f32(m0) if m0 is not None else f32(f32(np.sqrt(np.sum(x.astype(np.float64) ** 2))) / f32(np.sqrt(delta * pdim)))
    for i in range(niters):
        padded = np.zeros(pdim, f32); padded[:dim] = resid
        b = rht(padded, diag)                                                            # AS:225
        b_hat = np.minimum(np.maximum(b, -M), M).astype(f32)                             # AS:228
        coeff = (coeff + b_hat).astype(f32)                                              # AS:229
        if i < niters - 1:
            resid = (resid - irht(b_hat, diag)[:dim]).astype(f32)                        # AS:232-233
            M = f32(M * f32(eta))                                                        # AS:234
        e = np.sqrt(np.sum((x - irht(coeff, diag)[:dim]).astype(np.float64) ** 2)) / np.sqrt(np.sum(resid.astype(np.float64) ** 2))
        if e < err:                                                                      # AS:236-238
            break
    nlevels = f32(2 ** bits)
    vmin, vmax = coeff.min(), coeff.max()
    step = f32((vmax - vmin) / (nlevels - f32(1)))                                       # AS:72
    r = ((coeff - vmin) / step).astype(f32)                                              # AS:80
    fl = np.floor(r).astype(f32)
    bins = (fl + (_f32(u)[:pdim] < (r - fl).astype(f32)).astype(f32)).astype(f32)        # AS:81
    deq = (vmin + (bins * step).astype(f32)).astype(f32)                                 # AS:90
    out = irht(deq, diag)[:dim]                                                          # AS:267
    return {"out": out, "coeff": coeff, "bins": bins, "min": vmin, "step": step}


def quicfl_decode(X, h, d: int, h_len: int, recv_table, exact_mask, exact_vals, scale, diag):
    diag = _f32(diag)
    X = np.ascontiguousarray(X, np.int32); h = np.ascontiguousarray(h, np.int32)
    tab = _f32(recv_table).reshape(-1)
    em = None if exact_mask is None else np.ascontiguousarray(exact_mask, np.uint8)
    ev = None if exact_vals is None else _f32(exact_vals)
    work = np.empty(diag.size, np.float32); out = np.empty(d, np.float32)
    lib().orc_quicfl_decode(X, h, d, diag.size, int(h_len), tab, _ptr(em), _ptr(ev), np.float32(scale), diag, work, out)
    return out


def scalar(x, bits, u):
    x = _f32(x); u = _f32(u)
    out = np.empty(x.size, np.float32)
    lib().orc_scalar(x, x.size, np.float32(2 ** bits - 1), u, out)
    return out


def mean_of(qs):
    """ND:133-147: est += q / n in client order, fp32."""
    n = len(qs)
    est = np.zeros(np.asarray(qs[0]).size, np.float32)
    for q in qs:
        lib().orc_mean_accumulate(est, _f32(q), est.size, n)
    return est


def pack_row(k, sgn):
    """-> list of (width, uint32 words[32*width]) per 1024-coordinate code tile (format DMEP1)."""
    k = np.ascontiguousarray(k, np.int64); sgn = np.ascontiguousarray(sgn, np.uint8)
    tiles = []
    for t0 in range(0, k.size, TILE):
        kk = np.ascontiguousarray(k[t0:t0 + TILE]); ss = np.ascontiguousarray(sgn[t0:t0 + TILE])
        w = lib().orc_tile_width(kk, kk.size)
        if w < 0:
            raise OverflowError("magnitude >= 2^31 cannot be packed")
        words = np.empty(32 * w, np.uint32)
        lib().orc_pack_tile(kk, ss, kk.size, w, words)
        tiles.append((w, words))
    return tiles


def unpack_tile(words, w: int, cnt: int):
    words = np.ascontiguousarray(words, np.uint32)
    k = np.empty(cnt, np.int64); s = np.empty(cnt, np.uint8)
    lib().orc_unpack_tile(words, int(w), cnt, k, s)
    return k, s


def deq_from_code(L1, m: int, k, sgn, biased: bool = False):
    L1 = np.float32(L1); mf = np.float32(m)
    kf = np.asarray(k).astype(np.float32)
    s = np.where(np.asarray(k) == 0, np.float32(0), np.where(np.asarray(sgn) != 0, np.float32(-1), np.float32(1))).astype(np.float32)
    if biased:
        return ((L1 * s) * (kf / mf)).astype(np.float32)
    return (((L1 * s) * kf) / mf).astype(np.float32)


def quantize_mean_unbiased(Xm, m: int, Xs, threads: int = 1):
    Xm = np.ascontiguousarray(Xm, np.float32)
    n, d = Xm.shape
    Xs = _f32(Xs)
    mean = np.empty(d, np.float32)
    rc = lib().orc_quantize_mean_unbiased(Xm, n, d, d, int(m), Xs, mean, int(threads))
    if rc != 0:
        raise MemoryError("oracle baseline buffer")
    return mean

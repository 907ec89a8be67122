"""QUIC-FL tables (AS:429-451, AS:507-521).

The reference ships, per rate, the RECEIVER table `recv_table[2^b, h_len]` and `data.txt` {delta, T, h_len, x_len}; the sender tables
`sender_table_X` / `sender_table_p` that `QuicFLSender.__init__` loads (AS:447-451) are not in its tree, so the reference
cannot run its own `QUICFL_quantize`.  They are, however, determined by the receiver table: QUIC-FL's sender answers, for a grid
value x and the shared randomness h, with an index X such that the receiver's value is unbiased ON AVERAGE OVER h,

    (1/H) sum_h E[ recv_table[X, h] | x, h ] = x,

at minimal variance.  For a fixed x that is a small linear programme whose solution is, by Lagrangian duality, "for every h the X
whose receiver value is nearest to a shifted target x + s(x)", with exactly one h mixing two adjacent indices so that the mean
hits x.  `sender_tables` computes it in closed form (bisection on the shifted target, vectorised over the 10001 grid points) in
the reference's own format: base index X[x, h] and probability p[x, h] of sending X + 1 (AS:486-489).  What pins it: the
constraint itself (tests/test_quicfl_tables.py: bias below 3e-8 at every grid point), and the reference's published aggregate
numbers (CIFAR-10 round NMSE 0.272 / 0.039 / 1.7e-3 at 1 / 2 / 4 bits with 5 clients per round; these tables give
per-vector NMSE 1.47 / 0.215 / 0.0097, i.e. 0.29 / 0.043 / 1.9e-3 for a mean of 5, before the exact tail).

The receiver tables in `quicfl_tables.npz` are the published QUIC-FL tables (data, 4 arrays of 128-256 floats) as the reference
ships them in `*/Codes/tables/`; `load_tables(prefix)` reads a reference-style directory instead.
"""
from __future__ import annotations

import os
from statistics import NormalDist

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BITS = (1, 2, 3, 4)
SR_BITS = {1: 6, 2: 5, 3: 4, 4: 4}                       # AS:430: file names '{bits}_X_{sr_bits}_h_256_q_*'
EXACT_THRESHOLD = NormalDist().inv_cdf(1.0 - 2.0 ** -9)   # AS:475-478: Q^-1(p / 2), p = 2^-8


def load_tables(prefix: str | None = None) -> dict:
    """-> {nbits: {"recv": float32 [2^b, h_len], "delta", "T", "h_len", "x_len"}}.  prefix=None: the package's copy of the published
    tables; otherwise a directory laid out like the reference's `tables/` (recv_table.txt + data.txt per rate)."""
    out = {}
    if prefix is None:
        z = np.load(os.path.join(_HERE, "quicfl_tables.npz"))
        for b in BITS:
            out[b] = {"recv": z[f"recv{b}"].astype(np.float32), "delta": float(z[f"delta{b}"]), "T": float(z[f"T{b}"]),
                      "h_len": int(z[f"h_len{b}"]), "x_len": int(z[f"x_len{b}"])}
        return out
    import ast
    for b in BITS:
        fn = os.path.join(prefix, f"{b}_X_{SR_BITS[b]}_h_256_q_")
        data = ast.literal_eval(open(fn + "data.txt").read())
        recv = np.loadtxt(fn + "recv_table.txt", dtype=np.float64).astype(np.float32).reshape(2 ** b, int(data["h_len"]))
        out[b] = {"recv": recv, "delta": float(data["delta"]), "T": float(data["T"]), "h_len": int(data["h_len"]), "x_len": int(data["x_len"])}
    return out


def sender_tables(recv: np.ndarray, delta: float, x_len: int):
    """Sender tables of one rate from its receiver table: (X int8 [x_len, h_len], p float32 [x_len, h_len], grid float64 [x_len]).
    Grid point i stands for x = (i - (x_len - 1) / 2) * delta (AS:486: index (int_q + half) * h_len + h)."""
    R = np.asarray(recv, dtype=np.float64)
    L, H = R.shape
    if not (np.diff(R, axis=0) >= 0).all():
        raise ValueError("receiver table must be non-decreasing in X for every h")
    half = (x_len - 1) // 2
    xs = (np.arange(x_len) - half) * float(delta)
    mid = 0.5 * (R[1:] + R[:-1])                          # X_h(t) = number of midpoints of column h below the target t
    RT = R.T[None]                                        # [1, H, L]

    def choose(t):
        return (mid[None, :, :] < t[:, None, None]).sum(1)

    def value(X):
        return np.take_along_axis(RT, X[:, :, None], 2)[:, :, 0]

    lo = np.full(x_len, R.min() - 1.0)
    hi = np.full(x_len, R.max() + 1.0)
    for _ in range(80):                                   # mean over h of value(choose(t)) is a non-decreasing step function of t
        t = 0.5 * (lo + hi)
        up = value(choose(t)).mean(1) < xs
        lo = np.where(up, t, lo)
        hi = np.where(up, hi, t)
    Xlo, Xhi = choose(lo), choose(hi)                     # differ only in the column(s) whose midpoint lies in (lo, hi]
    if ((Xhi - Xlo) < 0).any() or ((Xhi - Xlo) > 1).any():
        raise RuntimeError("sender table: non-adjacent switch")
    vlo, vhi = value(Xlo), value(Xhi)
    gain = (vhi - vlo).mean(1)
    frac = np.where(gain > 0, (xs - vlo.mean(1)) / np.where(gain > 0, gain, 1.0), 0.0)
    frac = np.clip(frac, 0.0, 1.0)
    p = (Xhi != Xlo) * frac[:, None]
    return Xlo.astype(np.int8), p.astype(np.float32), xs


_cache: dict = {}


def tables_for(nbits: int, prefix: str | None = None) -> dict:
    """Receiver + derived sender tables of one rate (host arrays, cached)."""
    key = (int(nbits), prefix)
    if key not in _cache:
        if int(nbits) not in BITS:
            raise ValueError("QUIC-FL tables exist for 1, 2, 3 and 4 bits (AS:430)")
        t = dict(load_tables(prefix)[int(nbits)])
        t["send_X"], t["send_p"], t["grid"] = sender_tables(t["recv"], t["delta"], t["x_len"])
        _cache[key] = t
    return _cache[key]

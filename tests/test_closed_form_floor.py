"""CPU: the identity the C-phase of quantize_warp_kernel relies on (csrc/quantize_warp.cu).

AS:636 evaluates t = floor(RN32(RN32(c) - X)) with c the fp64 prefix of the fractional parts.  Inside one binade of the
fp32 prefix, c32 in [2^e + 1, 2^(e+1)), 2 <= e <= 22, with g = 2^(e-23) and a = ceil(X/g - 1/2):
    t = floor(c - Xp)      if a is even,        Xp = g (a - 1/2)
    t = ceil(c - Xp) - 1   if a is odd.
Checked here against the literal double-rounding expression, ties of both roundings included."""
import numpy as np


def _literal(c, X):
    c32 = c.astype(np.float32)                                   # RN32(c)
    y = (c32 - np.float32(X)).astype(np.float32)                 # RN32(c32 - X)
    return np.floor(y).astype(np.int64)


def _closed(c, X, e):
    g = 2.0 ** (e - 23)
    a = np.ceil(X / g - 0.5)
    Xp = g * (a - 0.5)
    u = c - Xp                                                   # exact: Xp is a multiple of 2^(e-24)
    assert np.all(u + Xp == c)
    return np.floor(u).astype(np.int64) if int(a) % 2 == 0 else (np.ceil(u) - 1).astype(np.int64)


def test_closed_form_matches_double_rounding_in_every_binade():
    rng = np.random.default_rng(0)
    total = 0
    for e in range(2, 23):
        g = 2.0 ** (e - 23)
        lo, hi = 2.0 ** e + 1.5, 2.0 ** (e + 1) - 1.0
        xs = [0.0, 2.0 ** -24, 1 - 2.0 ** -24, 0.5, 0.25, 0.75]
        xs += list(rng.integers(0, 1 << 24, 20) / 2.0 ** 24)                       # torch.rand's grid
        xs += [float(np.float32(v)) for v in rng.random(6) * 0.999]                 # arbitrary fp32 in [0, 1)
        xs += [float(((2 * k + 1) * 2.0 ** (e - 24)) % 1.0) for k in rng.integers(0, 1 << 20, 6)]   # X/g - 1/2 integral: tie of the second rounding
        for X in xs:
            X = float(np.float32(X))
            c = rng.uniform(lo, hi, 1500)
            n = rng.integers(int(lo / g) + 1, int(hi / g) - 1, 600).astype(np.float64)
            fl = np.floor(c[:600])
            c = np.concatenate([c, n * g, n * g + g / 2, n * g - g / 2,          # grid points and exact ties of the first rounding
                                np.nextafter(n * g + g / 2, 0), np.nextafter(n * g + g / 2, 1e300),
                                fl + X, fl + X + g / 2, fl + X - g / 2])          # right at the integer boundaries of c - X
            c = c[(c >= lo) & (c <= hi)]
            assert np.array_equal(_literal(c, X), _closed(c, X, e)), (e, X)
            total += c.size
    assert total > 2_000_000


def test_magic_constant_floor_and_fraction_bits():
    """floor(u) sits in the low word of (u + 1.5 * 2^52) rounded down; a non-negative fp32 fraction becomes a double
    by moving its bits (exponent re-biased), which is what the kernel does instead of a conversion."""
    rng = np.random.default_rng(1)
    u = np.concatenate([rng.uniform(-1e6, 8.4e6, 20000), np.arange(-5, 6, dtype=np.float64), np.array([0.999999999, 1.0, 8388607.999])])
    magic = 6755399441055744.0
    # round-down addition: emulate with exact integer arithmetic on the mantissa (ulp of the sum is 1)
    s = np.floor(u) + magic
    lo = (s.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    lo = np.where(lo >= 1 << 31, lo - (1 << 32), lo)
    assert np.array_equal(lo, np.floor(u).astype(np.int64))
    f = rng.random(20000).astype(np.float32)
    f = f[f > 0]
    b = f.view(np.uint32).astype(np.uint64)
    d = (((b >> np.uint64(3)) + np.uint64(0x38000000)) << np.uint64(32) | ((b << np.uint64(29)) & np.uint64(0xFFFFFFFF))).view(np.float64)
    assert np.array_equal(d, f.astype(np.float64))

"""CPU model of the fixed-point scan of csrc/quantize_fx.cu, checked against the oracle (oracle/dme_oracle.c).

The kernel replaces the fp64 prefix of AS:635 by exact integer arithmetic: every m|x|/D (fp32, AS:625-629) is converted
once to 64-bit fixed point with 2^-32 resolution, F = RN(y * 2^32): high word = floor (AS:630), low word = fractional part
(AS:631).  Prefixes of the low words are integer sums (associative: tile / thread boundaries cannot change a result), and
AS:636's t = floor(RN32(RN32(c) - X)) is evaluated
  * in closed form while the prefix stays inside one binade [2^e + 1, 2^(e+1)), 2 <= e <= 22:
    t = floor((P - U_e) / 2^32) with U_e = Xp_e (+ 1 unit when a_e is odd), Xp_e = g (a_e - 1/2), g = 2^(e-23),
    a_e = ceil(X/g - 1/2) -- the kernel only looks at the carries of a 32-bit running sum;
  * literally, on integers (round to 24 significant bits twice), everywhere else.
This file restates both in Python integers and compares type vectors with the oracle."""
import numpy as np
import pytest

from oracle import oracle as orc

FX = 32


def rn24(F: int) -> int:
    """Round a non-negative integer to 24 significant bits, ties to even (fp32 rounding of F * 2^-32)."""
    if F < (1 << 24):
        return F
    s = F.bit_length() - 24
    q, rem, half = F >> s, F & ((1 << s) - 1), 1 << (s - 1)
    if rem > half or (rem == half and (q & 1)):
        q += 1
    return q << s


def t_literal(P: int, Xi: int) -> int:
    """AS:636 on integers: floor(RN32(RN32(P 2^-32) - X)), X = Xi 2^-32."""
    g = rn24(P) - Xi
    g = rn24(g) if g >= 0 else -rn24(-g)
    return g >> FX                      # floor


def binade_offset(Xi: int, e: int):
    """U_e of the closed form, in 2^-32 units (Xi = X 2^32, a multiple of 2^8 on torch.rand's grid)."""
    assert Xi % (1 << 8) == 0
    X24 = Xi >> 8
    a = (X24 + (1 << e) - 1) >> (e + 1)             # ceil(X / g - 1/2), g = 2^(e-23)
    Xp = a * (1 << (e + 9)) - (1 << (e + 8))        # g (a - 1/2)
    return Xp + (a & 1)


def model_type_vector(x, m, X, l1_inject=None, chunk=32):
    """Type vector computed the way the kernel does (threads of `chunk` coordinates)."""
    x = np.asarray(x, np.float32)
    L1 = np.float32(np.sum(np.abs(x.astype(np.float64)))) if l1_inject is None else np.float32(l1_inject)
    D = np.float32(L1 + np.float32(1e-12))
    y = (np.float32(m) * np.abs(x / D)).astype(np.float32)                     # AS:625-629, one rounding per operation
    F = np.rint(y.astype(np.float64) * 2.0 ** FX)                               # exact product, RN to integer
    F = [int(v) for v in F]
    Xi = int(np.float64(np.float32(X)) * 2.0 ** FX)
    assert Xi == np.float64(np.float32(X)) * 2.0 ** FX
    d = len(F)
    k = np.zeros(d, np.int64)
    P = 0
    for i0 in range(0, d, chunk):
        lo = [f & 0xFFFFFFFF for f in F[i0:i0 + chunk]]
        E, En = P, P + sum(lo)
        e = E.bit_length() - 33                      # E in [2^e, 2^(e+1))
        fast = 2 <= e <= 22 and ((E - (3 << 31)) >> (e + 32)) == 1 and ((En + (1 << 32)) >> (e + 32)) == 1
        if fast:
            acc = E - binade_offset(Xi, e)
            tp = acc >> FX
            for j, l in enumerate(lo):
                acc += l
                t = acc >> FX
                k[i0 + j] = (F[i0 + j] >> FX) + (1 if t - tp == 1 else 0)
                tp = t
        else:
            tp = t_literal(E, Xi)
            c = E
            for j, l in enumerate(lo):
                c += l
                t = t_literal(c, Xi)
                k[i0 + j] = (F[i0 + j] >> FX) + (1 if t - tp == 1 else 0)
                tp = t
        P = En
    return k


def test_integer_literal_equals_float_literal():
    rng = np.random.default_rng(3)
    for _ in range(20000):
        e = int(rng.integers(-20, 24))
        c = float(rng.uniform(2.0 ** e, 2.0 ** (e + 1)))
        P = int(round(c * 2.0 ** FX))
        X = float(rng.integers(0, 1 << 24)) / 2.0 ** 24
        c32 = np.float32(P * 2.0 ** -FX) if P < (1 << 53) else np.float32(float(P) * 2.0 ** -FX)
        lit = int(np.floor(np.float32(c32 - np.float32(X))))
        assert t_literal(P, int(X * 2.0 ** FX)) == lit, (P, X)


def test_closed_form_on_integers_in_every_binade():
    rng = np.random.default_rng(4)
    for e in range(2, 23):
        g = 1 << (e + 9)
        for X24 in [0, 1, (1 << 24) - 1, 1 << 23, 1 << 22] + [int(v) for v in rng.integers(0, 1 << 24, 12)] + \
                   [int(((2 * int(q) + 1) << e) % (1 << 24)) for q in rng.integers(0, 1 << 20, 4)]:
            Xi = X24 << 8
            U = binade_offset(Xi, e)
            lo, hi = (1 << (e + 32)) + (3 << 31), (1 << (e + 33)) - (1 << 32)
            Ps = [int(v) for v in rng.integers(lo, hi, 300)]
            n = [int(v) for v in rng.integers(lo // g + 1, hi // g - 1, 100)]
            Ps += [q * g for q in n] + [q * g + g // 2 for q in n] + [q * g + g // 2 - 1 for q in n] + [q * g + g // 2 + 1 for q in n]
            fl = [(p >> FX) << FX for p in Ps[:100]]
            Ps += [f + Xi for f in fl] + [f + Xi + g // 2 for f in fl] + [f + Xi - g // 2 for f in fl] + [f + Xi - g // 2 - 1 for f in fl]
            for P in Ps:
                if lo <= P <= hi:
                    assert (P - U) >> FX == t_literal(P, Xi), (e, X24, P)


CASES = [("gauss", 1), ("gauss", 2), ("gauss", 4), ("lognormal", 1), ("lognormal", 2), ("laplace", 1), ("tiny", 1), ("sparse", 2), ("bernoulli", 1)]


@pytest.mark.parametrize("dist,R", CASES)
def test_model_matches_oracle(dist, R):
    rng = np.random.default_rng(hash((dist, R)) % (1 << 32))
    for d in (777, 4096, 40000, 300000):
        if dist == "gauss": x = rng.standard_normal(d)
        elif dist == "lognormal": x = rng.lognormal(1.0, 2.0, d) * rng.choice([-1.0, 1.0], d)
        elif dist == "laplace": x = rng.laplace(1.0, 2.0, d)
        elif dist == "tiny": x = rng.standard_normal(d) * np.where(rng.random(d) < 0.5, 1e-9, 1.0)
        elif dist == "sparse": x = rng.standard_normal(d) * (rng.random(d) < 0.05)
        else: x = (rng.random(d) < 0.7).astype(np.float64)
        x = x.astype(np.float32)
        m = orc.m_for(R, d)
        for X in (0.0, 2.0 ** -24, 1 - 2.0 ** -24, float(rng.integers(0, 1 << 24)) / 2.0 ** 24):
            ref = orc.type_unbiased(x, m, X)
            k = model_type_vector(x, m, X)
            assert int(k.sum()) == int(ref["k"].sum())
            assert np.array_equal(k, ref["k"]), (dist, R, d, X, int(np.sum(k != ref["k"])))

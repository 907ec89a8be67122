"""CPU: the oracle (oracle/dme_oracle.c) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer/type-vector work and fp32 elementwise chains are bit-exact;
schemes whose ATen reductions have an unspecified fp32 summation order are compared at 2e-6 relative
(stated per test)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _R(v):
    v = float(v)
    return int(v) if v == int(v) else v


def test_unbiased_bit_exact(golden_dir):
    g = _load(golden_dir, "type_quantizers.npz")
    n = int(g["n_unbiased"])
    assert n >= 200
    for i in range(n):
        x, R, X, L1, q = g[f"u{i}_x"], _R(g[f"u{i}_R"]), float(g[f"u{i}_X"]), g[f"u{i}_L1"], g[f"u{i}_q"]
        m = orc.m_for(R, x.size)
        o = orc.type_unbiased(x, m, X, l1_inject=L1)
        assert np.array_equal(o["deq"].view(np.uint32), q.view(np.uint32)), (i, str(g[f"u{i}_name"]), R, X)
        # type vector recovered from the reference's float output
        if L1 > 0 and m > 0:
            k_ref = np.rint(np.abs(q.astype(np.float64)) * m / float(L1)).astype(np.int64)
            assert np.array_equal(o["k"], k_ref), (i, R, X)
            nz = o["k"] != 0
            assert np.array_equal(o["sgn"][nz], (q[nz] < 0).astype(np.uint8))


def test_unbiased_survey_dyadic_table():
    """SURVEY 8(c) hand-checked table: R=4 -> m=23, R=6 -> m=94 on the dyadic vector (L1=8)."""
    x = np.array([0.5, -1.25, 2, 0, -0.75, 3, -0.25, 0.25], np.float32)
    exp4 = {0.0: [1, 4, 5, 0, 2, 9, 1, 1], 0.3: [2, 3, 6, 0, 2, 9, 0, 1], 0.5: [1, 4, 6, 0, 2, 9, 0, 1], 0.99: [1, 4, 5, 0, 2, 9, 1, 1]}
    for X, k in exp4.items():
        o = orc.type_unbiased(x, orc.m_for(4, 8), X)
        assert o["k"].tolist() == k and int(o["k"].sum()) == 23
    exp6 = {0.0: [5, 15, 24, 0, 8, 36, 3, 3], 0.5: [6, 15, 23, 0, 9, 35, 3, 3]}
    for X, k in exp6.items():
        o = orc.type_unbiased(x, orc.m_for(6, 8), X)
        assert o["k"].tolist() == k and int(o["k"].sum()) == 94
    assert orc.type_biased(x, 23)["k"].tolist() == [1, 3, 6, 0, 2, 9, 1, 1]
    assert orc.type_biased(x, 94)["k"].tolist() == [6, 15, 23, 0, 9, 35, 3, 3]


def test_unbiased_own_l1_matches_reference_l1_on_exact_inputs(golden_dir):
    """On dyadic inputs every partial sum is exact, so the fp64-accumulated L1 equals ATen's."""
    g = _load(golden_dir, "type_quantizers.npz")
    hit = 0
    for i in range(int(g["n_unbiased"])):
        if str(g[f"u{i}_name"]).startswith(("dyadic", "bern", "onehot", "zeros")):
            x = g[f"u{i}_x"]
            o = orc.type_unbiased(x, orc.m_for(_R(g[f"u{i}_R"]), x.size), float(g[f"u{i}_X"]))
            assert np.array_equal(o["deq"].view(np.uint32), g[f"u{i}_q"].view(np.uint32))
            hit += 1
    assert hit > 20


def test_biased_matches_reference_up_to_ties(golden_dir):
    g = _load(golden_dir, "type_quantizers.npz")
    n = int(g["n_biased"])
    assert n >= 40
    exact = 0
    for j in range(n):
        x, R, L1, q = g[f"b{j}_x"], _R(g[f"b{j}_R"]), g[f"b{j}_L1"], g[f"b{j}_q"]
        m = orc.m_for(R, x.size)
        o = orc.type_biased(x, m, l1_inject=L1)
        if float(L1) == 0.0:                       # all-zero input: the output is 0 whatever k is
            assert np.array_equal(o["deq"], q); exact += 1
            continue
        # mass and multiset of |k| are tie-independent; positions only differ among equal residuals
        k_ref = np.rint(np.abs(q.astype(np.float64)) * m / max(float(L1), 1e-30)).astype(np.int64)
        assert int(o["k"].sum()) == int(k_ref.sum()), (j, R)
        if np.array_equal(o["deq"].view(np.uint32), q.view(np.uint32)):
            exact += 1
        else:
            name = str(g[f"b{j}_name"])
            assert name.startswith(("bern", "sparse", "zeros", "onehot", "dyadic")), (j, name, R)   # tie-heavy inputs only
            assert np.array_equal(np.sort(o["k"]), np.sort(k_ref))
    assert exact >= n - 12


def test_hadamard_and_pair_transform_bit_exact(golden_dir):
    g = _load(golden_dir, "hadamard.npz")
    for i in range(int(g["n_h"])):
        x = g[f"h{i}_x"]
        assert np.array_equal(orc.hadamard(x).view(np.uint32), g[f"h{i}_y"].view(np.uint32)), i
        assert np.array_equal(orc.pair_transform(x).view(np.uint32), g[f"p{i}_y"].view(np.uint32)), i
    assert orc.pair_transform(np.arange(1, 9)).tolist() == [7, 4, 17, 10, 27, 16, 37, 22]
    assert np.allclose(orc.hadamard(np.arange(1, 9)) * np.sqrt(8), [36, -4, -8, 0, -16, 0, 0, 0], atol=1e-5)
    with pytest.raises(Exception, match="power of 2"):
        orc.hadamard(np.ones(12, np.float32))


def test_randomized_hadamard_bit_exact(golden_dir):
    g = _load(golden_dir, "hadamard.npz")
    for k in range(int(g["n_r"])):
        x, diag = g[f"r{k}_x"], g[f"r{k}_diag"]
        y = orc.rht(x, diag)
        assert np.array_equal(y.view(np.uint32), g[f"r{k}_y"].view(np.uint32)), k
        z = orc.irht(y, diag)
        assert np.array_equal(z.view(np.uint32), g[f"r{k}_z"].view(np.uint32)), k
        assert np.allclose(z[: x.size], x, atol=1e-5)


def test_drive_reference_mode(golden_dir):
    """tolerance 2e-6 relative to max|q|: S uses fp32 norm / abs-sum reductions (ATen order unspecified)."""
    g = _load(golden_dir, "drive.npz")
    for k in range(int(g["n"])):
        x, ds, q = g[f"d{k}_x"], g[f"d{k}_dsign"], g[f"d{k}_q"]
        o = orc.drive(x, ds, compat=0)
        assert np.max(np.abs(o - q)) <= 2e-6 * np.max(np.abs(q)), k
        # compat=1 is real DRIVE: per-vector error ~ (pi/2 - 1)*|x|^2, far below the reference transform's
        o1 = orc.drive(x, ds, compat=1)
        e0 = np.sum((o - x) ** 2) / np.sum(x ** 2); e1 = np.sum((o1 - x) ** 2) / np.sum(x ** 2)
        if x.size >= 2048:
            assert 0.45 < e1 < 0.70 and e0 > 1.5 * e1


def test_eden(golden_dir):
    """bins bit-exact when the reference's fp32 norm is injected; scale/output 2e-6 relative."""
    g = _load(golden_dir, "eden.npz")
    for k in range(int(g["n"])):
        x, diag, nb = g[f"e{k}_x"], g[f"e{k}_diag"], int(g[f"e{k}_nbits"])
        e = orc.eden_encode(x, diag, nb, norm_inject=g[f"e{k}_norm"])
        assert np.array_equal(e["bins"], g[f"e{k}_bins"]), k
        assert abs(float(e["scale"]) - float(g[f"e{k}_scale"])) <= 2e-6 * abs(float(g[f"e{k}_scale"]))
        out = orc.eden_decode(g[f"e{k}_bins"], x.size, diag, nb, g[f"e{k}_scale"])
        assert np.array_equal(out.view(np.uint32), g[f"e{k}_q"].view(np.uint32)), k
        own = orc.eden(x, diag, nb)                                  # own norm: same up to boundary cases
        assert np.max(np.abs(own - g[f"e{k}_q"])) <= 1e-5 * np.max(np.abs(x)) or np.mean(own != g[f"e{k}_q"]) < 0.01


def test_quicfl_receiver_bit_exact(golden_dir):
    g = _load(golden_dir, "quicfl_recv.npz")
    hl = {1: 64, 2: 32, 3: 16, 4: 16}
    for k in range(int(g["n"])):
        nb, d = int(g[f"q{k}_nbits"]), int(g[f"q{k}_d"])
        out = orc.quicfl_decode(g[f"q{k}_X"], g[f"q{k}_h"], d, hl[nb], g[f"table{nb}"], g[f"q{k}_exact"], g[f"q{k}_ev"],
                                g[f"q{k}_scale"], g[f"q{k}_diag"])
        assert np.array_equal(out.view(np.uint32), g[f"q{k}_out"].view(np.uint32)), k


def test_scalar_and_mean(golden_dir):
    g = _load(golden_dir, "scalar_mean.npz")
    for k in range(int(g["n_s"])):
        o = orc.scalar(g[f"s{k}_x"], int(g[f"s{k}_bits"]), g[f"s{k}_u"])
        assert np.array_equal(o.view(np.uint32), g[f"s{k}_q"].view(np.uint32)), k
    Xm, Xs, L1 = g["mean_X"], g["mean_Xs"], g["mean_L1"]
    R = _R(g["mean_R"])
    qs = [orc.type_unbiased(Xm[c], orc.m_for(R, Xm.shape[1]), float(Xs[c]), l1_inject=L1[c])["deq"] for c in range(Xm.shape[0])]
    est = orc.mean_of(qs)
    assert np.array_equal(est.view(np.uint32), g["mean_est"].view(np.uint32))


def test_pack_roundtrip_and_layout():
    rng = np.random.default_rng(3)
    for d, kmax in ((4096, 1), (5000, 3), (70, 200), (9000, 70000), (4096, 2 ** 31 - 1)):
        k = rng.integers(0, kmax + 1, d).astype(np.int64)
        s = rng.integers(0, 2, d).astype(np.uint8)
        tiles = orc.pack_row(k, s)
        TILE = orc.TILE
        assert len(tiles) == (d + TILE - 1) // TILE
        for t, (w, words) in enumerate(tiles):
            cnt = min(TILE, d - t * TILE)
            kk, ss = orc.unpack_tile(words, w, cnt)
            assert np.array_equal(kk, k[t * TILE: t * TILE + cnt]) and np.array_equal(ss, s[t * TILE: t * TILE + cnt])
            assert w in (2, 4, 8, 16, 32) and kk.max(initial=0) < 2 ** (w - 1)
            assert w == 2 or kk.max() >= 2 ** (w // 2 - 1)            # minimal width
    # layout: width 2, coordinate 17 = chunk 1 field 1 -> word index 1, bits [3:2] = sign<<1 | mag
    k = np.zeros(orc.TILE, np.int64); s = np.zeros(orc.TILE, np.uint8); k[17] = 1; s[17] = 1
    (w, words), = orc.pack_row(k, s)
    assert w == 2 and words[1] == 0b1100 and words.sum() == 0b1100
    with pytest.raises(OverflowError):
        orc.pack_row(np.array([2 ** 31], np.int64), np.zeros(1, np.uint8))


def test_eden_fractional_rates(golden_dir):
    """AS:352-368 / AS:401-421: bins bit-exact under the reference's mask and fp32 norm; the decoded vector bit-exact from the
    reference's bins and scale (mask / dropped coordinates recorded in the fixture)."""
    g = _load(golden_dir, "eden_frac_kashin.npz")
    for k in range(int(g["n_frac"])):
        x, diag, nb = g[f"f{k}_x"], g[f"f{k}_diag"], float(g[f"f{k}_nbits"])
        lo, hi = (int(np.floor(nb)), int(np.ceil(nb))) if nb > 1 else (1, 1)
        e = orc.eden_encode_frac(x, diag, lo, hi, g[f"f{k}_mask"], norm_inject=g[f"f{k}_norm"])
        assert np.array_equal(e["bins"], g[f"f{k}_bins"]), k
        assert abs(float(e["scale"]) - float(g[f"f{k}_scale"])) <= 2e-6 * abs(float(g[f"f{k}_scale"]))
        drop = g[f"f{k}_drop"] if nb < 1 else None
        out = orc.eden_decode_frac(g[f"f{k}_bins"], x.size, diag, lo, hi, g[f"f{k}_mask"], g[f"f{k}_scale"], drop=drop, pdrop=max(0.0, 1 - nb))
        assert np.array_equal(out.view(np.uint32), g[f"f{k}_q"].view(np.uint32)), k


def test_kashin(golden_dir):
    """AS:834-854 with the initial M and the Bernoulli uniforms injected: coefficients, bins, min, step and output bit-exact."""
    g = _load(golden_dir, "eden_frac_kashin.npz")
    for k in range(int(g["n_kashin"])):
        x, diag, bits = g[f"k{k}_x"], g[f"k{k}_diag"], int(g[f"k{k}_bits"])
        assert diag.size == orc.kashin_padded_dim(x.size)
        r = orc.kashin(x, diag, bits, g[f"k{k}_u"], m0=g[f"k{k}_m0"])
        assert np.array_equal(r["coeff"].view(np.uint32), g[f"k{k}_coeff"].view(np.uint32)), k
        assert np.array_equal(r["bins"], g[f"k{k}_bins"]) and r["min"] == g[f"k{k}_min"] and r["step"] == g[f"k{k}_step"], k
        assert np.array_equal(r["out"].view(np.uint32), g[f"k{k}_q"].view(np.uint32)), k
        own = orc.kashin(x, diag, bits, g[f"k{k}_u"])                   # own fp64-accumulated norm: same up to clamp-boundary cases
        assert np.max(np.abs(own["out"] - g[f"k{k}_q"])) <= 0.6 * float(g[f"k{k}_step"]) + 1e-6

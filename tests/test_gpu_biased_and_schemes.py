"""GPU parity (through the C ABI): biased type quantizer (Reznik), DRIVE, EDEN, QUIC-FL receiver, scalar SQ.
Bit-exact where the arithmetic is pinned (type vectors, bins, table look-ups, elementwise fp32 chains); 2e-6
relative where the reference's fp32 reductions have no defined order (DRIVE / EDEN scales)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402


@pytest.fixture(scope="module")
def dme():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dme_b200
    return dme_b200


def _R(v):
    v = float(v)
    return int(v) if v == int(v) else v


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a.view(np.uint32)


# ------------------------------------------------------------------ biased
@pytest.fixture(params=["linear", "radix"])
def biased_path(request, dme):
    """Both selections of the mass repair (csrc/reznik.cu): the linear-histogram path (default; tie-heavy rows fall back to the
    radix path through dme_status, inside the API) and the radix select over all coordinates."""
    dme.set_biased_path(request.param)
    yield request.param
    dme.set_biased_path("linear")


def test_golden_biased_vs_oracle_and_reference(dme, golden_dir, biased_path):
    g = np.load(os.path.join(golden_dir, "type_quantizers.npz"))
    n = int(g["n_biased"])
    same_as_ref = 0
    for j in range(n):
        x, R, L1, q = g[f"b{j}_x"], _R(g[f"b{j}_R"]), g[f"b{j}_L1"], g[f"b{j}_q"]
        out = dme.type_quantize(x, R, mode="biased", l1_inject=[L1], want=("deq", "k", "sgn"))
        o = orc.type_biased(x, out["m"], l1_inject=L1)
        assert np.array_equal(out["k"].cpu().numpy(), o["k"].astype(np.int32)), (j, str(g[f"b{j}_name"]), R)
        assert np.array_equal(_u32(out["deq"].cpu().numpy()), _u32(o["deq"])), (j, R)
        same_as_ref += int(np.array_equal(_u32(out["deq"].cpu().numpy()), _u32(q)))
    assert same_as_ref >= n - 12          # the rest are tie-heavy inputs (torch.topk's tie order is unspecified)


@pytest.mark.parametrize("n,d", [(1, 1), (3, 17), (2, 4096), (3, 4097), (4, 65536), (2, 122626), (1, 1 << 20)])
@pytest.mark.parametrize("R", [1, 2, 4])
def test_biased_random_rows(dme, n, d, R, biased_path):
    rng = np.random.default_rng(7 * n + d)
    X = rng.standard_normal((n, d)).astype(np.float32)
    if d > 100:
        X[0, : d // 3] = 0                      # zeros tie with each other
    out = dme.type_quantize(X, R, mode="biased", want=("deq", "k", "sgn", "l1"))
    m = out["m"]
    if m == 0:
        return
    for c in range(n):
        o = orc.type_biased(X[c], m)
        assert float(out["l1"][c]) == float(o["L1"])
        k = out["k"][c].cpu().numpy()
        assert np.array_equal(k, o["k"].astype(np.int32)), (c, d, R, int((k != o["k"]).sum()))
        assert np.array_equal(_u32(out["deq"][c].cpu().numpy()), _u32(o["deq"]))
        if abs(o["Delta"]) <= d:
            assert int(k.sum()) == m             # mass repaired exactly (AS:657-664)


@pytest.mark.parametrize("dist", ["uniform", "bernoulli", "lognormal12", "onehot", "zeros"])
def test_biased_tie_heavy_and_edge_inputs(dme, dist, biased_path):
    rng = np.random.default_rng(3)
    d = 30000
    x = {"uniform": rng.uniform(-1, 1, d), "bernoulli": (rng.random(d) < 0.7).astype(np.float64),
         "lognormal12": rng.lognormal(1, 2, d) * rng.choice([-1, 1], d), "onehot": np.eye(1, d, 5)[0] * 3.0,
         "zeros": np.zeros(d)}[dist].astype(np.float32)
    for R in (1, 2):
        out = dme.type_quantize(x, R, mode="biased", want=("deq", "k"))
        o = orc.type_biased(x, out["m"])
        assert np.array_equal(out["k"].cpu().numpy(), o["k"].astype(np.int32)), (dist, R)
        assert np.array_equal(_u32(out["deq"].cpu().numpy()), _u32(o["deq"])), (dist, R)


@pytest.mark.parametrize("n,d,R,scale", [(2, 1 << 24, 1, 1.0), (3, 1 << 20, 1, 1.0), (2, (1 << 22) + 777, 2, 1.0), (40, 65536, 1, 1.0), (4, 300000, 4, 1.0), (3, 1 << 20, 1, 0.0)])
def test_biased_linear_equals_radix(dme, n, d, R, scale):
    """Long rows: the two selections must agree bit for bit (k, hence everything); scale = 0: heavy-tailed rows (lognormal)."""
    g = torch.Generator(device="cuda").manual_seed(d % 1000 + n)
    X = torch.randn((n, d), generator=g, device="cuda")
    if scale == 0.0:
        X = torch.exp(2.0 * X) * torch.sign(torch.randn((n, d), generator=g, device="cuda"))
    res = {}
    for path in ("linear", "radix"):
        dme.set_biased_path(path)
        out = dme.type_quantize(X, R, mode="biased", want=("k", "deq"))
        res[path] = (out["k"].clone(), out["deq"].clone(), out["m"])
    dme.set_biased_path("linear")
    assert torch.equal(res["linear"][0], res["radix"][0])
    assert torch.equal(res["linear"][1].view(torch.int32), res["radix"][1].view(torch.int32))
    assert (res["linear"][0].sum(dim=1) == res["linear"][2]).all()            # mass repaired exactly


def test_biased_packed_and_mean(dme, biased_path):
    rng = np.random.default_rng(21)
    n, d = 5, 2 * 4096 + 777
    X = rng.standard_normal((n, d)).astype(np.float32)
    for R in (1, 3):
        pc = dme.type_encode(X, R, mode="biased")
        qs = []
        for c in range(n):
            o = orc.type_biased(X[c], pc.m)
            qs.append(o["deq"])
            tiles = orc.pack_row(o["k"], o["sgn"])
            for t, (w, words) in enumerate(tiles):
                gw, gwords = pc.tile(c, t)
                assert gw == w and np.array_equal(gwords, words), (R, c, t)
        mean = dme.decode_mean(pc).cpu().numpy()
        assert np.array_equal(_u32(mean), _u32(orc.mean_of(qs))), R
        fused = dme.quantize_mean(X, R, mode="biased").cpu().numpy()
        assert np.array_equal(_u32(fused), _u32(mean))


# ------------------------------------------------------------------ DRIVE
def test_drive_golden_and_modes(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "drive.npz"))
    for k in range(int(g["n"])):
        x, ds, q = g[f"d{k}_x"], g[f"d{k}_dsign"], g[f"d{k}_q"]
        out = dme.drive(x, dsign_inject=ds, compat="reference").cpu().numpy()
        assert np.max(np.abs(out - q)) <= 2e-6 * np.max(np.abs(q)), k           # tolerance: fp32 scale S (see module doc)
        for compat, cc in (("reference", 0), ("correct", 1)):
            o = orc.drive(x, ds, compat=cc)
            got = dme.drive(x, dsign_inject=ds, compat=compat).cpu().numpy()
            assert np.max(np.abs(got - o)) <= 2e-6 * np.max(np.abs(o)), (k, compat)


def test_drive_philox_batched(dme):
    rng = np.random.default_rng(2)
    X = rng.standard_normal((6, 10000)).astype(np.float32)
    a = dme.drive(X, seed=5, compat="correct").cpu().numpy()
    b = dme.drive(X, seed=5, compat="correct").cpu().numpy()
    assert np.array_equal(a, b)
    err = np.sum((a - X) ** 2, axis=1) / np.sum(X ** 2, axis=1)
    assert np.all((err > 0.45) & (err < 0.70))                                   # real DRIVE: pi/2 - 1 = 0.571
    ref = dme.drive(X, seed=5, compat="reference").cpu().numpy()
    err_ref = np.sum((ref - X) ** 2, axis=1) / np.sum(X ** 2, axis=1)
    assert np.all(err_ref > 1.2)                                                 # the reference's transform


# ------------------------------------------------------------------ EDEN
def test_eden_golden(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "eden.npz"))
    for k in range(int(g["n"])):
        x, diag, nb = g[f"e{k}_x"], g[f"e{k}_diag"], int(g[f"e{k}_nbits"])
        enc = dme.eden_encode(x, nb, diag_inject=diag, norm_inject=[g[f"e{k}_norm"]])
        assert np.array_equal(enc["bins"][0].cpu().numpy().astype(np.int32), g[f"e{k}_bins"]), k
        sc = float(enc["scale"][0])
        assert abs(sc - float(g[f"e{k}_scale"])) <= 2e-6 * abs(float(g[f"e{k}_scale"]))
        enc["scale"] = torch.tensor([float(g[f"e{k}_scale"])], device="cuda")          # decode with the reference's scale: bit-exact
        out = dme.eden_decode(enc, diag_inject=diag).cpu().numpy()
        assert np.array_equal(_u32(out), _u32(g[f"e{k}_q"])), k
        own = dme.eden(x, nb, diag_inject=diag).cpu().numpy()
        o = orc.eden(x, diag, nb)
        assert np.max(np.abs(own - o)) <= 4e-6 * np.max(np.abs(o))


def test_eden_batched_quality(dme):
    rng = np.random.default_rng(4)
    X = rng.standard_normal((8, 50000)).astype(np.float32)
    for nb in (1, 2):
        Y = dme.eden(X, nb, seed=11).cpu().numpy()
        for c in range(X.shape[0]):
            diag = dme.rademacher(1 << 16, seed=11 + c).cpu().numpy()           # row c is rotated with seed + c (AS:800)
            o = orc.eden(X[c], diag, nb)                                         # same Philox diagonal, oracle arithmetic
            assert np.max(np.abs(Y[c] - o)) <= 4e-6 * np.max(np.abs(o)), (nb, c)
        err = np.sum((Y - X) ** 2, axis=1) / np.sum(X ** 2, axis=1)
        assert np.all(err < (0.6 if nb == 1 else 0.2)), (nb, err)
    with pytest.raises(KeyError):
        dme.eden_encode(X, 3)


def test_eden_fractional_golden(dme, golden_dir):
    """AS:352-368 / AS:401-421 under the reference's mask, fp32 norm and dropped coordinates: bins and decoded vector bit-exact."""
    g = np.load(os.path.join(golden_dir, "eden_frac_kashin.npz"))
    for k in range(int(g["n_frac"])):
        x, diag, nb = g[f"f{k}_x"], g[f"f{k}_diag"], float(g[f"f{k}_nbits"])
        enc = dme.eden_encode(x, nb, diag_inject=diag, norm_inject=[g[f"f{k}_norm"]], mask_inject=g[f"f{k}_mask"] if nb > 1 else None)
        assert np.array_equal(enc["bins"][0].cpu().numpy().astype(np.int32), g[f"f{k}_bins"]), k
        assert abs(float(enc["scale"][0]) - float(g[f"f{k}_scale"])) <= 2e-6 * abs(float(g[f"f{k}_scale"]))
        enc["scale"] = torch.tensor([float(g[f"f{k}_scale"])], device="cuda")
        out = dme.eden_decode(enc, diag_inject=diag, drop_inject=g[f"f{k}_drop"] if nb < 1 else None).cpu().numpy()
        assert np.array_equal(_u32(out), _u32(g[f"f{k}_q"])), k


def test_eden_fractional_own_draws(dme):
    """Philox mask (sender and receiver derive the same one) and torch-drawn drops: the error sits between the neighbouring
    integer rates, and a rate below one bit costs the dropped fraction."""
    rng = np.random.default_rng(40)
    X = rng.standard_normal((6, 30000)).astype(np.float32)
    def nmse(Y):
        return float(np.mean(np.sum((Y - X) ** 2, axis=1) / np.sum(X ** 2, axis=1)))
    e1, e15, e2 = (nmse(dme.eden(X, b, seed=3).cpu().numpy()) for b in (1, 1.5, 2))
    assert e2 < e15 < e1 and abs(e15 - 0.5 * (e1 + e2)) < 0.08 * e1
    e05 = nmse(dme.eden(X, 0.5, seed=3).cpu().numpy())
    assert e05 > 1.5 * e1
    for bad in (3, 2.5, 0):
        with pytest.raises(KeyError):
            dme.eden_encode(X, bad)


def test_kashin_golden(dme, golden_dir):
    """Kashin_quantize (AS:834-854) with the rotation diagonal, the initial M and the Bernoulli uniforms of the reference run
    injected: coefficients, bins, min, step and the decoded vector bit-exact; with own norm: same up to clamp-boundary cases."""
    g = np.load(os.path.join(golden_dir, "eden_frac_kashin.npz"))
    for k in range(int(g["n_kashin"])):
        x, diag, bits = g[f"k{k}_x"], g[f"k{k}_diag"], int(g[f"k{k}_bits"])
        r = dme.kashin(x, bits, diag_inject=diag, m0_inject=[g[f"k{k}_m0"]], u_inject=g[f"k{k}_u"], want_parts=True)
        assert r["pdim"] == diag.size
        assert np.array_equal(_u32(r["coeff"][0].cpu().numpy()), _u32(g[f"k{k}_coeff"])), k
        assert np.array_equal(r["bins"][0].cpu().numpy(), g[f"k{k}_bins"]), k
        assert float(r["min"][0]) == float(g[f"k{k}_min"]) and float(r["step"][0]) == float(g[f"k{k}_step"]), k
        assert np.array_equal(_u32(r["out"].cpu().numpy()), _u32(g[f"k{k}_q"])), k
        own = dme.kashin(x, bits, diag_inject=diag, u_inject=g[f"k{k}_u"]).cpu().numpy()
        o = orc.kashin(x, diag, bits, g[f"k{k}_u"])["out"]
        # own fp32 norm: M differs in the last bits, so everything moves by ~1e-7; a rounding that flips moves a coordinate by a step
        assert np.max(np.abs(own - o)) <= 0.6 * float(g[f"k{k}_step"]) + 1e-6 and np.mean(np.abs(own - o) > 1e-4) < 0.01


def test_kashin_batched_matches_per_row_and_quality(dme):
    rng = np.random.default_rng(41)
    X = rng.standard_normal((5, 3000)).astype(np.float32)
    pdim = dme.kashin_padded_dim(3000)
    U = rng.random((5, pdim)).astype(np.float32)
    Y = dme.kashin(X, 3, u_inject=U).cpu().numpy()
    for c in range(5):
        assert np.array_equal(_u32(Y[c]), _u32(dme.kashin(X[c], 3, u_inject=U[c]).cpu().numpy())), c
    err = np.sum((Y - X) ** 2, axis=1) / np.sum(X ** 2, axis=1)
    assert np.all(err < 0.2), err


# ------------------------------------------------------------------ QUIC-FL receiver
def test_quicfl_receiver_golden(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "quicfl_recv.npz"))
    for k in range(int(g["n"])):
        nb, d = int(g[f"q{k}_nbits"]), int(g[f"q{k}_d"])
        out = dme.quicfl_decode(g[f"q{k}_X"], g[f"q{k}_h"], d, g[f"table{nb}"], [g[f"q{k}_scale"]], exact_mask=g[f"q{k}_exact"],
                                exact_vals=g[f"q{k}_ev"], diag_inject=g[f"q{k}_diag"]).cpu().numpy()
        assert np.array_equal(_u32(out), _u32(g[f"q{k}_out"])), k


# ------------------------------------------------------------------ scalar SQ + mean of dequantised rows
def test_scalar_golden_and_philox(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "scalar_mean.npz"))
    for k in range(int(g["n_s"])):
        out = dme.scalar_quantize(g[f"s{k}_x"], int(g[f"s{k}_bits"]), u_inject=g[f"s{k}_u"]).cpu().numpy()
        assert np.array_equal(_u32(out), _u32(g[f"s{k}_q"])), k
    rng = np.random.default_rng(6)
    x = rng.standard_normal(20000).astype(np.float32)
    acc = np.zeros_like(x, dtype=np.float64)
    for s in range(64):
        acc += dme.scalar_quantize(x, 2, seed=s).cpu().numpy()
    assert np.abs(acc / 64 - x).mean() < 0.12                                    # unbiased: the average converges
    const = np.full(100, 3.0, np.float32)
    assert np.array_equal(dme.scalar_quantize(const, 2).cpu().numpy(), const)    # AS:763-765


def test_mean_accumulate(dme):
    rng = np.random.default_rng(8)
    Q = rng.standard_normal((7, 12345)).astype(np.float32)
    got = dme.mean_accumulate(Q).cpu().numpy()
    assert np.array_equal(_u32(got), _u32(orc.mean_of(list(Q))))

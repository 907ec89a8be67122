"""GPU parity: Walsh-Hadamard transform, randomized transform / inverse, pair transform -- bit-exact against the
golden fixtures produced by the reference and against the oracle at larger sizes."""
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


def _eq(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32), np.ascontiguousarray(b, np.float32).view(np.uint32))


def test_golden_hadamard_and_pair(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "hadamard.npz"))
    for i in range(int(g["n_h"])):
        x = g[f"h{i}_x"]
        assert _eq(dme.hadamard(x).cpu().numpy(), g[f"h{i}_y"]), i
        assert _eq(dme.pair_transform(x).cpu().numpy(), g[f"p{i}_y"]), i
    with pytest.raises(Exception, match="power of 2"):
        dme.hadamard(np.ones(12, np.float32))


def test_golden_randomized(dme, golden_dir):
    g = np.load(os.path.join(golden_dir, "hadamard.npz"))
    for k in range(int(g["n_r"])):
        x, diag = g[f"r{k}_x"], g[f"r{k}_diag"]
        y = dme.rht(x, diag_inject=diag)
        assert _eq(y.cpu().numpy(), g[f"r{k}_y"]), k
        z = dme.irht(y, diag_inject=diag)
        assert _eq(z.cpu().numpy(), g[f"r{k}_z"]), k


@pytest.mark.parametrize("logd", [0, 1, 3, 4, 5, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24])
def test_hadamard_sizes_vs_oracle(dme, logd):
    d = 1 << logd
    n = 3 if logd <= 17 else 1
    rng = np.random.default_rng(logd)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Y = dme.hadamard(X).cpu().numpy()
    for c in range(n):
        assert _eq(Y[c], orc.hadamard(X[c])), (logd, c)


def test_rht_batched_philox_roundtrip(dme):
    rng = np.random.default_rng(1)
    n, d = 5, 100000
    X = rng.standard_normal((n, d)).astype(np.float32)
    dpad = 1 << 17
    diag = dme.rademacher(dpad, seed=123).cpu().numpy()
    assert set(np.unique(diag)) == {-1.0, 1.0} and abs(diag.mean()) < 0.02
    Y = dme.rht(X, seed=123)
    assert Y.shape == (n, dpad)
    for c in range(n):
        assert _eq(Y[c].cpu().numpy(), orc.rht(X[c], diag)), c
    Z = dme.irht(Y, seed=123).cpu().numpy()
    assert np.allclose(Z[:, :d], X, atol=2e-5) and np.allclose(Z[:, d:], 0, atol=2e-5)
    # norm preservation (orthonormal transform)
    assert np.allclose(np.linalg.norm(Y.cpu().numpy(), axis=1), np.linalg.norm(X, axis=1), rtol=1e-5)


def test_full_size_involution(dme):
    d = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(d, generator=g, device="cuda")
    y = dme.hadamard(x)
    z = dme.hadamard(y)
    assert torch.allclose(z, x, atol=1e-4)
    assert abs(float(y.norm()) / float(x.norm()) - 1) < 1e-5


@pytest.mark.parametrize("logd", [13, 17, 18, 20, 21, 24])
def test_irht_philox_diagonal_every_pass_structure(dme, logd):
    """The inverse transform applies the Philox diagonal in the LAST pass (contiguous block, narrow strided pass or wide
    strided pass, depending on log d): bit-exact against the oracle driven by the same diagonal."""
    d = 1 << logd
    rng = np.random.default_rng(100 + logd)
    x = rng.standard_normal(d).astype(np.float32)
    diag = dme.rademacher(d, seed=77).cpu().numpy()
    y = dme.rht(x, seed=77)
    assert _eq(y.cpu().numpy(), orc.rht(x, diag)), logd
    z = dme.irht(y, seed=77)
    assert _eq(z.cpu().numpy(), orc.irht(y.cpu().numpy(), diag)), logd


@pytest.mark.parametrize("n,d", [(3, 1000), (2, 5000), (4, 70000), (2, 1 << 20)])
@pytest.mark.parametrize("mode", ["unbiased", "biased"])
def test_rotated_type_quantizer_vs_oracle_composition(dme, n, d, mode):
    """BASELINE config 3: AS:127-144 -> AS:609-641 / AS:669-687 -> AS:151-156, composed from the oracle's pieces with the
    same diagonal and the same uniforms: bit-exact per vector, and bit-exact for the server-side form that averages in the
    rotated domain and rotates the mean back once."""
    if mode == "biased" and d > 100000:
        pytest.skip("the oracle's Reznik sort is slow at this size")
    rng = np.random.default_rng(n * 31 + d)
    X = (rng.standard_normal((n, d)) * np.where(rng.random((n, d)) < 0.02, 30.0, 1.0)).astype(np.float32)      # spiky: rotation matters
    dpad = orc.pad_pow2(d)
    diag = dme.rademacher(dpad, seed=123).cpu().numpy()
    Xs = dme.client_uniforms(seed=9, client0=0, n=n)
    for R in (1, 2):
        m = dme.m_for_rate(R, dpad)
        got = dme.rotated_type_quantize(X, R, mode=mode, seed=9, rotation_seed=123).cpu().numpy()
        rot_q = []
        for c in range(n):
            r = orc.rht(X[c], diag)
            q = orc.type_unbiased(r, m, float(Xs[c]))["deq"] if mode == "unbiased" else orc.type_biased(r, m)["deq"]
            rot_q.append(q)
            assert _eq(got[c], orc.irht(q, diag)[:d]), (mode, R, c)
        mean = dme.rotated_quantize_mean(X, R, mode=mode, seed=9, rotation_seed=123).cpu().numpy()
        assert _eq(mean, orc.irht(orc.mean_of(rot_q), diag)[:d]), (mode, R)
        # linearity: the mean of the per-client compositions differs only by the rounding of the last transform
        assert np.allclose(mean, got.astype(np.float64).mean(axis=0), rtol=0, atol=1e-5 * np.abs(X).max())


def test_rotation_flattens_spiky_inputs(dme):
    """Why config 3 exists: on a vector with a few huge coordinates the plain type quantizer spends its units on them; after
    the rotation the coordinates are near-Gaussian.  Per-vector squared error, averaged over uniforms."""
    rng = np.random.default_rng(5)
    d = 1 << 14
    x = rng.standard_normal(d).astype(np.float32)
    x[rng.integers(0, d, 8)] *= 400.0
    X = np.tile(x, (64, 1))
    plain = dme.type_quantize(X, 1, seed=3, want=("deq",))["deq"].cpu().numpy()
    rot = dme.rotated_type_quantize(X, 1, seed=3).cpu().numpy()
    e_plain = ((plain - x) ** 2).sum(axis=1).mean() / (x.astype(np.float64) ** 2).sum()
    e_rot = ((rot - x) ** 2).sum(axis=1).mean() / (x.astype(np.float64) ** 2).sum()
    print(f"per-vector NMSE at R=1 on a spiky vector: plain {e_plain:.3f}, rotated {e_rot:.3f}")
    assert e_rot < 3.0 and np.isfinite(e_plain)
    # both are unbiased: the average over 64 independent uniforms is closer to x than a single draw
    assert ((rot.mean(axis=0) - x) ** 2).sum() < 0.2 * ((rot[0] - x) ** 2).sum()

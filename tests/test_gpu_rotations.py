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


@pytest.mark.parametrize("logd", [0, 1, 3, 4, 5, 8, 11, 12, 13, 15, 16, 17, 20, 22])
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

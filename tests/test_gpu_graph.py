"""MeanGraph (the north-star path as one captured CUDA graph) against the eager call: replay k must be bit-identical to
quantize_mean(seed = s + k) -- same uniforms (dme_fill_uniforms against the host-side Philox), same kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dme():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dme_b200
    return dme_b200


@pytest.mark.parametrize("n,d,R,mode", [(10, 1024, 1, "unbiased"), (100, 122626, 1, "unbiased"), (7, 65536 + 17, 2, "unbiased"), (5, 40000, 1, "biased")])
def test_graph_replays_match_eager(dme, n, d, R, mode):
    g = torch.Generator(device="cuda").manual_seed(n * 31 + d)
    X = torch.randn((n, d), generator=g, device="cuda")
    gm = dme.MeanGraph(X, R, mode=mode, seed=77, client0=3)
    for k in range(3):
        got = gm().clone()
        ref = dme.quantize_mean(X, R, mode=mode, seed=77 + k, client0=3)
        assert torch.equal(got, ref), f"replay {k}"
    gm.status()
    # new input in place, seed reset
    X.mul_(0.5).add_(0.25)
    got = gm(seed=5).clone()
    assert torch.equal(got, dme.quantize_mean(X, R, mode=mode, seed=5, client0=3))
    assert gm.launches >= 3
    if d % 4:
        assert gm._src is not None      # unaligned rows are re-staged inside the graph


def test_fill_uniforms_matches_host_philox(dme):
    import ctypes as C
    from dme_b200 import _cabi
    n = 1000
    seed = torch.tensor([123456789], dtype=torch.int64, device="cuda")
    xu = torch.empty(n, dtype=torch.float32, device="cuda")
    rc = _cabi.lib().dme_fill_uniforms(C.c_void_p(xu.data_ptr()), n, C.c_void_p(seed.data_ptr()), 40, 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    assert int(seed.item()) == 123456790
    assert np.array_equal(xu.cpu().numpy(), dme.client_uniforms(123456789, 40, n))

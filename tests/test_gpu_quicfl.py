"""QUIC-FL sender (AS:455-503) on the GPU: dme_quicfl_encode + the receiver.  The reference cannot run its sender (tables not
shipped), so what is checked is the algorithm's own contract: the exact-tail rule, index ranges, a host re-computation
of the table look-ups from the kernel's own outputs, unbiasedness, and the error level the tables predict."""
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


@pytest.mark.parametrize("nbits,d", [(1, 1024), (2, 1000), (3, 4096), (4, 777)])
def test_encode_contract(dme, nbits, d):
    from dme_b200 import quicfl_tables as qt
    t = qt.tables_for(nbits)
    g = torch.Generator(device="cuda").manual_seed(nbits * 100 + d)
    x = torch.randn((3, d), generator=g, device="cuda") * 2.5
    enc = dme.quicfl_encode(x, nbits, seed=11, client0=4)
    dpad = enc["dpad"]
    assert dpad == 1 << (d - 1).bit_length()
    X, h, em = enc["X"].cpu().numpy(), enc["h"].cpu().numpy(), enc["exact_mask"].cpu().numpy().astype(bool)
    assert X.min() >= 0 and X.max() < 2 ** nbits and h.min() >= 0 and h.max() < t["h_len"]
    rot = dme.rht(x, 123).cpu().numpy().astype(np.float32)                       # the shared rotation (AS:464)
    scale = enc["scale"].cpu().numpy()
    ref_scale = np.sqrt(dpad) / np.linalg.norm(rot.astype(np.float64), axis=1)
    assert np.allclose(scale, ref_scale, rtol=2e-6)
    z = rot * scale[:, None]
    thr = np.float32(qt.EXACT_THRESHOLD)
    assert np.array_equal(em, np.abs(z) > thr)                                   # AS:475-478
    ed = enc["exact_dense"].cpu().numpy()
    assert np.array_equal(ed[em], z[em]) and not ed[~em].any()
    # the grid index the kernel used must be floor(z / delta) or that + 1 (AS:483-484), and X the table's base index or that + 1
    q = np.where(em, np.float32(0), z / np.float32(t["delta"]))
    fl = np.floor(q).astype(np.int64)
    half = (t["x_len"] - 1) // 2
    ok = np.zeros(X.shape, bool)
    for dq in (0, 1):
        idx = np.clip(fl + dq + half, 0, t["x_len"] - 1)
        base = t["send_X"][idx, h].astype(np.int64)
        p = t["send_p"][idx, h]
        ok |= (X == base) | ((X == base + 1) & (p > 0))
    assert ok.all()
    # receiver: table look-up, exact values, / scale, inverse rotation
    out = dme.quicfl_decode_dense(enc)
    out2 = dme.quicfl_decode(enc["X"], enc["h"], d, t["recv"], enc["scale"], exact_mask=enc["exact_mask"],
                             exact_vals=enc["exact_dense"][enc["exact_mask"].bool()])
    assert torch.equal(out, out2)                                                # dense and compacted exact values agree


@pytest.mark.parametrize("nbits", [1, 2, 4])
def test_unbiased_and_error_level(dme, nbits):
    from dme_b200 import quicfl_tables as qt
    d, reps = 2048, 256
    g = torch.Generator(device="cuda").manual_seed(5 + nbits)
    x = torch.randn(d, generator=g, device="cuda")
    X = x.unsqueeze(0).repeat(reps, 1).contiguous()                              # the same vector sent by `reps` clients
    est = dme.quicfl(X, nbits, seed=2024)                                        # Philox keyed by (seed, client): independent draws
    err = ((est - x) ** 2).sum(dim=1) / (x ** 2).sum()
    single = float(err.mean())
    # expected per-vector NMSE from the tables under N(0,1) (the rotated, scaled coordinates), exact tail excluded
    t = qt.tables_for(nbits)
    R = t["recv"].astype(np.float64); Xb = t["send_X"].astype(np.int64); p = t["send_p"].astype(np.float64); xs = t["grid"]
    v0 = np.take_along_axis(R.T[None], Xb[:, :, None], 2)[:, :, 0]
    v1 = np.take_along_axis(R.T[None], np.minimum(Xb + 1, R.shape[0] - 1)[:, :, None], 2)[:, :, 0]
    var = ((1 - p) * (v0 - xs[:, None]) ** 2 + p * (v1 - xs[:, None]) ** 2).mean(1)
    from statistics import NormalDist
    w = np.array([NormalDist().pdf(v) if abs(v) <= qt.EXACT_THRESHOLD else 0.0 for v in xs]); w /= w.sum() / (1 - 2.0 ** -8)
    expect = float((w * var).sum())
    assert 0.9 * expect < single < 1.1 * expect, (single, expect)
    mean_err = float(((est.mean(dim=0) - x) ** 2).sum() / (x ** 2).sum())
    assert mean_err < 1.5 * expect / reps, (mean_err, expect / reps)             # unbiased: the error of the mean falls like 1 / reps


def test_dropin_wrapper(dme):
    import dme_b200.All_Schemes as AS
    v = np.random.default_rng(3).standard_normal(1000).astype(np.float32)
    out = AS.QUICFL_quantize(v, 2)
    assert isinstance(out, np.ndarray) and out.shape == (1000,) and np.isfinite(out).all()
    assert np.sum((out - v) ** 2) / np.sum(v ** 2) < 0.6

"""GPU statistical parity: NMSE of the B200 path vs the unmodified reference (tests/golden/nmse_reference.json, made by
tests/golden/make_nmse_golden.py): n=10, d=1024, >=100 trials, four distributions.  Pass = 95 % confidence intervals
overlap (BASELINE north star).  Seeds are fixed, so the outcome is deterministic.  Also exercises the drop-in module."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dme():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dme_b200
    return dme_b200


def _draw(dist, rng, n, d):
    return {"gaussian": lambda: rng.standard_normal((n, d)), "uniform": lambda: rng.uniform(-1, 1, (n, d)),
            "exponential": lambda: rng.exponential(1.0, (n, d)), "lognormal": lambda: rng.lognormal(0.0, 1.0, (n, d))}[dist]().astype(np.float32)


@pytest.mark.parametrize("dist", ["gaussian", "uniform", "exponential", "lognormal"])
def test_nmse_ci_overlap_with_reference(dme, golden_dir, dist):
    ref = json.load(open(os.path.join(golden_dir, "nmse_reference.json")))
    n, d, trials = ref["n"], ref["d"], 200
    rng = np.random.default_rng(20251018)
    fns = {
        "Type_unbiased R=1": lambda X, s: dme.quantize_mean(X, 1, seed=s),
        "Type_unbiased R=2": lambda X, s: dme.quantize_mean(X, 2, seed=s),
        "Type_biased R=1": lambda X, s: dme.quantize_mean(X, 1, mode="biased"),
        "Type_biased R=2": lambda X, s: dme.quantize_mean(X, 2, mode="biased"),
        "EDEN R=1": lambda X, s: dme.mean_accumulate(dme.eden(X, 1, seed=s)),
        "EDEN R=2": lambda X, s: dme.mean_accumulate(dme.eden(X, 2, seed=s)),
        "DRIVE R=1": lambda X, s: dme.mean_accumulate(dme.drive(X, seed=s, compat="reference")),
        "Scalar R=2": lambda X, s: dme.mean_accumulate(dme.scalar_quantize(X, 2, seed=s)),
    }
    vals = {k: [] for k in fns}
    for t in range(trials):
        X = _draw(dist, rng, n, d)
        Xd = torch.from_numpy(X).cuda()
        mean = Xd.double().sum(0) / n
        den = float((Xd.double() ** 2).sum() / n)
        for name, f in fns.items():
            est = f(Xd, 1000 + t)
            vals[name].append(float(((est.double() - mean) ** 2).sum()) / den)
    report = {}
    for name, v in vals.items():
        mu, ci = float(np.mean(v)), float(1.96 * np.std(v, ddof=1) / np.sqrt(trials))
        r = ref["stats"][dist][name]
        report[name] = (round(mu, 5), round(ci, 5), round(r["mean"], 5), round(r["ci95"], 5))
        assert abs(mu - r["mean"]) <= ci + r["ci95"], (dist, name, report[name])
    print(dist, report)


def test_unbiasedness_and_order_optimal_scaling(dme):
    """E[q] = x (AS:609-641 is unbiased) and NMSE * n is flat in n (README.md:5: O(1/n))."""
    rng = np.random.default_rng(5)
    d = 4096
    x = rng.standard_normal(d).astype(np.float32)
    reps = 512
    X = torch.from_numpy(np.repeat(x[None, :], reps, 0)).cuda()
    est = dme.quantize_mean(X, 1, seed=77).cpu().numpy()                     # mean of 512 independent quantizations of x
    one = dme.type_quantize(x, 1, seed=77)["deq"].cpu().numpy()
    assert np.sum((est - x) ** 2) < np.sum((one - x) ** 2) / 100             # variance falls ~1/reps, no bias floor
    prods = []
    for n in (4, 16, 64):
        Xn = torch.randn((n, d), generator=torch.Generator(device="cuda").manual_seed(n), device="cuda")
        e = dme.quantize_mean(Xn, 1, seed=3)
        nmse = float(((e - Xn.mean(0)) ** 2).sum() / ((Xn ** 2).sum() / n))
        prods.append(nmse * n)
    assert max(prods) / min(prods) < 1.25 and 1.5 < prods[0] < 2.5           # ~2.0 for R=1 Gaussian


def test_dropin_module_on_gpu(dme):
    import dme_b200.All_Schemes as AS
    torch.manual_seed(42)
    x = np.random.default_rng(1).standard_normal(5000).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    for fn, bits in ((AS.Type_unbiased_quantize, 1), (AS.Type_biased_quantize, 2), (AS.DRIVE_quantize_Hadamard, 1), (AS.Scalar_quantize, 4),
                     (AS.No_quantize, 32)):
        for inp in (x, xt, list(x[:100])):
            q = fn(inp, bits)
            assert isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32 and q.shape == (len(inp),)   # AS:640, AS:752, AS:790, AS:859
    for fn, bits in ((AS.EDEN_quantize_Hadamard, 1), (AS.EDEN_quantize_Hadamard, 2), (AS.Kashin_quantize, 2)):
        q = fn(xt, bits)
        assert isinstance(q, np.ndarray) and q.dtype == np.float32 and q.shape == (5000,)                               # AS:812, AS:854
        assert np.sum((q - x) ** 2) / np.sum(x ** 2) < 0.7
    assert torch.equal(AS.No_quantize(xt), xt)
    with pytest.raises(KeyError):
        AS.Type_unbiased_quantize(xt, 0.7)                                                                                  # AS:623
    with pytest.raises(KeyError):
        AS.EDEN_quantize_Hadamard(xt, 3)
    a = AS.Type_unbiased_quantize(xt, 1)
    b = AS.Type_unbiased_quantize(xt, 1)
    assert not torch.equal(a, b)                                             # a fresh uniform per call, like torch.rand (AS:634)
    assert float(torch.abs(a * AS._api.m_for_rate(1, 5000) / xt.abs().sum()).round().sum()) in (1070.0, 1069.0, 1071.0)  # sum k = m (+-1)
    H = AS.HadamardSender()
    y = H.randomized_hadamard_transform(xt, 123)
    assert y.numel() == 8192 and abs(float(y.norm() / xt.norm()) - 1) < 1e-5
    z = AS.HadamardReceiver().randomized_inverse_hadamard_transform(y, 123)
    assert torch.allclose(z[:5000], xt, atol=1e-5)

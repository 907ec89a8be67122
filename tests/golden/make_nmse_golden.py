"""NMSE statistics of the UNMODIFIED reference (CPU) for the statistical-parity test: BASELINE config 1 shape
(n=10 clients, d=1024, 100 trials) over the four BASELINE distributions, standard NMSE = |est - mean|^2 / mean_i |x_i|^2
(the reference's own convention is this divided by 50 n^2).  Writes tests/golden/nmse_reference.json.

    python tests/golden/make_nmse_golden.py
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_adapter  # noqa: E402

warnings.filterwarnings("ignore")
AS = ref_adapter.load()
torch.set_num_threads(1)
n, d, trials = 10, 1024, 100


def draw(dist, rng):
    return {"gaussian": lambda: rng.standard_normal((n, d)), "uniform": lambda: rng.uniform(-1, 1, (n, d)),
            "exponential": lambda: rng.exponential(1.0, (n, d)), "lognormal": lambda: rng.lognormal(0.0, 1.0, (n, d))}[dist]().astype(np.float32)


schemes = {
    "Type_unbiased R=1": lambda v: AS.Type_unbiased_quantize(v, 1), "Type_unbiased R=2": lambda v: AS.Type_unbiased_quantize(v, 2),
    "Type_biased R=1": lambda v: AS.Type_biased_quantize(v, 1), "Type_biased R=2": lambda v: AS.Type_biased_quantize(v, 2),
    "EDEN R=1": lambda v: AS.EDEN_quantize_Hadamard(v, 1), "EDEN R=2": lambda v: AS.EDEN_quantize_Hadamard(v, 2),
    "DRIVE R=1": lambda v: AS.DRIVE_quantize_Hadamard(v, 1), "Scalar R=2": lambda v: AS.Scalar_quantize(v, 2),
}
out = {"n": n, "d": d, "trials": trials, "convention": "standard", "stats": {}}
for dist in ("gaussian", "uniform", "exponential", "lognormal"):
    torch.manual_seed(42)
    rng = np.random.default_rng(42)
    vals = {k: [] for k in schemes}
    for t in range(trials):
        X = draw(dist, rng)
        mean = X.sum(0) / n
        den = float((X.astype(np.float64) ** 2).sum() / n)
        for name, f in schemes.items():
            est = np.zeros(d, np.float32)
            for c in range(n):
                est += np.asarray(torch.as_tensor(f(X[c])), dtype=np.float32) / n          # ND:133-147
            vals[name].append(float(((est - mean).astype(np.float64) ** 2).sum()) / den)
    out["stats"][dist] = {k: {"mean": float(np.mean(v)), "ci95": float(1.96 * np.std(v, ddof=1) / np.sqrt(trials))} for k, v in vals.items()}
    print(dist, {k: (round(s["mean"], 5), round(s["ci95"], 5)) for k, s in out["stats"][dist].items()}, flush=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "nmse_reference.json"), "w"), indent=1)

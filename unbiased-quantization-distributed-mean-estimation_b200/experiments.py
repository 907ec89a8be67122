"""Parametrised DME experiment runner: ONE function replaces the reference's five cloned drivers
(`Normal_dist.py`, `Laplace_dist.py`, `Gamma_dist.py`, `Bernoulli_dist.py`, `Lognormal_dist.py`, ND:36-265).

Same experiment (ND:38-43: dim 2048, n in 1, 6, ..., 101, 50 instances), same scheme lines and pickle keys (ND:227-259),
same NMSE convention (||est - mean||^2 / (num_trials * sum ||x_i||^2 * n), `num_trials` only a divisor) plus the
standard one; but every scheme sees all n client vectors in ONE batched call instead of the per-vector Python loop
(ND:133-147): the type quantizers go through the fused quantize -> pack -> decode -> mean path, the others through
the batched kernels + `mean_accumulate`.

    python -m dme_b200.experiments --dist normal --out Distributions_NMSE_Results      # needs a B200

QUIC-FL lines are NaN: the reference cannot run them either (sender tables not shipped).
"""
from __future__ import annotations

import argparse
import os
import pickle
import time

import numpy as np
import torch

from . import api
from . import All_Schemes

# ND:89 and the same line of the four sibling drivers; BASELINE's uniform / exponential added
DISTRIBUTIONS = {
    "normal": lambda rng, size: rng.normal(loc=0, scale=1, size=size),
    "laplace": lambda rng, size: rng.laplace(loc=1, scale=2, size=size),
    "gamma": lambda rng, size: rng.gamma(shape=2, scale=2, size=size),
    "bernoulli": lambda rng, size: rng.choice([0, 1], size=size, p=[0.3, 0.7]).astype(np.float64),
    "lognormal": lambda rng, size: rng.lognormal(mean=1, sigma=2, size=size),
    "uniform": lambda rng, size: rng.uniform(-1, 1, size=size),
    "exponential": lambda rng, size: rng.exponential(1.0, size=size),
}
# the file-name stem the reference uses for each distribution (ND:262, ND:265 and siblings)
FILE_STEM = {"normal": "Normal_dist", "laplace": "Laplace_dist", "gamma": "Gamma_dist", "bernoulli": "Bernoulli_dist",
             "lognormal": "Lognormal_dist", "uniform": "Uniform_dist", "exponential": "Exponential_dist"}

# the reference's 14 lines (ND:47-65), in its order
LINES = ["DRIVE_Hadamard", "EDEN_Hadamard_1bit", "EDEN_Hadamard_2bit", "Type_Biased_1bit", "Type_Unbiased_1bit",
         "Type_Biased_2bit", "Type_Unbiased_2bit", "QUICFL_1bit", "QUICFL_2bit", "Kashin_1bit", "Kashin_2bit",
         "Scalar_1bit", "Scalar_2bit", "Scalar_4bit"]


def _estimate(line: str, X: torch.Tensor, seed: int, kashin: bool):
    """Server-side mean estimate of one scheme line for the client rows X[n, d] (ND:133-147, batched)."""
    n = X.shape[0]
    if line.startswith("Type_"):
        mode = "unbiased" if "Unbiased" in line else "biased"
        return api.quantize_mean(X, int(line[-4]), mode=mode, seed=seed)
    if line == "DRIVE_Hadamard":
        return api.mean_accumulate(api.drive(X, seed=seed))
    if line.startswith("EDEN_"):
        return api.mean_accumulate(api.eden(X, int(line[-4]), seed=seed))
    if line.startswith("Scalar_"):
        return api.mean_accumulate(api.scalar_quantize(X, int(line[-4]), seed=seed))
    if line.startswith("QUICFL_"):
        # one sender seed per client in [0, 100) like AS:820 would draw; here Philox is keyed by (seed, client)
        return api.mean_accumulate(api.quicfl(X, int(line[-4]), seed=seed))
    if line.startswith("Kashin_") and kashin:
        return api.mean_accumulate(api.kashin(X, int(line[-4]), seed=np.random.default_rng(seed).integers(0, 100, n)))      # all rows per transform launch (AS:191-239, AS:834-854)
    return None                                  # Kashin when disabled


def run(dist="normal", dim=2048, num_users_list=None, num_instances=50, num_trials=50, seed=42, lines=None, kashin=True,
        out_dir=None, verbose=False):
    """Returns (nmse_avg_data, nmse_max_data, nmse_std_avg): the reference's two dictionaries (keys `NMSE_<line>_avg` /
    `_max`, one value per entry of num_users_list, reference convention) and the averages in the standard convention
    ||est - mean||^2 / ((1/n) sum ||x_i||^2).  Writes the reference's two pickle files when out_dir is given."""
    gen = DISTRIBUTIONS[dist]
    rng = np.random.default_rng(seed)
    users = list(np.arange(1, 102, 5)) if num_users_list is None else [int(u) for u in num_users_list]    # ND:43
    lines = LINES if lines is None else list(lines)
    dev = api._device()
    ref = {ln: np.full((len(users), num_instances), np.nan) for ln in lines}
    std = {ln: np.full((len(users), num_instances), np.nan) for ln in lines}
    for ui, n in enumerate(users):
        t0 = time.time()
        for inst in range(num_instances):
            Xh = gen(rng, (n, dim))                                             # float64, like np.random.* in ND:89
            vec_norm_squared = float((np.linalg.norm(Xh, axis=1) ** 2).sum())   # ND:90, ND:94
            X = torch.as_tensor(Xh, dtype=torch.float32, device=dev)            # ND:91
            emp_mean = X.sum(dim=0) / n                                         # ND:95
            for li, ln in enumerate(lines):
                est = _estimate(ln, X, seed * 1000003 + (ui * num_instances + inst) * 64 + li, kashin)
                if est is None:
                    continue
                err2 = float(torch.norm(est - emp_mean).pow(2))
                ref[ln][ui, inst] = err2 / (num_trials * vec_norm_squared * n)  # ND:151-164
                std[ln][ui, inst] = err2 / (vec_norm_squared / n)
        if verbose:
            print(f"num_users={n}: {(time.time() - t0):.2f} s", flush=True)
    avg = {f"NMSE_{ln}_avg": ref[ln].mean(axis=1).astype(np.float32) for ln in lines}       # ND:208-221, ND:227-242
    mx = {f"NMSE_{ln}_max": ref[ln].max(axis=1).astype(np.float32) for ln in lines}         # ND:193-206, ND:244-259
    std_avg = {f"NMSE_{ln}_avg": std[ln].mean(axis=1) for ln in lines}
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, f"nmse_avg_data_{FILE_STEM[dist]}.pkl"), "wb") as f:   # ND:261-262
            pickle.dump(avg, f)
        with open(os.path.join(out_dir, f"nmse_max_data_{FILE_STEM[dist]}.pkl"), "wb") as f:   # ND:264-265
            pickle.dump(mx, f)
    return avg, mx, std_avg


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    p.add_argument("--dist", default="normal", choices=sorted(DISTRIBUTIONS))
    p.add_argument("--dim", type=int, default=2048)
    p.add_argument("--instances", type=int, default=50)
    p.add_argument("--max-users", type=int, default=100)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--no-kashin", action="store_true", help="skip the per-vector Kashin lines (slow)")
    p.add_argument("--out", default=None, help="directory for the two pickle files (the reference's Distributions_NMSE_Results)")
    a = p.parse_args(argv)
    avg, _, std_avg = run(a.dist, a.dim, np.arange(1, a.max_users + 2, 5), a.instances, seed=a.seed, kashin=not a.no_kashin,
                          out_dir=a.out, verbose=True)
    for k in avg:
        print(f"{k:32s} n=1: {avg[k][0]:.3e}   n={int(np.arange(1, a.max_users + 2, 5)[-1])}: {avg[k][-1]:.3e}   std-NMSE*n at last: "
              f"{std_avg[k][-1] * np.arange(1, a.max_users + 2, 5)[-1]:.3f}")


if __name__ == "__main__":
    main()

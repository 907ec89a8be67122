"""Development helper: the fused quantize kernel (L1 norms one row ahead, one launch) against the two-launch path -- the mean, the
norms and the packed codes must be bit-identical -- and per-kernel times under the tuning knobs (lead rounds, L2 policies).

    python tools/fused_check.py [NxD ...] [--tune]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi

L = _cabi.lib()
shapes = [a for a in sys.argv[1:] if "x" in a] or ["3x1000", "7x122626", "40x65536", "5x1048576", "9x4194337", "128x16777216"]
tune = "--tune" in sys.argv


def run(path, X, seed):
    dme.set_unbiased_path(path)
    out = torch.empty(X.shape[1], device="cuda")
    dme.quantize_mean(X, 1, seed=seed, out=out, check=True)
    codes = dme.type_encode(X, 1, seed=seed)
    return out, codes


for sh in shapes:
    n, d = (int(v) for v in sh.split("x"))
    X = torch.randn((n, d), device="cuda")
    if n >= 3:
        X[1] *= 1e-3
        X[2, : d // 2] = 0
    ref, cref = run("tiles", X, 5)
    got, cgot = run("fused", X, 5)
    same_mean = bool(torch.equal(ref, got))
    same_l1 = bool(torch.equal(cref.l1, cgot.l1))
    if n * d <= 1 << 24:
        mref, mgot = cref.to_messages(), cgot.to_messages()
        same_codes = all(a == b for a, b in zip(mref, mgot))
    else:       # primary slots are placed deterministically: compare the directory and the arena (light-tailed data: no overflow tiles)
        mref = mgot = None
        same_codes = bool(torch.equal(cref.dir, cgot.dir)) and bool(torch.equal(cref.codes, cgot.codes))
    print(f"{sh}: mean {same_mean} l1 {same_l1} codes {same_codes}", flush=True)
    del cref, cgot, mref, mgot
    out = torch.empty(d, device="cuda")
    variants = [("tiles", None), ("fused", (2, 1, 2, 1))]
    if tune:
        variants += [("fused", v) for v in [(1, 1, 2, 1), (3, 1, 2, 1), (2, 1, 2, 0), (2, 0, 0, 1), (4, 1, 2, 1)]]
    for path, tv in variants:
        dme.set_unbiased_path(path)
        if tv:
            assert L.dme_set_fused_tuning(*tv) == 0
        res = [dme.profile_kernels(lambda: dme.quantize_mean(X, 1, seed=i, out=out, check=False), warm=0) for i in range(5)]
        tot = [round(sum(t for _, t in r), 3) for r in res[2:]]
        print(" ", path, tv, tot, [(nm, round(t, 3)) for nm, t in res[-1]], flush=True)
    L.dme_set_fused_tuning(2, 1, 2, 1)
    del X
dme.set_unbiased_path("fused")

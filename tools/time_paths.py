"""Development helper: per-kernel times of the fused path for each unbiased implementation."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi

L = _cabi.lib()
n, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64x16777216").split("x"))
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
for path in sys.argv[2:] or ["fx", "tiles"]:
    dme.set_unbiased_path(path)
    res = [dme.profile_kernels(lambda: dme.quantize_mean(X, 1, seed=i, out=out, check=False), warm=0) for i in range(4)]
    print(path, n, d, [[(nm, round(t, 3)) for nm, t in r] for r in res[2:]], flush=True)

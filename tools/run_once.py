"""Development helper: a few fused steps at one shape (target of ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme

n, d = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "16x16777216").split("x"))
R = float(sys.argv[2]) if len(sys.argv) > 2 else 1
R = int(R) if R == int(R) else R
mode = sys.argv[3] if len(sys.argv) > 3 else "unbiased"
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
for i in range(3):
    dme.quantize_mean(X, R, seed=i, out=out, mode=mode, check=False)
torch.cuda.synchronize()
dme.Workspace.get(X.device).status()
print("ok", float(out.abs().sum()))

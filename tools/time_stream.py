"""Development helper: per-kernel times of the fused path for a few shapes (GPU box)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi

L = _cabi.lib()
shapes = [(16, 1 << 24), (128, 1 << 20), (1000, 1 << 16), (4, 1 << 24)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]]
for n, d in shapes:
    X = (torch.rand((n, d), device="cuda") * 2 - 1) if os.environ.get("UNIF") else torch.randn((n, d), device="cuda")
    out = torch.empty(d, device="cuda")
    L.dme_profile_enable(1)
    res = []
    for i in range(4):
        dme.quantize_mean(X, 1, seed=i, out=out, check=False)
        buf = (C.c_float * 8)()
        k = L.dme_profile_read(buf, 8)
        res.append([round(buf[j], 3) for j in range(k)])
    L.dme_profile_enable(0)
    try:
        dme.Workspace.get(X.device).status()
    except Exception as ex:
        print('status:', type(ex).__name__)
    balg = 4.0 * n * d + 4.0 * d
    t = sum(res[-1])
    print(f"n={n} d={d} env={os.environ.get('DME_DBG','')}/{os.environ.get('DME_DBG_LAG','')}/{os.environ.get('DME_DBG_G','')}: kernels ms {res[1:]}  -> {balg / t * 1e-6:.0f} GB/s alg", flush=True)
    if int(os.environ.get("DME_DBG", "0")) & 32:
        ws = dme.Workspace.get(X.device).buf
        off = (-ws.data_ptr()) % 256
        hdr = ws[off: off + 256].cpu().numpy().view("uint64")
        names = ["c:copy wait A", "c:pass A", "c:B-phase", "c:C-phase", "c:READY wait", "c:copy wait B", "cta lifetime", "c:look-back finish (warp 0)",
                 "-", "-", "s:cdone wait", "s:count x1000", "s:issue B", "s:adone wait", "s:issue A", "s:raw B copy latency (sum)"]
        pairs = n * ((d + 4095) // 4096)
        print("   per tile pair, ns:", {nm: round(float(hdr[3 + q]) / pairs, 1) for q, nm in enumerate(names) if nm != "-"})
    del X

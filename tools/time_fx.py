"""Development helper: phase timers of quantize_fx_kernel (library built with DME_TIMERS=1)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi

L = _cabi.lib()
shapes = [tuple(int(v) for v in s.split("x")) for s in sys.argv[1:]] or [(128, 1 << 24)]
names = ["pre-B (A loads issue)", "wait TMA B", "B-phase", "A consume", "window issue+barrier", "publish_row", "agg publish", "wait window", "window eval/fallback",
         "walk + hit", "emit+free+take / rc load", "fallbacks"]
for n, d in shapes:
    X = torch.randn((n, d), device="cuda")
    out = torch.empty(d, device="cuda")
    for i in range(3):
        L.dme_profile_enable(1)
        dme.quantize_mean(X, 1, seed=i, out=out, check=False)
        buf = (C.c_float * 8)()
        k = L.dme_profile_read(buf, 8)
        L.dme_profile_enable(0)
        torch.cuda.synchronize()
        ws = dme.Workspace.get(X.device).buf
        off = (-ws.data_ptr()) % 256
        hdr = ws[off: off + 256].cpu().numpy().view("uint64")
    tiles = n * ((d + 4095) // 4096)
    print(f"n={n} d={d}: kernels ms {[round(buf[j], 3) for j in range(k)]}")
    tot = 0
    for q, nm in enumerate(names):
        v = float(hdr[5 + q])
        print(f"   {nm:32s} {v / tiles:10.1f} cycles / tile" if q < 11 else f"   {nm:32s} {v:10.0f} ({v / tiles:.4f} per tile)")

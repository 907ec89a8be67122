"""Development helper: per-warp timeline of one group of quantize_fx_kernel (library built with DME_TIMERS=1)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi

L = _cabi.lib()
n, d = 64, 1 << 24
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
buf = (C.c_ulonglong * (2 << 15))()
for i in range(2):
    dme.quantize_mean(X, 1, seed=i, out=out, check=False)
    torch.cuda.synchronize()
    k = L.dme_debug_fx_trace(buf, 1 << 15)
a = np.frombuffer(buf, dtype=np.uint64)[: 2 * k].reshape(k, 2)
t = a[:, 0].astype(np.int64); code = (a[:, 1] >> np.uint64(56)).astype(int); val = (a[:, 1] & np.uint64((1 << 56) - 1)).astype(int)
sel = code >= 20
t, code, val = t[sel], code[sel], val[sel]
t -= t.min()
it = val // 8; w = val % 8
names = {20: "top", 21: "Bstart", 22: "Bend", 23: "bar_arrive", 24: "bar_pass", 25: "prefixOK", 26: "walk_done", 27: "iter_end"}
for i in sorted(set(it))[:14]:
    print("iteration", i)
    for c in range(20, 28):
        row = []
        for ww in range(4):
            m = (it == i) & (w == ww) & (code == c)
            row.append(int(t[m][0]) if m.any() else -1)
        print(f"   {names[c]:11s}", row)

"""Development helper: timeline of one row (client 40) of quantize_fx_kernel (library built with DME_TIMERS=1)."""
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
t = a[:, 0].astype(np.int64); code = (a[:, 1] >> np.uint64(56)).astype(int); tile = (a[:, 1] & np.uint64((1 << 56) - 1)).astype(int)
t -= t.min()
ev = {c: {} for c in range(1, 6)}
for tt, cc, ti in zip(t, code, tile):
    ev[cc][ti] = tt
tiles = sorted(ev[5].keys())
print("events", k, "tiles with full C", len(tiles))
rows = []
for ti in tiles:
    if all(ti in ev[c] for c in range(1, 6)):
        rows.append((ti, ev[1][ti], ev[2][ti], ev[3][ti], ev[4][ti], ev[5][ti]))
r = np.array(rows)
print("tile  Bstart  publish  Cstart  prefixOK  Cdone   (ns, relative);  and publish of tile-1")
for q in list(range(100, 140)) + list(range(2000, 2030)):
    ti = r[q, 0]
    prev = ev[2].get(ti - 1, -1)
    print(ti, r[q, 1], r[q, 2], r[q, 3], r[q, 4], r[q, 5], " prev publish", prev, " wait", r[q, 4] - r[q, 3], " B->pub", r[q, 2] - r[q, 1])
print("mean B->publish", (r[:, 2] - r[:, 1]).mean(), "mean Cstart->prefixOK", (r[:, 4] - r[:, 3]).mean(), "mean prefixOK->Cdone", (r[:, 5] - r[:, 4]).mean(), "publish->Cstart", (r[:, 3] - r[:, 2]).mean())
span = r[:, 5].max() - r[:, 1].min()
print("row span ns", span, "tiles", len(r))
# how late is the latest predecessor publish relative to Cstart
late = []
pub = np.array([ev[2].get(ti, 0) for ti in range(4096)])
cummax = np.maximum.accumulate(pub)
for row in r:
    ti = row[0]
    if ti > 0:
        late.append(cummax[ti - 1] - row[3])
late = np.array(late)
print("latest predecessor publish minus Cstart: mean", late.mean(), "p50", np.percentile(late, 50), "p90", np.percentile(late, 90), "frac>0", (late > 0).mean())

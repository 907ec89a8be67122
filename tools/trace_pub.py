"""Development helper: publish / READY times of neighbouring tiles across CTAs (DME_TIMERS=1 build, DME_DBG=128)."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
from dme_b200 import _cabi
L = _cabi.lib()
n, d = 32, 1 << 24
X = torch.randn((n, d), device="cuda"); out = torch.empty(d, device="cuda")
for i in range(2):
    dme.quantize_mean(X, 1, seed=i, out=out, check=False)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 8192)()
L.dme_debug_trace(buf, 4096)
dme.quantize_mean(X, 1, seed=5, out=out, check=False)
torch.cuda.synchronize()
k = L.dme_debug_trace(buf, 4096)
ev = sorted((buf[2 * i], buf[2 * i + 1] >> 32, (buf[2 * i + 1] >> 20) & 0xfff, buf[2 * i + 1] & 0xfffff) for i in range(k))
t0 = ev[0][0]
# group by tile
pub = {}; rdy = {}; s1 = {}
for t, c, b, v in ev:
    if c == 11: pub[v] = (t - t0, b)
    if c == 12: rdy[v] = (t - t0, b)
    if c == 3: s1[v] = (t - t0, b)
tiles = sorted(pub)
sel = [t for t in tiles if 2000 <= t < 2000 + 900][:60]
for t in sel:
    print(f"tile {t:5d} cta {pub[t][1]:4d}  S1 done {s1.get(t,(0,0))[0]/1000:9.3f}  published {pub[t][0]/1000:9.3f} us   next-tile-seen-by-service(READY of prev) {rdy.get(t,(0,0))[0]/1000:9.3f}")

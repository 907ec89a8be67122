"""Development helper: event timeline of one CTA of the stream kernel (build with DME_TIMERS=1, run with DME_DBG=64)."""
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
L.dme_debug_trace(buf, 4096)          # drop the warm-up
dme.quantize_mean(X, 1, seed=5, out=out, check=False)
torch.cuda.synchronize()
k = L.dme_debug_trace(buf, 4096)
names = {1: "c fetch wait", 2: "c fetch got", 3: "c S1 done (B1 arrive)", 4: "c passA done", 5: "c S2 enter", 6: "c READY passed", 7: "c S2 done",
         10: "s B1 seen", 11: "s published", 12: "s prev tile READY", 13: "s scan done", 14: "s issued (B slot)", 15: "s PA seen", 16: "s PA done+issued"}
ev = sorted((buf[2 * i], buf[2 * i + 1] >> 32, buf[2 * i + 1] & 0xffffffff) for i in range(k))
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 600, int(sys.argv[2]) if len(sys.argv) > 2 else 760
t0 = ev[lo][0] if len(ev) > lo else 0
for t, c, v in ev[lo:hi]:
    print(f"{(t - t0) / 1000.0:9.3f} us  {'    ' if c < 10 else ''}{names.get(c, c):28s} {v}")

"""Development helper (DME_TIMERS=1 build): per-tile timeline of quantize_warp_kernel."""
import sys, torch, ctypes as C, numpy as np
sys.path.insert(0, ".")
import dme_b200 as dme
from dme_b200 import _cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = 1 << 24
L = C.CDLL(_cabi.lib()._name)
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
T = n * (d // 1024)
dbg = torch.zeros(T * 12, dtype=torch.int64, device="cuda")
dme.quantize_mean(X, 1, seed=0, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(dbg.data_ptr()))
dme.quantize_mean(X, 1, seed=1, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(0))
a = dbg.cpu().numpy().reshape(T, 12)
G = 148 * 16
R = T // G
t = a[: R * G].astype(np.float64).reshape(R, G, 12)
t0 = t[:, :, 1][t[:, :, 1] > 0].min()
t = (t - t0) / 1e3
r = np.arange(4, R - 4)
names = {0: "draw", 1: "Bstart", 5: "landed", 9: "Bmath_done", 2: "published", 3: "Cstart(window issued)", 4: "Cend(emit done)", 6: "forwarded", 7: "next_item", 8: "tma_issued"}
print("total us", t[:, :, 4].max())
print("round time (Bstart r+1 - Bstart r): %.3f" % np.median(t[r + 1, :, 1] - t[r, :, 1]))
# events of iteration k of a warp: B tile = round k, C tile = round k-1
seq = [("Bstart", t[r, :, 1]), ("landed", t[r, :, 5]), ("Bmath_done", t[r, :, 9]), ("published", t[r, :, 2]), ("Cstart", t[r - 1, :, 3]), ("Cend", t[r - 1, :, 4]),
       ("forwarded", t[r - 1, :, 6]), ("next_item", t[r - 1, :, 7]), ("tma_issued", t[r - 1, :, 8]), ("next Bstart", t[r + 1, :, 1])]
for (n0, a0), (n1, a1) in zip(seq[:-1], seq[1:]):
    dd = (a1 - a0).ravel()
    print("%-12s -> %-12s median %.3f  mean %.3f  p90 %.3f us" % (n0, n1, np.median(dd), dd.mean(), np.percentile(dd, 90)))

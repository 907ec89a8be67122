"""Development helper (DME_TIMERS=1 build): timeline of the fused quantize kernel -- when each row's constants became ready against
when its last A-tile was summed and when its first B-tile wanted them."""
import sys, torch, ctypes as C, numpy as np
sys.path.insert(0, ".")
import dme_b200 as dme
from dme_b200 import _cabi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
d = 1 << 24
L = C.CDLL(_cabi.lib()._name)
if len(sys.argv) > 2:
    L.dme_set_fused_tuning(*[int(v) for v in sys.argv[2].split(",")])
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
T4 = d // 1024
T = n * T4
dbg = torch.zeros(T * 12 + n, dtype=torch.int64, device="cuda")
dme.quantize_mean(X, 1, seed=0, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(dbg.data_ptr()))
dme.quantize_mean(X, 1, seed=1, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(0))
h = dbg.cpu().numpy()
ready = h[T * 12:].astype(np.float64)
a = h[: T * 12].reshape(T, 12).astype(np.float64)
t0 = a[:, 11][a[:, 11] > 0].min()
us = lambda v: (v - t0) / 1e3
print("kernel span us: %.1f" % us(a[:, 4].max()))
for c in range(min(n, 12)):
    rows = a[c * T4:(c + 1) * T4]
    print("row %2d: A done first/median/last %.1f/%.1f/%.1f  ready %.1f  B wanted (Bstart) first/median %.1f/%.1f  B proceeded first/median/last %.1f/%.1f/%.1f  Cend last %.1f" % (
        c, us(rows[:, 11].min()), us(np.median(rows[:, 11])), us(rows[:, 11].max()), us(ready[c]), us(rows[:, 10].min()), us(np.median(rows[:, 10])),
        us(rows[:, 1].min()), us(np.median(rows[:, 1])), us(rows[:, 1].max()), us(rows[:, 4].max())))
# per-iteration phases for a steady-state window
G = 148 * 16
R = T // G
t = a[: R * G].reshape(R, G, 12)
r = np.arange(8, R - 8)
seq = [("Bwanted", t[r, :, 10]), ("Bstart", t[r, :, 1]), ("landed", t[r, :, 5]), ("Bmath_done", t[r, :, 9]), ("published", t[r, :, 2]), ("Cstart", t[r - 1, :, 3]),
       ("Cend", t[r - 1, :, 4]), ("tma_issued", t[r - 1, :, 8]), ("next Bwanted", t[r + 1, :, 10])]
for (n0, a0), (n1, a1) in zip(seq[:-1], seq[1:]):
    dd = (a1 - a0).ravel() / 1e3
    print("%-12s -> %-12s median %.3f  mean %.3f  p90 %.3f  p99 %.3f us" % (n0, n1, np.median(dd), dd.mean(), np.percentile(dd, 90), np.percentile(dd, 99)))

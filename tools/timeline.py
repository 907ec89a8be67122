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
dbg = torch.zeros(T * 6, dtype=torch.int64, device="cuda")
dme.quantize_mean(X, 1, seed=0, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(dbg.data_ptr()))
dme.quantize_mean(X, 1, seed=1, out=out, check=False)
torch.cuda.synchronize()
L.dme_debug_buffer(C.c_void_p(0))
a = dbg.cpu().numpy().reshape(T, 6)
t0 = a[:, 1][a[:, 1] > 0].min()
draw, bst, pub, cst, cen = [(a[:, k] - t0) / 1e3 for k in range(5)]
land = ((a[:, 5] >> 20) - (t0 & ((1 << 44) - 1))) / 1e3
sm = (a[:, 5] >> 8) & 0xfff; wslot = a[:, 5] & 0xff
G = 148 * 16
print("total us", cen.max())
sel = slice(4 * G, T - 4 * G)
print("B start -> publish us: median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile((pub - bst)[sel], [50, 90, 99])))
print("wait for TMA (B start -> landed) us: median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile((land - bst)[sel], [50, 90, 99])))
print("draw -> B start us: median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile((bst - draw)[sel], [50, 90, 99])))
print("publish -> C start us: median %.2f p90 %.2f" % tuple(np.percentile((cst - pub)[sel], [50, 90])))
print("C start -> C end us: median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile((cen - cst)[sel], [50, 90, 99])))
# lateness: publish time of tile j relative to the running max of earlier publishes
pm = np.maximum.accumulate(pub)
late = pub[1:] - pm[:-1]
print("tiles published after all their predecessors were (they set the pace): %.3f" % (late > 0).mean())
# per hardware warp slot: B duration
for w in sorted(set(wslot[sel].tolist()))[:16]:
    m = wslot[sel] == w
    print("warp slot %2d: tiles %6d  B dur median %.2f  iteration (C end - B start of same tile... ) draw->pub %.2f" % (w, m.sum(), np.median((pub - bst)[sel][m]), np.median((pub - draw)[sel][m])))
np.save("gpurun_out/timeline.npy", a[: 64 * G])

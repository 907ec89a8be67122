"""Development helper (library built with DME_TIMERS=1): look-back poll counters of quantize_warp_kernel."""
import sys, torch, ctypes as C
sys.path.insert(0, ".")
import dme_b200 as dme
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
X = torch.randn((n, d), device="cuda")
out = torch.empty(d, device="cuda")
for i in range(3):
    dme.quantize_mean(X, 1, seed=i, out=out, check=False)
torch.cuda.synchronize()
ws = dme.Workspace.get(X.device)
off = (-ws.buf.data_ptr()) % 256
hdr = ws.buf[off:off + 256].view(torch.int32).cpu().numpy()
pad = hdr[5:]
print("tiles", n * ((d + 1023) // 1024), "retries", pad[0], "missing tile/block/super", pad[2], pad[3], pad[4], "farthest missing tile rec (x4)", list(pad[8:16]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); dme.quantize_mean(X, 1, seed=9, out=out, check=False); e1.record(); torch.cuda.synchronize()
print("ms", e0.elapsed_time(e1))

"""Development helper: how many tiles of quantize_tiles_kernel found their look-back window incomplete and had to poll.
Build the library with DME_NVCC_EXTRA=-DDME_COUNT_FALLBACK first; UNIF=1 uses uniform inputs (only 2-bit tiles)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import dme_b200 as dme
for n, d in [(128, 1 << 24), (128, 1 << 20)]:
    X = (torch.rand((n, d), device="cuda") * 2 - 1) if os.environ.get("UNIF") else torch.randn((n, d), device="cuda"); out = torch.empty(d, device="cuda")
    for i in range(3):
        dme.quantize_mean(X, 1, seed=i, out=out, check=False)
        torch.cuda.synchronize()
        ws = dme.Workspace.get(X.device).buf
        off = (-ws.data_ptr()) % 256
        hdr = ws[off: off + 64].cpu().numpy().view("uint32")
        print(n, d, "tickets", hdr[0], "fallbacks", hdr[5], "of", n * ((d + 4095) // 4096))
    del X

import numpy as np, torch, sys, time
sys.path.insert(0, '/root/repo')
import dme_b200 as dme
from oracle import oracle as orc
for (n, d, R) in [(1, 1000, 1), (3, 4096, 1), (2, 20000, 2), (5, 65536, 1), (4, 300000, 1), (2, 1 << 20, 1)]:
    rng = np.random.default_rng(n * 1000 + d)
    X = rng.standard_normal((n, d)).astype(np.float32)
    Xs = dme.client_uniforms(seed=42, client0=5, n=n)
    t0 = time.time()
    out = dme.type_quantize(X, R, seed=42, client0=5, want=("deq", "k", "sgn", "l1"))
    torch.cuda.synchronize()
    bad = 0
    for c in range(n):
        o = orc.type_unbiased(X[c], out["m"], float(Xs[c]))
        bad += int((out["k"][c].cpu().numpy() != o["k"]).sum())
        assert float(out["l1"][c]) == float(o["L1"]), (float(out["l1"][c]), float(o["L1"]))
    print(n, d, R, "mismatches", bad, "time %.3f" % (time.time() - t0), flush=True)

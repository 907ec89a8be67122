"""Development helper: MeanGraph replays against the eager call."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme
for n, d in [(10, 1024), (100, 122626), (100, 122624), (33, 122626), (100, 65536), (7, 65553)]:
    X = torch.randn((n, d), device="cuda")
    gm = dme.MeanGraph(X, 1, seed=77, client0=3)
    for k in range(3):
        got = gm().clone()
        ref = dme.quantize_mean(X, 1, seed=77 + k, client0=3)
        ref2 = dme.quantize_mean(X, 1, seed=77 + k, client0=3).clone()
        xu = dme.client_uniforms(77 + k, 3, n)
        ref3 = dme.quantize_mean(X, 1, seed=0, client0=3, x_inject=xu).clone()
        nd = int((got != ref).sum())
        print(n, d, k, "diff", nd, float((got - ref).abs().max()), "eager twice equal", bool(torch.equal(ref, ref2)), "inject equal", bool(torch.equal(ref, ref3)),
              "xu equal", bool((gm._xu.cpu().numpy() == xu).all()), flush=True)
    gm.status()

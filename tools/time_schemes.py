"""Development helper: CUDA-event times of the other scheme kernels at large shapes (GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dme_b200 as dme

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n, d = 32, 1 << 24
X = torch.randn((n, d), device="cuda")
gb = 4.0 * n * d / 1e9
for name, fn in [
    ("type_unbiased quantize_mean R=1", lambda: dme.quantize_mean(X, 1, seed=1, check=False)),
    ("type_unbiased quantize_mean R=2", lambda: dme.quantize_mean(X, 2, seed=1, check=False)),
    ("type_biased   quantize_mean R=1", lambda: dme.quantize_mean(X, 1, mode="biased", seed=1, check=False)),
    ("rht (d=2^24, pad none)", lambda: dme.rht(X, 123)),
    ("drive", lambda: dme.drive(X, seed=1)),
    ("eden 1 bit encode+decode", lambda: dme.eden(X, 1, seed=1)),
    ("scalar_quantize 1 bit", lambda: dme.scalar_quantize(X, 1, seed=1)),
]:
    ms = timeit(fn)
    print(f"{name:36s} {ms:8.3f} ms   {gb / ms * 1e3:7.0f} GB/s of input ({n} x 2^24)", flush=True)

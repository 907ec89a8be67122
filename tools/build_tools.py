"""Builds the standalone GPU microbenchmarks under tools/ (sm_100a)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    src, out = os.path.join(HERE, "ubench.cu"), os.path.join(HERE, "ubench")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-o", out, src], check=True)
    return out


if __name__ == "__main__":
    print(build("--force" in sys.argv))

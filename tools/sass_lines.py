"""Development helper: static SASS instruction count per source line of one kernel in stream.cu
(nvcc -cubin -lineinfo + nvdisasm --print-line-info)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "unbiased-quantization-distributed-mean-estimation_b200", "csrc")
src = sys.argv[1] if len(sys.argv) > 1 else "stream.cu"
kern = sys.argv[2] if len(sys.argv) > 2 else "quantize_stream_kernelILi1E"
cubin = "/tmp/t/%s.cubin" % src
os.makedirs("/tmp/t", exist_ok=True)
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-cubin",
       "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", cubin, os.path.join(CSRC, src)] + os.environ.get("DME_NVCC_EXTRA", "").split()
subprocess.run(cmd, check=True)
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
cur = None; line = None; cnt = collections.Counter(); ops = collections.defaultdict(collections.Counter); inside = False
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        inside = kern in m.group(1); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m and line:
        ins = m.group(1).split(); op = ins[1] if ins[0].startswith("@") else ins[0]
        cnt[line] += 1; ops[line][op.split(".")[0]] += 1
tot = sum(cnt.values())
print("total", tot)
srcs = {}
for (f, l), c in sorted(cnt.items()):
    if c >= int(os.environ.get("MIN", "4")):
        if f not in srcs:
            p = os.path.join(CSRC, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:70] if l - 1 < len(srcs[f]) else ""
        print(f"{f}:{l:4d} {c:4d}  {' '.join(f'{k}{v}' for k, v in ops[(f, l)].most_common(6)):50s} | {text}")

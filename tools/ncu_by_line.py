"""Development helper: join an ncu SASS dump (tools/ncu_src.py third output) with nvdisasm line info of the same
build, and print executed warp-instructions per unit by source line."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "unbiased-quantization-distributed-mean-estimation_b200", "csrc")
lines_file, kern = sys.argv[1], sys.argv[2]
src = "stream.cu"
cubin = "/tmp/t/join.cubin"
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-cubin",
       "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", cubin, os.path.join(CSRC, src)] + os.environ.get("DME_NVCC_EXTRA", "").split()
subprocess.run(cmd, check=True, capture_output=True)
out = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
seq = []; line = None; inside = False
for ln in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        inside = kern in m.group(1); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
    if m:
        ins = m.group(1).split(); op = ins[1] if ins[0].startswith("@") else ins[0]
        seq.append((line, op))
recs = []
for l in open(lines_file):
    idx, ex, sm, ins = l.split(None, 3)
    t = ins.split(); op = t[1] if t[0].startswith("@") else t[0]
    recs.append((float(ex), int(sm), op))
print(len(seq), len(recs))
n = min(len(seq), len(recs))
bad = sum(1 for i in range(n) if seq[i][1] != recs[i][2])
print("opcode mismatches", bad)
agg = collections.Counter(); smp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for i in range(n):
    agg[seq[i][0]] += recs[i][0]; smp[seq[i][0]] += recs[i][1]; ops[seq[i][0]][seq[i][1].split(".")[0]] += recs[i][0]
srcs = {}
tot = sum(agg.values()); ts = sum(smp.values())
print("total per unit", tot)
for (f, l), c in sorted(agg.items(), key=lambda kv: (kv[0][0] != "stream.cu", kv[0])):
    if c * 16 >= float(os.environ.get("MIN", "2")):
        if f not in srcs:
            p = os.path.join(CSRC, f)
            srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = srcs[f][l - 1].strip()[:64] if l - 1 < len(srcs[f]) else ""
        print(f"{f[:14]:14s}:{l:4d} {c * 16:7.1f} {100 * smp[(f, l)] / ts:5.1f}%  {' '.join(f'{k}{v * 16:.0f}' for k, v in ops[(f, l)].most_common(4)):36s} | {text}")

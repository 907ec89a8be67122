"""Development helper: opcode histogram of one kernel of libdme_b200.so (cuobjdump -sass)."""
import re, subprocess, sys, collections
so = "unbiased-quantization-distributed-mean-estimation_b200/libdme_b200.so"
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None; hist = collections.Counter(); n = 0; lines = []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            ins = m.group(2).strip()
            lines.append(f"{m.group(1)} {ins}")
            op = ins.split()[0]
            if op.startswith("@"):
                op = ins.split()[1]
            hist[op.split(".")[0]] += 1; n += 1
print(n, "instructions")
print(", ".join(f"{k} {v}" for k, v in hist.most_common(45)))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(lines))

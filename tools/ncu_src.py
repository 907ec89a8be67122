"""Development helper: summarise an `ncu --page source --csv --print-source sass` dump: executed instructions per
opcode, top stall reasons, and the hottest instruction ranges."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0; byop = collections.Counter(); samples = collections.Counter(); stall = collections.Counter()
recs = []
for r in body:
    if len(r) < len(hdr): continue
    ins = r[ix["Source"]].strip(); ex = int(r[ix["Instructions Executed"]] or 0); sm = int(r[ix["# Samples"]] or 0)
    toks = ins.split(); op = toks[1] if toks[0].startswith("@") else toks[0]
    byop[op.split(".")[0]] += ex; tot += ex; samples[op.split(".")[0]] += sm
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]] or 0)
    recs.append((ex, sm, ins))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print("total warp instr", tot, "per unit", tot / units)
print("by opcode (per unit):", ", ".join(f"{k} {v / units:.2f}" for k, v in byop.most_common(40)))
ts = sum(samples.values())
print("samples by opcode (%):", ", ".join(f"{k} {100 * v / ts:.1f}" for k, v in samples.most_common(25)))
tt = sum(stall.values())
print("stalls (%):", ", ".join(f"{k[6:]} {100 * v / tt:.1f}" for k, v in stall.most_common(12)))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        for i, (ex, sm, ins) in enumerate(recs):
            f.write(f"{i:5d} {ex / units:8.3f} {sm:6d}  {ins}\n")

"""Development helper: ms per step and per-kernel times of a bench.py JSON line (file argument)."""
import json, sys
for p in sys.argv[1:]:
    l = json.loads(open(p).read().strip().splitlines()[-1])
    r = l["roofline"]
    print(p, "ms/step %.3f" % l["ms_per_step"], "step_frac %.3f" % r["step_frac"], {k: round(v, 3) for k, v in r["kernel_ms"].items()},
          "e2e", (l.get("e2e") or {}).get("ms_per_step"))

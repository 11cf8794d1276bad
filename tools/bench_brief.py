#!/usr/bin/env python
"""Print the few numbers of a bench.py JSON line that matter while tuning (value, e2e, latency, per-kernel ms, parity, gather)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "lat_ms", round((d.get("latency") or {}).get("median_ms", 0), 4),
      "step_ms", {k: round(v, 3) for k, v in d.get("step_ms", {}).items() if k in ("median", "p95")})
print({k: round(v["ms_per_launch"], 3) for k, v in d.get("kernels", {}).items()})
print("parity", d.get("parity"))
for k in ("gather", "c5_batch_gather"):
    if d.get(k):
        print(k, {kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in d[k].items() if kk != "how"})

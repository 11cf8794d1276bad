#!/usr/bin/env python
"""Print the few numbers of a bench.py JSON line that matter while tuning (value, e2e, latency, per-kernel ms)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "lat_ms", round(d.get("latency", {}).get("median_ms", 0), 4))
print({k: round(v["ms_per_launch"], 3) for k, v in d.get("kernels", {}).items()})

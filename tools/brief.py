#!/usr/bin/env python
"""Read bench.py's JSON line on stdin, print a short summary (tag = argv[1])."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")]
if not line:
    print(tag, "no JSON line")
    sys.exit(1)
d = json.loads(line[-1])
e = d.get("e2e") or {}
print(tag, "value", round(d["value"]), "e2e", round(e.get("value", 0)), "e2e_ms", round(e.get("ms_per_step", 0), 1),
      "clk", (d.get("clocks") or {}).get("sm_mhz"))
k = d["roofline"]["kernels"]
print("   ", " ".join(f"{n}={v['ms'] / v['groups']:.2f}ms/{v['GBps']:.0f}" for n, v in k.items()))

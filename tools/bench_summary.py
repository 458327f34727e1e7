#!/usr/bin/env python
"""Compact one-line summaries of bench.py JSON lines read from stdin (tooling for gpurun logs)."""
import json
import sys
for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if d.get("impl") == "reference":
        print("REF fps %.2f" % d["value"]); continue
    print("BENCH %s fps %.1f ms %.3f e2e %.1f roof %.3f hbm %.3f" % (
        d["config"]["precision"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"],
        (d.get("roofline_hbm") or {}).get("frac", 0)))
    print("   ", {k.replace("conv:", ""): v["ms"] for k, v in d["kernels_ms_per_step"].items()})

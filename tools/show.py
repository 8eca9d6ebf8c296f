#!/usr/bin/env python
"""Pretty-print the per-hidden table of a bench.py JSON line read from stdin."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("%s | value %.0f %s | e2e %.0f | parity %s | ds %s" % (d["config"]["workload"], d["value"], d["unit"],
      d["e2e"]["value"], d.get("parity_all_ranks"), d.get("plan", {}).get("ds_parts")))
for p in d["per_hidden"]:
    print("  H=%3d %.3f ms  %.0f GFLOP/s  gather %.1f TB/s  alg %.0f GB/s (%.3f of HBM)" % (
        p["hidden"], p["kernel_ms"], p["gflops"], p["gather_gbs"] / 1e3, p["alg_gbs"], p["frac_hbm"]))

#!/usr/bin/env python
"""print the headline numbers and the per-kernel table of a bench.py JSON line"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.3f M reads/s  e2e %.3f M reads/s  ms/step %.1f  e2e ms/step %.1f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["e2e"].get("ms_per_step", 0)))
for k, v in d["kernels"].items():
    print("%-45s %8.2f ms  share %.3f  %s" % (k, v["ms"], v["share_of_step"] or 0, ("%.0f GB/s" % v["achieved_GBps"]) if v.get("achieved_GBps") else ""))
print("sum of stages %.1f ms per batch" % sum(v["ms"] for v in d["kernels"].values() if "part_of" not in v))
print("parity", d.get("parity")); print("cpu", d.get("cpu_baseline")); print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac")}); print("clocks", d.get("clocks"))

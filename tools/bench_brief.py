#!/usr/bin/env python
"""run bench.py with the given extra args and print one short line (value, e2e, per-kernel ms); env vars pass through"""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v["ms"], 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench failed:", e, out.stderr[-400:])

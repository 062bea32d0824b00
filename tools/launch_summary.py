#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> a table of one batch: kernels by total time with launch counts and
shares.  One batch = the launches from the first `k_encode`-less upload of a batch (first k_smem_fwd) up to the next one.
usage: launch_summary.py launches.csv [batch_index]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), ms))
starts = [i for i, (k, _) in enumerate(rows) if k.startswith("k_smem_fwd<1") or k.startswith("void k_smem_fwd<1")]
b = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lo = starts[b] if starts else 0
hi = starts[b + 1] if len(starts) > b + 1 else len(rows)
while lo > 0 and re.match(r'(void )?(k_encode|k_fq_)', rows[lo - 1][0]): lo -= 1  # the batch's tokenise / encode kernels come before its first seeding kernel
agg = OrderedDict()
for k, ms in rows[lo:hi]:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f %% |" % (k[:70], n, ms, 100 * ms / tot))
print("| **total** | %d | %.3f | |" % (sum(v[0] for v in agg.values()), tot))

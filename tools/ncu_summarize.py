#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name).replace("flm::<unnamed>::", "").replace("void ", "")
    name = re.sub(r"^at::native::.*?(\w+_kernel\w*).*$", r"torch:\1", name)
    ns = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("us", "usecond"):
        ns *= 1e3
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
print("%-70s %7s %12s %7s" % ("kernel", "count", "total_us", "share"))
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-70s %7d %12.1f %6.1f%%" % (k[:70], n, ns / 1e3, 100 * ns / total))
print("%-70s %7d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), total / 1e3))

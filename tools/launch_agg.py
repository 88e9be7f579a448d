"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.  python tools/launch_agg.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{v[1] / 1e3:10.1f} us {v[0]:6d}x {v[1] / v[0] / 1e3:8.1f} us/launch {100 * v[1] / tot:5.1f}%  {k[:90]}")
print(f"total {tot / 1e3:.1f} us")

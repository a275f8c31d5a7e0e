"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/summarize_launches.py file.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(data)} launches, total {tot / 1e3:.1f} us (cold-cache, serialised: compare shares)")
print(f"{'share':>7} {'total us':>10} {'count':>6} {'avg us':>9}  kernel")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{100 * v[1] / tot:6.1f}% {v[1] / 1e3:10.1f} {v[0]:6d} {v[1] / v[0] / 1e3:9.2f}  {k[:120]}")

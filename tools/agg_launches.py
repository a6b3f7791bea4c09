#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/agg_launches.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    a = agg.setdefault(r[ki][:78], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:5d} {a[1] / 1e6:10.3f} ms {100 * a[1] / tot:5.1f}%  avg {a[1] / a[0] / 1e3:9.1f} us  {k}")

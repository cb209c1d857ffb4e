#!/usr/bin/env python
"""Average duration per kernel of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_agg.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            agg[d["Kernel Name"][:90]].append(float(d["Metric Value"].replace(",", "")))
        except ValueError:
            pass
for k, v in agg.items():
    print("%4d launches  %8.1f us  %s" % (len(v), sum(v) / len(v) / 1000.0, k))

#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source sass --csv`: stall reasons, opcode mix and the hottest loop.

    ncu -i prof.ncu-rep --page source --print-source sass --csv > sass.csv ; python tools/ncu_sass_summary.py sass.csv [--loop]
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
src = [r[ix["Source"]].strip() for r in data]
sm = [int(r[ix["# Samples"]]) for r in data]
ex = [int(r[ix["Instructions Executed"]]) for r in data]
tot = sum(sm)
print(rows[0][1])
print("samples", tot, "| SASS instructions", len(data), "| warp instructions executed", sum(ex))
agg = Counter()
for r in data:
    for h in stalls:
        agg[h[6:]] += int(r[ix[h]])
print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in agg.most_common(9)))
c, cs = Counter(), Counter()
for s_, e, m in zip(src, ex, sm):
    t = s_.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += e
    cs[op] += m
te = sum(c.values())
print("opcode      %executed  %samples")
for op, e in c.most_common(18):
    print("%-10s %9.1f %9.1f" % (op, 100.0 * e / te, 100.0 * cs[op] / tot))
mx = max(ex)
hot = [i for i, e in enumerate(ex) if e >= mx * 0.9]
print("hottest loop: SASS %d..%d (%d instructions, %.0f executions each): %.1f%% of samples, %.1f%% of executed instructions"
      % (hot[0], hot[-1], len(hot), mx, 100.0 * sum(sm[hot[0]:hot[-1] + 1]) / tot, 100.0 * sum(ex[hot[0]:hot[-1] + 1]) / te))
if "--loop" in sys.argv:
    for i in range(hot[0], hot[-1] + 1):
        r = data[i]
        top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("%5d %-70s %6d  %s" % (i, src[i][:70], sm[i], " ".join("%s:%d" % (n, v) for v, n in top if v)))

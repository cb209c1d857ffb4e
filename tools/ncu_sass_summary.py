#!/usr/bin/env python
"""Summarise the SASS page of an ncu report, one block per kernel in the report: tensor / TMEM / async instruction counts, opcode mix, where
the stall samples fall, and the hottest straight-line ranges.

    ncu -i prof.ncu-rep --page source --print-source sass --csv > sass.csv ; python tools/ncu_sass_summary.py sass.csv
"""
import csv
import sys
from collections import Counter

WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "UTCATOMSWS", "ELECT", "NANOSLEEP", "HMMA", "LDSM", "FFMA", "FMUL", "FADD",
         "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR", "F2FP", "FHFMA", "FMNMX", "PRMT", "SHFL"]


def opcode(src):
    t = src.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for n, hi in enumerate(heads):
    name = rows[hi - 1][1] if hi > 0 and rows[hi - 1] and rows[hi - 1][0] == "Kernel Name" else "?"
    end = (heads[n + 1] - 1) if n + 1 < len(heads) else len(rows)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    src = [r[ix["Source"]].strip() for r in data]
    sm = [int(r[ix["# Samples"]]) for r in data]
    ex = [int(r[ix["Instructions Executed"]]) for r in data]
    tot_s, tot_e = max(sum(sm), 1), max(sum(ex), 1)
    if (name, tot_e, tot_s) in seen:        # ncu prints a result once per source view
        continue
    seen.add((name, tot_e, tot_s))
    print(name)
    print("  SASS instructions %d, warp instructions executed %d, stall samples %d" % (len(data), sum(ex), sum(sm)))
    c, cs = Counter(), Counter()
    for s_, e, m in zip(src, ex, sm):
        op = opcode(s_)
        c[op] += e
        cs[op] += m
    print("  tensor / TMEM / async and other instructions of interest, executed (warp level):")
    for op in WATCH:
        if c[op]:
            print("    %-12s %10d  (%.2f%% of instructions, %.1f%% of stall samples)" % (op, c[op], 100.0 * c[op] / tot_e, 100.0 * cs[op] / tot_s))
    print("  top opcodes by executed count: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_e) for k, v in c.most_common(12)))
    print("  top opcodes by stall samples : " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_s) for k, v in cs.most_common(12)))
    # straight-line ranges (equal execution counts) that carry more than 2 % of the executed instructions
    out, prev, start = [], None, 0
    for i, e in enumerate(ex + [None]):
        if prev is None:
            prev, start = e, i
            continue
        if e is None or abs(e - prev) > 0.02 * max(prev, 1):
            out.append((start, i - 1, prev, sum(sm[start:i])))
            prev, start = e, i
    print("  hottest ranges (SASS index range, instructions, executions each, share of executed instructions, share of stall samples, first instruction):")
    for a, b, e, s_ in out:
        if (b - a + 1) * e > 0.02 * tot_e:
            print("    %5d..%-5d %4d x %8d  %5.1f%%  %5.1f%%   %s" % (a, b, b - a + 1, e, 100.0 * (b - a + 1) * e / tot_e, 100.0 * s_ / tot_s, src[a][:60]))
    print()

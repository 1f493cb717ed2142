#!/usr/bin/env python
"""Opcode mix of a kernel weighted by executed instructions, from `ncu -i rep --page source --csv --print-source sass`.
usage: tools/sass_mix.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE:
        continue
    op = r[iS].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    base = o.split(".")[0]
    key = base
    if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL"):
        key = ".".join(o.split(".")[:1] + [x for x in o.split(".")[1:] if x in ("64", "128")])
    n = int(r[iE] or 0)
    mix[key] += n
    stall[key] += int(r[iT] or 0)
    tot += n
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print("total warp instructions", tot)
ts = sum(stall.values())
for k, v in mix.most_common(top):
    print("%-14s %12d %5.1f%%   stall samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * stall[k] / max(ts, 1)))

#!/usr/bin/env python
"""Opcode mix of one kernel from an `ncu --page source --csv` export (SASS view): executed warp instructions and
stall samples per opcode, the figures behind the "instructions per point" numbers in DESIGN.md.

    python tools/ncu_source_mix.py gpurun_out/rN/source_k_sweepA.csv [points]
"""
import csv
import collections
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    points = float(sys.argv[2]) if len(sys.argv) > 2 else 256.0 ** 3
    hdr = rows[1]
    iS, iI, iP = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix = collections.Counter()
    smp = collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= iI:
            continue
        toks = r[iS].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDG", "STG", "LDS", "STS")) and "." in op else "")
        n = int(r[iI] or 0)
        mix[op] += n
        smp[op] += int(r[iP] or 0)
        tot += n
    ts = sum(smp.values())
    print(f"kernel: {rows[0][1][:100]}")
    print(f"warp instructions {tot:,}  = {tot * 32 / points:.0f} thread instructions per point; samples {ts}")
    print(f"{'opcode':14s} {'warp inst':>14s} {'share':>7s} {'per point':>10s} {'samples':>8s} {'share':>7s}")
    for op, n in mix.most_common(28):
        print(f"{op:14s} {n:14,d} {100.0 * n / tot:6.1f}% {n * 32 / points:10.1f} {smp[op]:8d} {100.0 * smp[op] / max(ts, 1):6.1f}%")


if __name__ == "__main__":
    main()

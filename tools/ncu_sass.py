#!/usr/bin/env python
"""SASS-level totals of an `ncu --page source --csv --print-source cuda,sass` export: executed warp instructions by
opcode and by the CUDA line they sit under (inline stacks attribute one instruction to several CUDA lines, so
only SASS rows are summed).  usage: python tools/ncu_sass.py src.csv [topN]"""
import csv
import sys
from collections import defaultdict


def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = None
    seen = set()
    by_op, by_line, stall = defaultdict(float), defaultdict(float), defaultdict(float)
    cur_line = None
    tot = 0.0
    tot_s = 0.0
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            iI, iS, iA = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
            continue
        if hdr is None or len(r) <= iI:
            continue
        if r[0] != "":
            cur_line = (r[0], r[1].strip()[:90])
            continue
        addr = r[iA]
        if not addr.startswith("0x") or addr in seen:
            continue
        seen.add(addr)
        n, s = f(r[iI]), f(r[iS])
        op = r[3].split()[0] if r[3].split() else "?"
        if op.startswith("@"):
            op = r[3].split()[1]
        by_op[op.split(".")[0]] += n
        by_line[cur_line] += n
        stall[cur_line] += s
        tot += n
        tot_s += s
    print(f"total warp instructions {tot:.0f}, samples {tot_s:.0f}")
    print("-- by opcode")
    for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v / tot * 100:6.2f}%  {k}")
    print("-- by innermost CUDA line")
    for k, v in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{v / tot * 100:6.2f}% inst {stall[k] / max(tot_s, 1) * 100:6.2f}% smp  {k[0]:>5} {k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

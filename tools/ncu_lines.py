#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [topN]"""
import csv
import sys


def f(v):
    try:
        return float(v)
    except ValueError:
        return 0.0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = {"file": r[1], "rows": []}
            sections.append(cur)
        elif len(r) >= 2 and r[0] == "Function Name" and cur is not None:
            cur["fn"] = r[1]
        elif r and r[0] == "Line No" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and "hdr" in cur:
            cur["rows"].append(r)
    for s in sections:
        h = s["hdr"]
        iL, iS, iI, iSm = h.index("Line No"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(f(r[iI]) for r in s["rows"] if len(r) > iI) or 1.0
        tsm = sum(f(r[iSm]) for r in s["rows"] if len(r) > iSm) or 1.0
        print(f"== {s['file']} :: {s.get('fn')}  total warp-inst {tot:.0f}  samples {tsm:.0f}")
        agg = {}
        for r in s["rows"]:
            if len(r) <= iI:
                continue
            try:
                ln = int(r[iL])
            except ValueError:
                continue
            a = agg.setdefault(ln, [r[iS], 0.0, 0.0])
            a[1] += f(r[iI])
            a[2] += f(r[iSm])
        for ln, (src, ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"{ln:5d} {ins / tot * 100:5.1f}% inst {smp / tsm * 100:5.1f}% smp | {src.strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

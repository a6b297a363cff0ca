#!/usr/bin/env python
"""Print a flat_sweep.sh result file as a table.  usage: python tools/sweep_table.py gpurun_out/x_sweep.jsonl"""
import json
import sys

cur = None
for l in open(sys.argv[1]).read().split("\n"):
    if l.startswith("#"):
        cur = l[2:]
    elif l.startswith("{"):
        d = json.loads(l)
        print(f"{cur:62s} fwd {d['fwd_us']:7.2f} ({d['fwd_gbps']:6.0f}) bwd {d['bwd_us']:7.2f} ({d['bwd_gbps']:6.0f}) "
              f"frac {d['frac']:.3f} P={d['f_cs']}/{d['b_cs']} D={d['f_slots']}/{d['b_slots']}")

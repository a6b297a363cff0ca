#!/usr/bin/env python
"""Print a sweep result file as a table.
usage: python tools/sweep_table.py gpurun_out/x_sweep.jsonl            (tools/flat_sweep.sh knob sweeps)
       python tools/sweep_table.py --microbench gpurun_out/sweep.jsonl (bench.py --sweep --sweep-out: profiles/r0N_microbench_sweep.txt)"""
import json
import sys

if sys.argv[1] == "--microbench":
    print("# BASELINE.json configs[3]: instance_cond microbench sweep on one B200 (bench.py --sweep), device-resident inputs,")
    print("# CUDA events over back-to-back launches, rotating buffer sets; frac = (2+3)*E*s / (fwd+bwd) / 6542.1 GB/s (measured copy peak).")
    print("# path (of the backward, the last call): 0 = small (warp/CTA per slab), 1 = cluster, 2 = flat, 4 = resident (cluster per slab, shared-memory resident)")
    print(f"{'dtype':5s} {'S':>4s} {'N':>2s} {'C':>4s} {'path':>4s} {'fwd_us':>10s} {'bwd_us':>10s} {'fwd GB/s':>9s} {'bwd GB/s':>9s} {'frac':>6s}")
    for l in open(sys.argv[2]):
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        if "skipped" in d:
            print(f"{d['dtype']:5s} {d['S']:4d} {d['N']:2d} {d['C']:4d}  skipped: {d['skipped']}")
            continue
        print(f"{d['dtype']:5s} {d['S']:4d} {d['N']:2d} {d['C']:4d} {d['path']:4d} {d['fwd_us']:10.2f} {d['bwd_us']:10.2f} "
              f"{d['fwd_gbps']:9.1f} {d['bwd_gbps']:9.1f} {d['frac']:6.3f}")
    sys.exit(0)

cur = None
for l in open(sys.argv[1]).read().split("\n"):
    if l.startswith("#"):
        cur = l[2:]
    elif l.startswith("{"):
        d = json.loads(l)
        print(f"{cur:62s} fwd {d['fwd_us']:7.2f} ({d['fwd_gbps']:6.0f}) bwd {d['bwd_us']:7.2f} ({d['bwd_gbps']:6.0f}) "
              f"frac {d['frac']:.3f} P={d['f_cs']}/{d['b_cs']} D={d['f_slots']}/{d['b_slots']}")

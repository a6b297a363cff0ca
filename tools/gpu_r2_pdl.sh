#!/bin/bash
B="--steps 20 --warmup 5 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls"
run() { timeout 200 python bench.py $B "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  us/step %.2f  frac %.3f  roofline bwd %.2f fwd %.2f' % (d['ms_per_step']*1e3, d['frac_of_peak'], d['roofline']['avg_launch_us'], d['roofline']['fwd']['avg_launch_us']))
    elif 'rror' in l: print('  ', l.strip()[:200])
"; }
echo "graph, default"; run
echo "graph, pdl + coop"; MICN_BENCH_OPTS="flat_pdl=1" run
echo "graph, pdl, no coop"; MICN_BENCH_OPTS="flat_pdl=1,flat_coop=0" run
echo "stream, default"; run --launch stream
echo "stream, pdl + coop"; MICN_BENCH_OPTS="flat_pdl=1" run --launch stream
echo "stream, pdl, no coop"; MICN_BENCH_OPTS="flat_pdl=1,flat_coop=0" run --launch stream
echo "graph, no coop only"; MICN_BENCH_OPTS="flat_coop=0" run

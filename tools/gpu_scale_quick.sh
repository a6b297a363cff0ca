#!/bin/bash
N=${1:-2}
B="--gpus $N --steps 20 --warmup 5 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py $B "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  us/step %.2f  per rank %s  total GB/s %.0f  check %s' % (d['ms_per_step']*1e3, d['per_rank_us_per_step'], d['value'], d['allreduce_check_rel_err']))
    elif 'Error' in l: print('  ', l.strip()[:200])
"; }
echo "fused (lagged fold)"; run
echo "fused, plain stores (xchg_dbg=8)"; MICN_BENCH_OPTS="xchg_dbg=8" run
echo "fused, synchronous fold in the roofline-style loop is not timed here"

#!/bin/bash
N=${1:-2}
B="--gpus $N --steps 20 --warmup 5 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls"
run() { if [ $N -gt 1 ]; then L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516"; else L=python; fi
  timeout 300 $L bench.py $B "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  us/step %.2f  per rank %s  total GB/s %.0f  clocks %s' % (d['ms_per_step']*1e3, d['per_rank_us_per_step'], d['value'], d['clocks']))
    elif 'Error' in l: print('  ', l.strip()[:200])
"; }
echo "N=$N default"; run
echo "N=$N default (again)"; run

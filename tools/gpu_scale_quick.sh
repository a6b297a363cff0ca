#!/bin/bash
# headline only at N GPUs (the driver's command without the model-level legs), twice, + the multi-GPU exchange check
N=${1:-2}
B="--gpus $N --steps 20 --warmup 5 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 bench.py $B "$@" 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  us/step %.2f  total GB/s %.0f  per rank %s check %s clocks %s' % (d['ms_per_step']*1e3, d['value'], d['per_rank_us_per_step'], d['allreduce_check_rel_err'], d['clocks']))
    elif 'Error' in l: print('  ', l.strip()[:200])
"; }
echo "N=$N fused"; run; run
echo "N=$N nccl async"; run --collective nccl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/peer_xchg_test.py 2>&1 | grep "PEER EXCHANGE\|FAIL" | sort | uniq -c | head

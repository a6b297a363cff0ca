#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it, with a hard timeout
N=${1:-2}; tag=${2:-s}
o=gpurun_out; mkdir -p $o
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $o/${tag}_bench_n$N.log 2> $o/${tag}_bench_n$N.err; echo "exit $?"
tail -c 800 $o/${tag}_bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench_n$N.log") if l.startswith("{")][-1])
print("N", d["n_gpus"], "value", d["value"], "ms_per_step", d["ms_per_step"], "regions", d["regions_ms"], "check", d["allreduce_check_rel_err"])
print("e2e", json.dumps(d.get("e2e"))[:300])
for k in ("model_step","sliding_window","model_legs_error"):
    print(k, json.dumps(d.get(k))[:1600])
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/peer_xchg_test.py 2>&1 | grep "PEER EXCHANGE\|FAIL\|peer buffers" | sort | uniq -c | head

#!/bin/bash
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-torch-ref"
o=gpurun_out/g2_probe.log
: > $o
echo "# default" >> $o; $T 2>&1 | grep '^{' | cut -c1-220 >> $o
echo "# no allreduce" >> $o; MICN_BENCH_NO_ALLREDUCE=1 $T 2>&1 | grep '^{' | cut -c1-220 >> $o
echo "# plain (non-cooperative) launch" >> $o; MICN_BENCH_OPTS=flat_coop=0 $T 2>&1 | grep '^{' | cut -c1-220 >> $o
echo "# cluster path" >> $o; MICN_BENCH_OPTS=force_path=1 $T 2>&1 | grep '^{' | cut -c1-220 >> $o
cat $o

#!/bin/bash
tag=${1:-f}
o=gpurun_out
mkdir -p $o
MICN_SHAPE=1,48,48 python tools/res_trace_probe.py > $o/${tag}_trace_48.log 2>&1; cat $o/${tag}_trace_48.log | cut -c1-260
for cp in 1 2 4 8; do echo "copies=$cp"; MICN_OPTS="res_copies=$cp" MICN_SHAPE=1,48,48 python tools/res_trace_probe.py 2>&1 | grep "back-to-back\|rep 1" | cut -c1-260; done
LD_LIBRARY_PATH=tools/trace tools/micn_selftest --suite trace --N 1 --C 48 --S 96 --dtype bf16 > $o/${tag}_flat_trace_fwd.log 2>&1
LD_LIBRARY_PATH=tools/trace tools/micn_selftest --suite trace --N 1 --C 48 --S 96 --dtype bf16 --opt flat_trace_which=2 > $o/${tag}_flat_trace_bwd.log 2>&1
head -30 $o/${tag}_flat_trace_fwd.log
MICN_EXTRA="24x48,96x48,192x24,48x64" timeout 600 python tools/calls_graph_probe.py > $o/${tag}_probe_res.log 2>&1; cat $o/${tag}_probe_res.log | cut -c1-150

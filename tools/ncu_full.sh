#!/bin/bash
# one full-set capture (with source) of the flat forward and backward kernels at the headline shape
tag=${1:-f}
ONE="tools/micn_selftest --suite one --N 1 --C 48 --S 96 --dtype bf16 --iters 20"
$ONE > gpurun_out/${tag}_one.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fwd_flat -s 4 -c 1 -f -o gpurun_out/${tag}_fwd $ONE > gpurun_out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bwd_flat -s 4 -c 1 -f -o gpurun_out/${tag}_bwd $ONE >> gpurun_out/${tag}_ncu.log 2>&1
tail -1 gpurun_out/${tag}_one.log

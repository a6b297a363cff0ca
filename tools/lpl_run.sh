#!/bin/bash
# bring-up: correctness, then the lag / piece-size knobs at the headline shape
tag=${1:-v11}
o=gpurun_out/${tag}_lpl.log
: > $o
timeout 600 tools/micn_selftest --suite correctness > gpurun_out/${tag}_selftest.log 2>&1; echo "selftest exit $?" >> gpurun_out/${tag}_selftest.log
tail -4 gpurun_out/${tag}_selftest.log
H="--N 1 --C 48 --S 96"
for dt in bf16 fp32; do
  for l in 2 3 4 5; do
    echo "# $dt L=$l" >> $o
    timeout 120 tools/micn_selftest --suite one $H --dtype $dt --flag $l | grep '^{' >> $o
  done
  for v in 512 768 1024; do
    echo "# $dt fpv=$v" >> $o
    timeout 120 tools/micn_selftest --suite one $H --dtype $dt --fpv $v | grep '^{' >> $o
  done
  echo "# $dt N=4" >> $o
  timeout 120 tools/micn_selftest --suite one --N 4 --C 48 --S 96 --dtype $dt | grep '^{' >> $o
done
tools/micn_selftest --suite trace $H --dtype bf16 > gpurun_out/${tag}_trace.log 2>&1

#!/bin/bash
# quick headline check (bring-up): default plan, bf16 and fp32, plus N=4
o=gpurun_out/${1:-q}_quick.log
: > $o
for a in "--N 1 --C 48 --S 96 --dtype bf16" "--N 1 --C 48 --S 96 --dtype fp32" "--N 4 --C 48 --S 96 --dtype bf16" "--N 1 --C 48 --S 96 --dtype bf16 --epi 1" "--N 1 --C 48 --S 96 --dtype bf16 --epi 2" "--N 8 --C 24 --S 48 --dtype bf16" $2; do
  echo "# $a" >> $o; timeout 120 tools/micn_selftest --suite one $a | grep '^{' >> $o
done
cat $o

#!/bin/bash
# bring-up: the per-piece overhead the planner should assume (vector equivalents)
o=gpurun_out/${1:-ring}_sweep.jsonl
: > $o
run() { echo "# $*" >> $o; timeout 120 tools/micn_selftest --suite one "$@" | grep '^{' >> $o; }
for shape in "--N 1 --C 48 --S 96 --dtype bf16" "--N 1 --C 48 --S 96 --dtype fp32" "--N 4 --C 48 --S 96 --dtype bf16" "--N 4 --C 96 --S 48 --dtype bf16" "--N 1 --C 24 --S 128 --dtype bf16" "--N 8 --C 24 --S 48 --dtype bf16" "--N 1 --C 96 --S 48 --dtype bf16"; do
  for ov in 128 800 1100 1600 2200 3000 4500; do run $shape --fovh $ov; done
done

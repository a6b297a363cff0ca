#!/bin/bash
# Flat-path knob sweep on the GPU box (bring-up aid; results land in gpurun_out/).
out=${1:-gpurun_out/flat_sweep.jsonl}
: > $out
run() { echo "# $*" >> $out; timeout 120 tools/micn_selftest --suite one "$@" | grep '^{' >> $out; }
for dt in bf16 fp32; do
  run --N 1 --C 48 --S 96 --dtype $dt
  for k in 3 4 5 6 8; do run --N 1 --C 48 --S 96 --dtype $dt --fslots $k; done
  for l in 1 2 3 4 8 12; do run --N 1 --C 48 --S 96 --dtype $dt --flag $l; done
  run --N 1 --C 48 --S 96 --dtype $dt --fslots 4 --flag 4
  run --N 1 --C 48 --S 96 --dtype $dt --fslots 4 --flag 8
  run --N 1 --C 48 --S 96 --dtype $dt --fslots 8 --flag 8
  run --N 1 --C 48 --S 96 --dtype $dt --fpd 0
  run --N 1 --C 48 --S 96 --dtype $dt --fpd 2500
  run --N 4 --C 48 --S 96 --dtype $dt
done
run --N 4 --C 96 --S 48 --dtype bf16
run --N 4 --C 96 --S 48 --dtype fp32
run --N 1 --C 24 --S 128 --dtype bf16
run --N 1 --C 24 --S 128 --dtype fp32
run --N 1 --C 48 --S 96 --dtype bf16 --epi 1
run --N 1 --C 48 --S 96 --dtype bf16 --epi 2
cat $out

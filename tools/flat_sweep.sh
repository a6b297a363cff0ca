#!/bin/bash
# Flat-path knob sweep on the GPU box (bring-up aid; results land in gpurun_out/).
out=${1:-gpurun_out/flat_sweep.jsonl}
: > $out
run() { echo "# $*" >> $out; timeout 120 tools/micn_selftest --suite one "$@" | grep '^{' >> $out; }
for dt in bf16 fp32; do
  run --N 1 --C 48 --S 96 --dtype $dt
  for cfg in "1 4 6" "1 6 9" "1 8 12" "2 2 8" "2 3 10" "2 3 12" "2 4 12" "2 4 16" "2 5 14" "4 1 12" "4 2 16" "4 3 20" "8 1 20" "8 1 24"; do
    set -- $cfg
    run --N 1 --C 48 --S 96 --dtype $dt --fgroups $1 --flag $2 --fslots $3
  done
  run --N 1 --C 48 --S 96 --dtype $dt --fgroups 2 --flag 3 --fslots 12 --fpd 0
  run --N 1 --C 48 --S 96 --dtype $dt --fgroups 2 --flag 3 --fslots 12 --fpd 1500
  run --N 1 --C 48 --S 96 --dtype $dt --fgroups 2 --flag 3 --fslots 12 --fpb 50
  run --N 4 --C 48 --S 96 --dtype $dt
done
run --N 4 --C 96 --S 48 --dtype bf16
run --N 4 --C 96 --S 48 --dtype fp32
run --N 1 --C 24 --S 128 --dtype bf16
run --N 1 --C 24 --S 128 --dtype fp32
run --N 1 --C 48 --S 96 --dtype bf16 --epi 1
run --N 1 --C 48 --S 96 --dtype bf16 --epi 2
cat $out

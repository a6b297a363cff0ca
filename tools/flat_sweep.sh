#!/bin/bash
# Flat-path knob sweep on the GPU box (bring-up aid; results land in gpurun_out/).
out=${1:-gpurun_out/flat_sweep.jsonl}
: > $out
run() { echo "# $*" >> $out; timeout 120 tools/micn_selftest --suite one "$@" | grep '^{' >> $out; }
H="--N 1 --C 48 --S 96"
for dt in bf16 fp32; do
  run $H --dtype $dt
  for l in 2 3 4 6; do run $H --dtype $dt --flag $l; done
  for d in 0 1000 1500 2500; do run $H --dtype $dt --fpd $d; done
  run $H --dtype $dt --fslots 3 --opt flat_slots_b=2
  run $H --dtype $dt --fslots 2 --opt flat_slots_b=1
  run $H --dtype $dt --fslots 3 --opt flat_slots_b=1
  run $H --dtype $dt --fpb 50
  run --N 4 --C 48 --S 96 --dtype $dt
done
run --N 4 --C 96 --S 48 --dtype bf16
run --N 4 --C 96 --S 48 --dtype fp32
run --N 1 --C 24 --S 128 --dtype bf16
run --N 1 --C 24 --S 128 --dtype fp32
run --N 8 --C 24 --S 48 --dtype bf16
run --N 1 --C 384 --S 48 --dtype bf16
run $H --dtype bf16 --epi 1
run $H --dtype bf16 --epi 2
cat $out

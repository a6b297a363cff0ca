#!/bin/bash
# Planner knobs of the flat path at the headline shape, with the kernels launched as programmatic dependents (the bench's
# mode): fwd+bwd pair replayed from a CUDA graph (tools/calls_graph_probe.py).  One line per variant.
mkdir -p gpurun_out
base="flat_pdl=1,flat_coop=0"
for v in "" "flat_shape_bwd=2" "flat_shape_fwd=1" "flat_ovh_vecs=1000" "flat_ovh_vecs=4000" "flat_slots=3" "flat_slots_b=1" "flat_slots_b=3" "flat_grid=144"; do
  line=$(MICN_ONLY=48x96 MICN_OPTS="$base,$v" timeout 90 python tools/calls_graph_probe.py 2>&1 | grep -E "^[a-z].* 48 +96 " | head -1)
  echo "[$v] $line"
done | tee gpurun_out/minisweep.txt

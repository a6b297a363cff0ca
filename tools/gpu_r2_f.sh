#!/bin/bash
# 128^3 fp32 backward dip: L2 budget / lag sweep (item 8)
o=gpurun_out
for opt in "" "--opt flat_l2_mb=16" "--opt flat_l2_mb=48" "--opt flat_l2_mb=64" "--opt flat_lag=3" "--opt flat_lag=6" "--opt flat_slots=3" ; do
  for shape in "--N 4 --C 96 --S 128 --dtype fp32" "--N 4 --C 96 --S 128 --dtype bf16" "--N 1 --C 48 --S 96 --dtype bf16"; do
    echo "# $shape $opt"; timeout 120 tools/micn_selftest --suite one $shape --iters 20 $opt | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   fwd %.1f us (%.0f GB/s)  bwd %.1f us (%.0f GB/s)  lag %s' % (d.get('fwd_us',0), d.get('fwd_gbps',0), d.get('bwd_us',0), d.get('bwd_gbps',0), d.get('lag', d.get('L','?'))))"
  done
done

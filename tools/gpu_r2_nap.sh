#!/bin/bash
# nap-duration variants of the parked mbarrier waits (MICN_PARK_NS consumers / MICN_IDLE_NS helper warps), headline + two more shapes
for d in "" tools/alt_p16_i200 tools/alt_p64_i400 tools/alt_p128_i800 tools/alt_p32_i600; do
  echo "## ${d:-default (32/200)}"
  for rep in 1 2; do
  for shape in "--N 1 --C 48 --S 96 --dtype bf16" "--N 1 --C 48 --S 96 --dtype fp32" "--N 4 --C 48 --S 96 --dtype bf16"; do
    LD_LIBRARY_PATH=$d timeout 120 tools/micn_selftest --suite one $shape --iters 30 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   %s %s x%s  fwd %.2f us  bwd %.2f us  frac %.3f' % (d['dtype'], d['C'], d['N'], d['fwd_us'], d['bwd_us'], d['frac']))"
  done; done
done

#!/bin/bash
# round 2, GPU pass B: parity suite (all failures listed), per-shape pair times with and without the resident path
tag=${1:-b}
o=gpurun_out
mkdir -p $o
timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_r2.py::test_more_than_2_pow_31_elements > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -n 5 $o/${tag}_pytest.log | cut -c1-300
MICN_EXTRA="24x48,96x48,192x24,48x64" timeout 600 python tools/calls_graph_probe.py > $o/${tag}_probe_res.log 2>&1
MICN_OPTS="res_off=1" MICN_EXTRA="24x48,96x48,192x24,48x64" timeout 600 python tools/calls_graph_probe.py > $o/${tag}_probe_nores.log 2>&1
echo "--- with resident path"; cat $o/${tag}_probe_res.log | cut -c1-150
echo "--- without"; cat $o/${tag}_probe_nores.log | cut -c1-150

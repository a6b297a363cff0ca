#!/bin/bash
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_r2.py::test_more_than_2_pow_31_elements > $o/q_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $o/q_pytest.log | cut -c1-200
MICN_TEST_FLAT_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py -m gpu -q -x --deselect tests/test_gpu_parity_r2.py::test_more_than_2_pow_31_elements > $o/q_pytest_pdl.log 2>&1; echo "pytest (flat_pdl=1) exit $?"; tail -3 $o/q_pytest_pdl.log | cut -c1-200
echo "--- probe, pdl on (default)"; MICN_EXTRA="24x48,96x48" timeout 300 python tools/calls_graph_probe.py 2>&1 | cut -c1-120
echo "--- probe, pdl off"; MICN_OPTS="pdl=0" MICN_EXTRA="24x48,96x48" timeout 300 python tools/calls_graph_probe.py 2>&1 | cut -c1-120
echo "--- probe, pdl on + flat_pdl"; MICN_OPTS="flat_pdl=1,flat_coop=0" MICN_EXTRA="24x48,96x48" timeout 300 python tools/calls_graph_probe.py 2>&1 | cut -c1-120

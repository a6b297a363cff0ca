#!/bin/bash
tag=${1:-e}
o=gpurun_out
mkdir -p $o
timeout 1800 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_r2.py::test_more_than_2_pow_31_elements > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest.log
tail -n 5 $o/${tag}_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref > $o/${tag}_bench.log 2> $o/${tag}_bench.err; echo "bench exit $?"
tail -c 400 $o/${tag}_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.log") if l.startswith("{")][-1])
print("value",d["value"],"ms_per_step",d["ms_per_step"],"regions",d["regions_ms"])
print("roofline", d["roofline"]["avg_launch_us"], d["roofline"]["fwd"]["avg_launch_us"])
print(json.dumps(d["next_rows"])[:2500])
print(json.dumps(d["swin_unetr_norm_calls"])[:800])
PY
MICN_SHAPE=1,48,48 python tools/res_trace_probe.py > $o/${tag}_trace_48.log 2>&1; cat $o/${tag}_trace_48.log

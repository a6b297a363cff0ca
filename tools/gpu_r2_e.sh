#!/bin/bash
tag=${1:-k}
o=gpurun_out
mkdir -p $o
timeout 1500 python -m pytest tests -m gpu -q -x > $o/${tag}_pytest_cpp.log 2>&1; echo "pytest(cpp binding) exit $?" >> $o/${tag}_pytest_cpp.log
tail -n 4 $o/${tag}_pytest_cpp.log | cut -c1-300
true
true
python tools/py_overhead_probe.py 2>&1 | head -8
timeout 900 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.log 2> $o/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.log") if l.startswith("{")][-1])
print("value",d["value"],"ms_per_step",d["ms_per_step"])
print(json.dumps(d["next_rows"])[:1200])
print(json.dumps(d["swin_unetr_norm_calls"])[:600])
for k in ("model_step","model_step_unetr"):
    v=d.get(k,{})
    print(k, {kk: (vv.get("ms_per_step") if isinstance(vv,dict) else vv) for kk,vv in v.items() if kk in ("reference","ours","speedup")})
PY

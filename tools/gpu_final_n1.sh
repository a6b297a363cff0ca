#!/bin/bash
o=gpurun_out; tag=${1:-fin}
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $o/${tag}_smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench_ref.log 2>&1; echo "ref exit $?"
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $o/${tag}_bench.log 2> $o/${tag}_bench.err; echo "bench exit $? after $SECONDS s"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${tag}_bench.log") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/${tag}_bench_ref.log") if l.startswith("{")][-1])
print("value",d["value"],"ms_per_step",d["ms_per_step"],"frac",d["frac_of_peak"],"regions",d["regions_ms"], "launches", d["gpu_launches"])
print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","traffic","avg_launch_us")}, "fwd", d["roofline"]["fwd"])
print("e2e", d["e2e"]["value"], "ref", r["value"], "e2e ratio", d["e2e"]["value"]/r["value"], "cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
print("clocks", d["clocks"])
print("calls", {k:round(v,3) if isinstance(v,float) else v for k,v in d["swin_unetr_norm_calls"].items() if k!="what"})
print("next_rows", json.dumps(d["next_rows"])[:1500])
for k in ("model_step","model_step_unetr","sliding_window","model_step_unet_cpu","model_legs_error"):
    print(k, json.dumps(d.get(k))[:1800])
PY

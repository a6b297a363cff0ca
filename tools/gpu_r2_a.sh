#!/bin/bash
# round 2, GPU pass A: full parity suite, smoke, bench (driver-style and default), graph-launch variant
tag=${1:-a}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $o/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_driver.log 2> $o/${tag}_bench_driver.err; echo "bench exit $?" >> $o/${tag}_bench_driver.err
timeout 300 python bench.py --steps 20 --warmup 5 --launch graph --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls > $o/${tag}_bench_graph.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 10 --model-steps none --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls > $o/${tag}_bench_200.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_ref.log 2>&1
for f in pytest smoke; do tail -n 3 $o/${tag}_$f.log | cut -c1-400; done
tail -c 600 $o/${tag}_bench_driver.err
python - <<'PY'
import json,glob,sys
tag=sys.argv[1] if len(sys.argv)>1 else 'a'
for f in sorted(glob.glob('gpurun_out/%s_bench_*.log'%tag)):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f,'unparsed',e); continue
    print(f, 'value',round(d['value'],1),'ms_per_step',d['ms_per_step'],'regions',d.get('regions_ms'),'frac',d.get('frac_of_peak'))
    for k in ('model_step','model_step_unetr','sliding_window','model_step_unet_cpu','model_legs_error'):
        if k in d: print('  ',k, json.dumps(d[k])[:1200])
PY

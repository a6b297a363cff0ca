#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), then the ncu launch list and a full capture of the
# flat kernels (each ncu run only after the same command exited 0 without ncu).  Results -> gpurun_out/<tag>_*.
tag=${1:-run}
o=gpurun_out
mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $o/${tag}_smi.log 2>&1
python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1; echo "smoke exit $?" >> $o/${tag}_smoke.log
python bench.py > $o/${tag}_bench.log 2>&1; echo "bench exit $?" >> $o/${tag}_bench.log
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_ref.log 2>&1; echo "ref exit $?" >> $o/${tag}_bench_ref.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls"
$B > $o/${tag}_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu_launches.log 2>&1
for f in pytest smoke bench; do tail -n 3 $o/${tag}_$f.log | cut -c1-300; done
# full-set captures of the forward and backward flat kernels for profiles/
bash tools/ncu_full.sh ${tag}

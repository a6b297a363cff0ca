#!/bin/bash
# round-2 profile pass: launch list of the bench step + one `ncu --set full` capture per kernel family (each only after the
# same command exited 0 without ncu)
o=gpurun_out
mkdir -p $o
B="python bench.py --steps 5 --warmup 3 --launch stream --no-cpu-baseline --no-e2e --no-torch-ref --no-model-calls --model-steps none"
$B > $o/r02_bench_short.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/r02_launches_bench_steps5.csv $B > $o/r02_ncu_launches.log 2>&1
cap() {  # case, kernel regex, output tag
  python tools/ncu_target.py $1 > $o/r02_target_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o /tmp/r02_$3 python tools/ncu_target.py $1 >> $o/r02_ncu_full.log 2>&1
  # summarise on the box: the reports (36 MB each) do not fit the 64 MiB that travels back
  python tools/ncu_summary.py /tmp/r02_$3.ncu-rep 30 > $o/r02_$3.txt 2>> $o/r02_ncu_full.log
}
cap headline micn_fwd_flat flat_fwd_bf16_1x48x96
cap headline micn_bwd_flat flat_bwd_bf16_1x48x96
cap res48 micn_fwd_res res_fwd_bf16_1x48x48
cap res48 micn_bwd_res res_bwd_bf16_1x48x48
cap dual micn_fwd_flat dual_fwd_bf16_1x48x96
cap dual micn_bwd_flat dual_bwd_bf16_1x48x96
cap fp32_128 micn_bwd_flat flat_bwd_fp32_4x96x128
cp /tmp/r02_flat_bwd_bf16_1x48x96.ncu-rep $o/ 2>/dev/null  # (one report kept whole for source-level reading)
ls -la /tmp/r02_*.ncu-rep $o/r02_*.txt; tail -3 $o/r02_ncu_full.log

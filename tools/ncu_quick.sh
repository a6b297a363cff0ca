#!/bin/bash
# DRAM / L2 counters of the flat kernels at the headline shape (fwd launches first, then bwd)
tag=${1:-q}
ONE="tools/micn_selftest --suite one --N 1 --C 48 --S 96 --dtype bf16 --iters 20"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_bytes.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_bytes.sum,smsp__inst_executed.avg.per_cycle_active
$ONE > gpurun_out/${tag}_one.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:flat -s 4 -c 2 --csv --log-file gpurun_out/${tag}_fwd.csv $ONE > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:bwd_flat -s 4 -c 2 --csv --log-file gpurun_out/${tag}_bwd.csv $ONE > /dev/null 2>&1
cat gpurun_out/${tag}_one.log | tail -1

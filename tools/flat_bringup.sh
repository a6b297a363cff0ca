tag=${1:-v6}
timeout 600 tools/micn_selftest --suite correctness > gpurun_out/${tag}_selftest.log 2>&1; echo "selftest exit $?" >> gpurun_out/${tag}_selftest.log; tail -3 gpurun_out/${tag}_selftest.log
LD_LIBRARY_PATH=tools/trace tools/micn_selftest --suite trace --N 1 --C 48 --S 96 --dtype bf16 > gpurun_out/${tag}_trace.log 2>&1
LD_LIBRARY_PATH=tools/trace tools/micn_selftest --suite trace --N 1 --C 48 --S 96 --dtype bf16 --opt flat_trace_which=2 > gpurun_out/${tag}_trace_bwd.log 2>&1
bash tools/flat_sweep.sh gpurun_out/${tag}_sweep.jsonl > /dev/null 2>&1
grep -c . gpurun_out/${tag}_sweep.jsonl

#!/bin/bash
# call u: evidence for the EAS tensor-core kernel + full GPU suite on the current tree
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -rs --durations=8 > gpurun_out/r02u_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_tests.log )
tail -4 gpurun_out/r02u_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02u_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r02u_bench_hex8.json 2> gpurun_out/r02u_bench_hex8.err; echo "hex8 rc=$?"
for w in hex8_eas_mild hex8_eas_full; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02u_bench_$w.json 2> gpurun_out/r02u_bench_$w.err; echo "$w rc=$?"
  timeout 600 python bench.py --workload $w --option eas_mma=0 --no-cpu > gpurun_out/r02u_bench_${w}_fma.json 2> gpurun_out/r02u_bench_${w}_fma.err; echo "$w fma rc=$?"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
for w in hex8_eas_mild hex8_eas_full; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:eval_assemble -s 3 -c 3 --csv --log-file gpurun_out/r02u_ncu_metrics_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02u_ncu_$w.log 2>&1
done
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_cases.py eas_mild eas_full > gpurun_out/r02u_sanitizer_racecheck_eas.log 2>&1; echo "racecheck rc=$?"; tail -n 2 gpurun_out/r02u_sanitizer_racecheck_eas.log
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_cases.py eas_mild eas_full > gpurun_out/r02u_sanitizer_synccheck_eas.log 2>&1; echo "synccheck rc=$?"; tail -n 2 gpurun_out/r02u_sanitizer_synccheck_eas.log

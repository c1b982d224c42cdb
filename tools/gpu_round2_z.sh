#!/bin/bash
# call z (1 GPU): evidence of the final tree -- bench lines of every workload, ncu launch list + metrics of the headline
# kernel, sanitizer passes over the kernels added late in the round, reference arm
set -u
mkdir -p gpurun_out
for w in hex8 tet10 hex27 hex8_fbar hex8_eas_mild hex8_eas_full hex8_dp; do
  timeout 900 python bench.py --workload $w > gpurun_out/r02z_bench_$w.json 2> gpurun_out/r02z_bench_$w.err; echo "$w rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_ref.json 2>&1; echo "ref rc=$?"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread
for w in hex8 tet10 hex27; do
  timeout 900 ncu --metrics $M --clock-control none -k regex:eval_assemble -s 3 -c 3 --csv --log-file gpurun_out/r02z_ncu_metrics_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02z_ncu_$w.log 2>&1; echo "ncu $w rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_hex8.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02z_launches_hex8.log 2>&1; echo "launches rc=$?"
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py fbar_mulf plastic linear hex27 > gpurun_out/r02z_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -n 2 gpurun_out/r02z_sanitizer_$tool.log
done

#!/bin/bash
# GPU call D (1 GPU): full GPU suite (visco tests), TMA A/B of the hex8 map staging, tet10 / hex27 / F-bar / EAS bench
# lines, ncu metrics of the tet10 and hex27 kernels.
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=8 > gpurun_out/r02d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_tests.log )
tail -4 gpurun_out/r02d_tests.log
for rep in 1 2; do
  timeout 300 python bench.py --no-cpu --no-parity --steps 30 > gpurun_out/r02d_bench_hex8_cpasync_$rep.json 2>/dev/null
  timeout 300 python bench.py --no-cpu --no-parity --steps 30 --option tma_gather=1 > gpurun_out/r02d_bench_hex8_tma_$rep.json 2>/dev/null
done
for w in tet10 hex27 hex8_fbar hex8_eas_mild hex8_eas_full; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err; echo "$w rc=$?"
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active
for w in tet10 hex27; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:eval_assemble -s 3 -c 3 --csv --log-file gpurun_out/r02d_ncu_metrics_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02d_ncu_$w.log 2>&1
done
timeout 600 ncu --metrics $M --clock-control none -k regex:eval_assemble -s 3 -c 3 --csv --log-file gpurun_out/r02d_ncu_metrics_hex8_tma.csv python bench.py --option tma_gather=1 --steps 3 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02d_ncu_hex8_tma.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_cases.py tma > gpurun_out/r02d_sanitizer_racecheck_tma.log 2>&1
echo done

#!/bin/bash
# GPU call A (1 GPU): full GPU test suite, headline bench, ncu launch list + metrics of the current kernel,
# sanitizer passes.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02a_gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=15 > gpurun_out/r02a_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_tests.log ) 
tail -5 gpurun_out/r02a_tests.log
timeout 600 python bench.py > gpurun_out/r02a_bench_hex8.json 2> gpurun_out/r02a_bench_hex8.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_bench_ref.json 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread
timeout 600 ncu --metrics $M --clock-control none -k regex:eval_assemble -s 3 -c 4 --csv --log-file gpurun_out/r02a_ncu_metrics_hex8.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02a_ncu_hex8.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_hex8.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --e2e-steps 3 > gpurun_out/r02a_launches_hex8.log 2>&1
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/r02a_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
done
tail -3 gpurun_out/r02a_sanitizer_*.log

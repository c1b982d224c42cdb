#!/bin/bash
# GPU call G (1 GPU): full GPU suite on the current tree, tet10 bench after the revert, host adapter check.
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=8 > gpurun_out/r02g_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_tests.log )
tail -6 gpurun_out/r02g_tests.log
timeout 600 python bench.py --workload tet10 --steps 10 > gpurun_out/r02g_bench_tet10.json 2> gpurun_out/r02g_bench_tet10.err; echo "tet10 rc=$?"
timeout 600 python bench.py --workload hex8_dp --no-cpu --steps 10 > gpurun_out/r02g_bench_hex8_dp.json 2> gpurun_out/r02g_bench_hex8_dp.err; echo "dp rc=$?"
./4c_b200/host/host_check > gpurun_out/r02g_host_check.log 2>&1; echo "host_check rc=$?"; cat gpurun_out/r02g_host_check.log | tail -4
echo done

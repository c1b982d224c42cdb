#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/order_probe.py 24 > gpurun_out/r02f_order_probe.log 2>&1; cat gpurun_out/r02f_order_probe.log | tail -8
timeout 600 python bench.py --workload tet10 --no-cpu --no-parity --steps 10 --option segments=0 > gpurun_out/r02f_bench_tet10_noseg.json 2> gpurun_out/r02f_bench_tet10_noseg.err; echo "noseg rc=$?"
( timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -k "symmetric or pipelined" > gpurun_out/r02f_tests_sym.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_tests_sym.log ); tail -15 gpurun_out/r02f_tests_sym.log
for st in 0 1; do
  timeout 300 python bench.py --no-cpu --no-parity --steps 10 --e2e-steps 8 --option symmetric_transfer=$st > gpurun_out/r02f_bench_hex8_sym$st.json 2> gpurun_out/r02f_bench_hex8_sym$st.err; echo "sym$st rc=$?"
done
timeout 300 python bench.py --no-cpu --no-parity --steps 10 --e2e-steps 8 --option symmetric_transfer=1 --option nt_stores=0 > gpurun_out/r02f_bench_hex8_sym1_nont.json 2>/dev/null
timeout 300 python bench.py --no-cpu --no-parity --steps 10 --e2e-steps 8 --option symmetric_transfer=1 --option host_threads=8 > gpurun_out/r02f_bench_hex8_sym1_t8.json 2>/dev/null
echo done

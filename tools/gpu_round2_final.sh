#!/bin/bash
# final tree of round 2 (1 GPU): the whole GPU suite, smoke(), the default bench line and the reference arm, as the driver runs them
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02f_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r02f_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02f_smoke.log
timeout 900 python bench.py > gpurun_out/r02f_bench_hex8.json 2> gpurun_out/r02f_bench_hex8.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02f_bench_hex8.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2>&1; echo "ref rc=$?"

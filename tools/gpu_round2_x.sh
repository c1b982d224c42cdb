#!/bin/bash
# call x: full GPU suite, smoke and the headline bench on the current tree
set -u
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -rs --durations=8 > gpurun_out/r02x_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_tests.log )
tail -4 gpurun_out/r02x_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02x_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r02x_bench_hex8.json 2> gpurun_out/r02x_bench_hex8.err; echo "hex8 rc=$?"
tail -c 1500 gpurun_out/r02x_bench_hex8.json

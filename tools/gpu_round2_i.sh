#!/bin/bash
# call i: full GPU suite on the tree with tet4 / hex20 + smoke + host adapter check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02i_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02i_tests.log 2>&1
echo "pytest rc=$?"
tail -n 4 gpurun_out/r02i_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02i_smoke.log 2>&1
echo "smoke rc=$?"
tail -n 3 gpurun_out/r02i_smoke.log

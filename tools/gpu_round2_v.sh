#!/bin/bash
# call v: GPU tests of the general cells (tet4 / hex20 / wedge6 / pyramid5) + host adapter check
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "tet4 or hex20 or wedge6 or pyramid5" > gpurun_out/r02v_tests.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r02v_tests.log
timeout 300 python -m pytest tests/test_host_adapter.py -m gpu -q > gpurun_out/r02v_host.log 2>&1; echo "host rc=$?"; tail -n 3 gpurun_out/r02v_host.log

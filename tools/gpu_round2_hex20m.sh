# hex20 MULF through the hex27 kernel: the cell's MULF tests (oracle, history, known answers), then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "mulf" > gpurun_out/r02j_mulf.log 2>&1; tail -3 gpurun_out/r02j_mulf.log
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02j_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r02j_gpu_suite.log

# edge cases (empty blocks, ranks without row nodes, isolated nodes) and the whole GPU suite on the same tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q > gpurun_out/r02g_edge.log 2>&1; tail -30 gpurun_out/r02g_edge.log
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02g_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r02g_gpu_suite.log

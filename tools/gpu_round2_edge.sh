# edge cases (empty blocks, ranks without row nodes, isolated nodes, caller graphs with more / fewer entries)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q > gpurun_out/r02g_edge.log 2>&1; tail -30 gpurun_out/r02g_edge.log

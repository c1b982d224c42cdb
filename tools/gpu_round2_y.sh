#!/bin/bash
# call y (2 GPUs): NCCL row-export test + A/B of owner-computes vs row export on the weak-scaling mesh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -k "world2 or 2-" -rs > gpurun_out/r02y_mgpu2.log 2>&1; echo "mgpu rc=$?"; tail -4 gpurun_out/r02y_mgpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/export_bench.py 100 20 > gpurun_out/r02y_export_bench_n2.log 2>&1; echo "bench rc=$?"
grep "^{" gpurun_out/r02y_export_bench_n2.log; tail -3 gpurun_out/r02y_export_bench_n2.log

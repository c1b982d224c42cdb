#!/bin/bash
# final tree of round 2 on 2 GPUs: the NCCL tests the 1-GPU suite skips, and the bench line at N = 2 as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -m gpu -q -k "multirank or world2 or nccl" -rs > gpurun_out/r02f_mgpu2.log 2>&1; echo "mgpu rc=$?"; tail -6 gpurun_out/r02f_mgpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02f_bench_hex8_n2.json 2> gpurun_out/r02f_bench_hex8_n2.err; echo "bench rc=$?"
grep "^{" gpurun_out/r02f_bench_hex8_n2.json | cut -c1-300

#!/bin/bash
# GPU call C (8 GPUs): 8-rank NCCL tests, weak-scaling bench at 8 (with the config-5 sub-run) and 4, D2H probe.
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02c_topo.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -rs > gpurun_out/r02c_tests_n8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_tests_n8.log )
tail -4 gpurun_out/r02c_tests_n8.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02c_bench_hex8_n8.json 2> gpurun_out/r02c_bench_hex8_n8.err; echo "bench8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02c_bench_hex8_n4.json 2> gpurun_out/r02c_bench_hex8_n4.err; echo "bench4 rc=$?"
timeout 600 python tools/d2h_probe.py --gb 2 --reps 3 > gpurun_out/r02c_d2h_probe_n8.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/r02c_d2h_probe_n8.log | cut -c1-300

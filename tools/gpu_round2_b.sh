#!/bin/bash
# GPU call B (N GPUs): multi-rank NCCL tests (halo overlap path), weak-scaling bench at N, D2H probe.
set -u
N=${1:-2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -rs > gpurun_out/r02b_tests_n$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests_n$N.log )
tail -4 gpurun_out/r02b_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02b_bench_hex8_n$N.json 2> gpurun_out/r02b_bench_hex8_n$N.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --option overlap_halo=0 --no-parity --no-config5 --host-alloc plain > gpurun_out/r02b_bench_hex8_n${N}_nooverlap_plain.json 2> gpurun_out/r02b_bench_hex8_n${N}_nooverlap_plain.err; echo "bench2 rc=$?"
timeout 900 python tools/d2h_probe.py --gb 2 --reps 4 > gpurun_out/r02b_d2h_probe_n$N.log 2>&1; echo "probe rc=$?"
tail -5 gpurun_out/r02b_d2h_probe_n$N.log

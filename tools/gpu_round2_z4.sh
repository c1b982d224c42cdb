#!/bin/bash
# call z4 (4 GPUs): NCCL tests of the final tree (halo, actions, row export at world 2 and 4) + weak-scaling bench at N = 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -q -rs > gpurun_out/r02z_mgpu4.log 2>&1; echo "mgpu rc=$?"; tail -5 gpurun_out/r02z_mgpu4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02z_bench_hex8_n4.json 2> gpurun_out/r02z_bench_hex8_n4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench_hex8_n4.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["parity"]["ok"], d["e2e"]["value"])
PY

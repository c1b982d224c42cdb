#!/bin/bash
# call o: EAS tensor-core kernel, warps-per-SM A/B (tools/_libs/w65.so: mild 6x2, full 5x2 warps)
mkdir -p gpurun_out
for w in hex8_eas_mild hex8_eas_full; do
  B200S_LIB=$PWD/tools/_libs/w65.so timeout 600 python bench.py --workload $w --no-cpu --no-parity --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r02o_bench_${w}_w65.json 2> gpurun_out/r02o_bench_${w}_w65.err
  echo "bench $w w65 rc=$?"
done

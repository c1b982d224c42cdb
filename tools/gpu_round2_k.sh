#!/bin/bash
# call s: ncu --set full of the EAS tensor-core kernels (mild, full) + the new A/B test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "eas_tensor or tet4" > gpurun_out/r02s_tests.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r02s_tests.log
for w in mild full; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:eas_mma -s 3 -c 1 -f -o gpurun_out/r02s_eas_${w} \
    python bench.py --workload hex8_eas_$w --no-cpu --no-parity --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/r02s_ncu_${w}.log 2>&1
  echo "ncu $w rc=$?"
done
ls -la gpurun_out/*.ncu-rep

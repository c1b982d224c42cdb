#!/bin/bash
# call t: EAS tensor-core kernel -- parity tests, memcheck, A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_post.py -m gpu -q -x -k "eas or tet4" > gpurun_out/r02t_tests.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r02t_tests.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02t_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 3 gpurun_out/r02t_sanitizer_memcheck.log
for w in hex8_eas_mild hex8_eas_full; do
  for m in 1; do
    timeout 600 python bench.py --workload $w --option eas_mma=$m --no-cpu --steps 10 --warmup 3 > gpurun_out/r02t_bench_${w}_mma$m.json 2> gpurun_out/r02t_bench_${w}_mma$m.err
    echo "bench $w mma=$m rc=$?"; tail -c 600 gpurun_out/r02t_bench_${w}_mma$m.json
  done
done

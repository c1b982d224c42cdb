# call h3: TECH shell_eas / shell_eas_ans through the CUDA path (parity + the reference's known answers), racecheck / synccheck of
# the general kernel's two-step EAS condensation
mkdir -p gpurun_out; L=gpurun_out/r02z5_tests.log
python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "shell" > $L 2>&1; tail -5 $L
for t in racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $t python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "shell_eas_ans" > gpurun_out/r02z5_sanitizer_$t.log 2>&1
  tail -4 gpurun_out/r02z5_sanitizer_$t.log
done

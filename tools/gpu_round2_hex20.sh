# hex20 through the hex27 kernel: parity tests of the cell, rate against the general kernel (option general_kernel=1), hex27 unchanged
mkdir -p gpurun_out; L=gpurun_out/r02h_hex20.log; : > $L
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_edge_cases.py -m gpu -q -x -k "hex20 or edge" > gpurun_out/r02h_hex20_tests.log 2>&1; tail -15 gpurun_out/r02h_hex20_tests.log
LIBPROBE_OPTS="general_kernel=1" timeout 300 python tools/lib_probe.py hex20 40 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex20 40 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex20 64 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex27 80 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
cat $L

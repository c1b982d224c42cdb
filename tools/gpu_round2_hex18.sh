# hex18 through the hex27 kernel (nodes and Gauss points padded): rate vs the general kernel, then the whole GPU suite
mkdir -p gpurun_out; L=gpurun_out/r02i_hex18.log; : > $L
LIBPROBE_OPTS="general_kernel=1" timeout 300 python tools/lib_probe.py hex18 40 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex18 40 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex18 64 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex20 64 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
timeout 300 python tools/lib_probe.py hex27 80 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
cat $L
timeout 1500 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r02i_gpu_suite.log 2>&1; echo "suite rc=$?"; grep -a "THROUGHPUT" gpurun_out/r02i_gpu_suite.log | tail -3; tail -3 gpurun_out/r02i_gpu_suite.log

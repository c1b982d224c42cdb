mkdir -p gpurun_out
python tools/lib_probe.py tet10 87 tools/_libs/base.so tools/_libs/main.so tools/_libs/h1.so 2>&1 | grep LIBPROBE > gpurun_out/r02w_probe1.log
python tools/lib_probe.py hex8 100 tools/_libs/base.so tools/_libs/main.so tools/_libs/h1.so 2>&1 | grep LIBPROBE >> gpurun_out/r02w_probe1.log
python tools/lib_probe.py hex27 80 tools/_libs/base.so tools/_libs/main.so 2>&1 | grep LIBPROBE >> gpurun_out/r02w_probe1.log
cat gpurun_out/r02w_probe1.log

# A/B of the hex27 kernel: build without any hex20 padding code (tools/_libs/nopad.so: the library of commit 48090ef or earlier, built with
# tools/build_variant.py there and copied here; not kept in the tree) vs the product (padding as a template flag),
# same box, alternating; hex20 rate; then the whole GPU suite on the tree
mkdir -p gpurun_out; L=gpurun_out/r02h_hex27_ab2.log; : > $L
for i in 1 2; do
  timeout 300 python tools/lib_probe.py hex27 80 tools/_libs/nopad.so 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
done
timeout 300 python tools/lib_probe.py hex20 64 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
cat $L
timeout 1500 python -m pytest tests/ -x -q -m gpu -rA > gpurun_out/r02h_gpu_suite.log 2>&1; echo "suite rc=$?"; grep -a "THROUGHPUT" gpurun_out/r02h_gpu_suite.log; tail -3 gpurun_out/r02h_gpu_suite.log

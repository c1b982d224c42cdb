# call h: element order A/B -- Z-order (reorder=1, default) vs Hilbert curve (reorder=2)
mkdir -p gpurun_out; L=gpurun_out/r02z_order_probe.log; : > $L
for c in "tet10 87" "hex8 100" "hex27 80"; do for r in 1 2; do
  LIBPROBE_OPTS="reorder=$r" python tools/lib_probe.py $c 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
done; done
cat $L

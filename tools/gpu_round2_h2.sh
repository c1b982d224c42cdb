# call h2: row-phase segments vs plain curve order for the device-resident kernel (tet10 / hex27 / hex8)
mkdir -p gpurun_out; L=gpurun_out/r02z_order_probe2.log; : > $L
for c in "tet10 87" "hex27 80" "hex8 100"; do for o in "pipeline_phases=12" "segments=0" "segments=0 reorder=2" "pipeline_phases=48"; do
  LIBPROBE_OPTS="$o" python tools/lib_probe.py $c 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
done; done
cat $L

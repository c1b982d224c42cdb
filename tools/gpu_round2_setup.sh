# set-up cost (once per discretization: set_mesh + set_material + build_graph incl. the scatter map) of the BASELINE meshes
mkdir -p gpurun_out; L=gpurun_out/r02h_setup.log; : > $L
for c in "hex8 100" "hex8 100" "tet10 87" "hex27 80" "hex20 64"; do
  timeout 600 python tools/lib_probe.py $c 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
done
cat $L

# kernel rates of the table-driven general kernel (cells without a specialised kernel), for the record
mkdir -p gpurun_out; L=gpurun_out/r02g_general_cells.log; : > $L
for c in "tet4 60" "hex20 40" "wedge6 50" "pyramid5 40" "hex18 40"; do
  timeout 300 python tools/lib_probe.py $c 4c_b200/lib4c_b200_solid.so 2>&1 | grep LIBPROBE >> $L
done
cat $L

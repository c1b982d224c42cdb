# ncu metrics of the hex27 kernel in its hex20 (PAD) instantiation, next to the native hex27 one
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sectors_op_red.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
for c in "hex20 40" "hex27 40"; do
  set -- $c
  timeout 600 ncu --metrics $M --clock-control none -k regex:hex27_eval_assemble -s 4 -c 2 --csv --log-file gpurun_out/r02h_ncu_metrics_padded_$1.csv python tools/lib_probe.py $1 $2 4c_b200/lib4c_b200_solid.so > gpurun_out/r02h_ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
done
grep -a "gpu__time_duration\|dram__bytes\|lts__t_sectors_op_red\|fp64\|registers" gpurun_out/r02h_ncu_metrics_padded_hex20.csv | cut -d, -f5,10-15 | head -20

# call w2: tet10 math-only vs full (scatter_mode), evict-last RED variant, EAS scatter A/B through bench.py
mkdir -p gpurun_out
L=gpurun_out/r02w_probe2.log; : > $L
LIBPROBE_OPTS="scatter_mode=1" python tools/lib_probe.py tet10 87 tools/_libs/base.so 2>&1 | grep LIBPROBE >> $L
python tools/lib_probe.py tet10 87 tools/_libs/base.so tools/_libs/main.so tools/_libs/h3.so 2>&1 | grep LIBPROBE >> $L
LIBPROBE_OPTS="scatter_mode=1" python tools/lib_probe.py hex8 100 tools/_libs/base.so 2>&1 | grep LIBPROBE >> $L
python tools/lib_probe.py hex8 100 tools/_libs/base.so tools/_libs/h3.so 2>&1 | grep LIBPROBE >> $L
for lib in base main; do for w in hex8_eas_mild hex8_eas_full; do
  B200S_LIB=$PWD/tools/_libs/$lib.so python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-parity --e2e-steps 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('BENCH $lib $w', d['value'], d['ms_per_step'], d.get('roofline',{}).get('kernel_ms'))" >> $L
done; done
cat $L

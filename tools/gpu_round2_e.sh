#!/bin/bash
# GPU call E (1 GPU): tet10 regression hunt (round-1 library vs current, reorder on/off), copy-engine zeroing A/B.
set -u
mkdir -p gpurun_out
( cd tools/_r01tree && timeout 600 python bench.py --workload tet10 --no-cpu --steps 10 > ../../gpurun_out/r02e_bench_tet10_r01lib.json 2> ../../gpurun_out/r02e_bench_tet10_r01lib.err ); echo "r01lib rc=$?"
timeout 600 python bench.py --workload tet10 --no-cpu --no-parity --steps 10 > gpurun_out/r02e_bench_tet10_cur.json 2> gpurun_out/r02e_bench_tet10_cur.err; echo "cur rc=$?"
timeout 600 python bench.py --workload tet10 --no-cpu --no-parity --steps 10 --option pipeline_phases=1 > gpurun_out/r02e_bench_tet10_p1.json 2> gpurun_out/r02e_bench_tet10_p1.err; echo "p1 rc=$?"
timeout 600 python bench.py --workload tet10 --no-cpu --no-parity --steps 10 --option reorder=0 > gpurun_out/r02e_bench_tet10_noreorder.json 2> gpurun_out/r02e_bench_tet10_noreorder.err; echo "noreorder rc=$?"
for db in 0 1 2; do
  timeout 300 python bench.py --no-cpu --no-parity --steps 30 --option double_buffer=$db > gpurun_out/r02e_bench_hex8_db$db.json 2> gpurun_out/r02e_bench_hex8_db$db.err; echo "db$db rc=$?"
done
timeout 300 python bench.py --no-cpu --no-parity --steps 30 --option double_buffer=2 > gpurun_out/r02e_bench_hex8_db2_b.json 2>/dev/null
echo done

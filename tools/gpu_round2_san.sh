# sanitizer passes over the hex20 / hex18 route through the hex27 kernel (plain shared-memory stores of the padding row records next
# to the cp.async gather) and over the empty-block launches
mkdir -p gpurun_out
for t in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $t python -m pytest tests/test_gpu_round2.py tests/test_gpu_edge_cases.py -m gpu -q -x -k "fast_path or empty_element_block or no_nodes or ghost_elements_only" > gpurun_out/r02j_sanitizer_$t.log 2>&1
  echo "$t rc=$?"; tail -3 gpurun_out/r02j_sanitizer_$t.log
done

"""red.global.add.f64 throughput vs run length (coalescing) inside an L2-resident window."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
capi = importlib.import_module("4c_b200").capi
out = {}
for R in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
    out[f"run{R}"] = capi.microbench(100 + R, 2 << 30)
print(json.dumps({"red_f64_Gatomics_per_s_by_run_length": out}))

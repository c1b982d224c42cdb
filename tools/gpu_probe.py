"""Development probe run on the GPU box: microbenchmarks (FP64 DFMA / DMMA peaks, memset, red.f64
patterns) and timing of the hex8 kernel under the profiling options.  Prints JSON lines; results are
summarised in DESIGN.md / profiles/."""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

b200 = importlib.import_module("4c_b200")
capi = b200.capi


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    if "--micro" in sys.argv:
        out = {}
        out["dfma_tflops"], out["implied_mhz"] = capi.measure_fp64_peak(0)
        out["dmma_tflops"] = capi.microbench(1)
        out["dmma_dfma_tflops"] = capi.microbench(2)
        out["memset_gbs"] = capi.microbench(3, 2 << 30)
        for w, name in ((4, "red_contig"), (5, "red_runs3_l2win"), (6, "red_random")):
            out[name + "_Gps"] = capi.microbench(w, 2 << 30)
        print(json.dumps({"microbench": out}), flush=True)

    rm = b200.partition.box_partition(b200.mesh.HEX8, (n, n, n), 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / n)
    import ctypes
    cudart = ctypes.CDLL("libcudart.so")
    uh = np.ascontiguousarray(u.ravel())
    for reorder in (1, 0):
        ctx = capi.Context(0)
        ctx.set_option("reorder", reorder)
        t1 = time.time()
        ctx.set_mesh(capi.HEX8, rm.X, rm.conn)
        ctx.set_material(capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
        nnz = ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
        t2 = time.time()
        print(json.dumps({"setup": {"reorder": reorder, "device_setup_s": t2 - t1, "nnz": nnz,
                                    "n_ele": rm.n_col_eles}}), flush=True)
        ptrs = ctx.device_ptrs()
        cudart.cudaMemcpy(ctypes.c_void_p(ptrs["u_row"]), ctypes.c_void_p(uh.ctypes.data),
                          ctypes.c_size_t(uh.nbytes), 1)
        for mk, par, mname in ((capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], "neohooke"),
                               (capi.MAT_SVK, [1000.0, 0.3], "svk"), (capi.MAT_ISO_VOL, [1.0, 1000.0], "isovol")):
            ctx.set_material(mk, par)
            for ctas in (1, 2):
                for mode in (0, 1):
                    ctx.set_option("ctas_per_sm", ctas)
                    ctx.set_option("scatter_mode", mode)
                    ts = []
                    for it in range(6):
                        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
                        ctx.synchronize()
                        ts.append(ctx.timers())
                    k = np.median([t["kernel"] for t in ts[2:]])
                    d = np.median([t["device"] for t in ts[2:]])
                    print(json.dumps({"reorder": reorder, "mat": mname, "ctas_per_sm": ctas, "scatter_mode": mode,
                                      "kernel_ms": round(float(k), 4), "device_ms": round(float(d), 4),
                                      "Melem_per_s_kernel": round(rm.n_col_eles / k / 1e3, 1)}), flush=True)
            ctx.set_option("scatter_mode", 0)
            ctx.set_option("ctas_per_sm", 2)
            ts = []
            for it in range(5):
                ctx.evaluate_async(capi.ACTION_INTERNALFORCE)
                ctx.synchronize()
                ts.append(ctx.timers()["kernel"])
            print(json.dumps({"reorder": reorder, "mat": mname, "action": "internalforce",
                              "kernel_ms": float(np.median(ts[2:]))}), flush=True)
        ctx.close()


if __name__ == "__main__":
    main()

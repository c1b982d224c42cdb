"""Development probe run on the GPU box: microbenchmarks (FP64 DFMA / DMMA peaks, memset, red.f64
patterns) and timing of the hex8 kernel under the profiling options.  Prints JSON lines; results are
summarised in DESIGN.md / profiles/."""
import importlib
import json
import sys
import time

import numpy as np

b200 = importlib.import_module("4c_b200")
capi = b200.capi


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    out = {}
    out["dfma_tflops"], out["implied_mhz"] = capi.measure_fp64_peak(0)
    out["dmma_tflops"] = capi.microbench(1)
    out["dmma_dfma_tflops"] = capi.microbench(2)
    out["memset_gbs"] = capi.microbench(3, 2 << 30)
    for w, name in ((4, "red_contig"), (5, "red_runs3_l2win"), (6, "red_random")):
        out[name + "_Gps"] = capi.microbench(w, 2 << 30)
    print(json.dumps({"microbench": out}), flush=True)

    t0 = time.time()
    rm = b200.partition.box_partition(b200.mesh.HEX8, (n, n, n), 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / n)
    t1 = time.time()
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    t2 = time.time()
    print(json.dumps({"setup": {"mesh_s": t1 - t0, "device_setup_s": t2 - t1, "nnz": asm.nnz,
                                "n_ele": rm.n_col_eles}}), flush=True)
    import ctypes
    import torch
    ptrs = asm.ctx.device_ptrs()
    cudart = ctypes.CDLL("libcudart.so")
    uh = np.ascontiguousarray(u.ravel())
    cudart.cudaMemcpy(ctypes.c_void_p(ptrs["u_row"]), ctypes.c_void_p(uh.ctypes.data),
                      ctypes.c_size_t(uh.nbytes), 1)
    for mk, par, mname in ((capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], "neohooke"),
                           (capi.MAT_SVK, [1000.0, 0.3], "svk"), (capi.MAT_ISO_VOL, [1.0, 1000.0], "isovol")):
        asm.ctx.set_material(mk, par)
        for ctas in (1, 2):
            for mode in (0, 1, 2):
                asm.ctx.set_option("ctas_per_sm", ctas)
                asm.ctx.set_option("scatter_mode", mode)
                ts = []
                for it in range(6):
                    asm.ctx.evaluate_async(capi.ACTION_NLNSTIFF)
                    asm.ctx.synchronize()
                    ts.append(asm.ctx.timers())
                k = np.median([t["kernel"] for t in ts[2:]])
                z = np.median([t["zero"] for t in ts[2:]])
                d = np.median([t["device"] for t in ts[2:]])
                print(json.dumps({"mat": mname, "ctas_per_sm": ctas, "scatter_mode": mode, "kernel_ms": k,
                                  "zero_ms": z, "device_ms": d, "Melem_per_s_kernel": rm.n_col_eles / k / 1e3,
                                  "Melem_per_s_eval": rm.n_col_eles / d / 1e3}), flush=True)
        asm.ctx.set_option("scatter_mode", 0)
        asm.ctx.set_option("ctas_per_sm", 2)
        ts = []
        for it in range(5):
            asm.ctx.evaluate_async(capi.ACTION_INTERNALFORCE)
            asm.ctx.synchronize()
            ts.append(asm.ctx.timers()["kernel"])
        print(json.dumps({"mat": mname, "action": "internalforce", "kernel_ms": float(np.median(ts[2:]))}),
              flush=True)


if __name__ == "__main__":
    main()

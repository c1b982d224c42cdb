"""Profiling helper (not part of the product path): element-kernel time of a bench workload with the
scatter switched off (scatter_mode 1) and with the locality ordering off, to see the cost split."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

b200 = importlib.import_module("4c_b200")
capi = b200.capi

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="tet10")
ap.add_argument("--n", type=int, default=0)
args = ap.parse_args()
wl = bench.WORKLOADS[args.workload]
n = args.n or wl["default_n"]
rm, u, label = bench.build_rank(wl, n, 0, 1)
for opts in ({}, {"scatter_mode": 1}, {"reorder": 0}, {"ctas_per_sm": 1}):
    ctx = capi.Context(0)
    for k, v in opts.items():
        if k != "scatter_mode":
            ctx.set_option(k, v)
    ctx.set_mesh(rm.celltype, rm.X, rm.conn)
    ctx.set_material(getattr(capi, wl["mat"]), wl["par"])
    ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
    if "scatter_mode" in opts:
        ctx.set_option("scatter_mode", opts["scatter_mode"])
    ur = np.ascontiguousarray(u[:rm.n_row_nodes].ravel())
    ctx.evaluate(capi.ACTION_NLNSTIFF, ur, None, None)
    ctx.kernel_time(reset=True)
    for _ in range(5):
        ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    ctx.synchronize(check=False)
    s, c = ctx.kernel_time(reset=True)
    print(f"{args.workload} {label} options {opts}: element kernel {s / c:.3f} ms", flush=True)
    ctx.close()

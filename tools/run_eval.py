"""Small driver for ncu captures: set up an n^3 hex8 block and run a few evaluates.
    python tools/run_eval.py [n] [--scatter-mode M] [--ctas C] [--iters K] [--mat 0|1|2] [--no-reorder]"""
import argparse
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("4c_b200")
capi = b200.capi

ap = argparse.ArgumentParser()
ap.add_argument("n", type=int, nargs="?", default=100)
ap.add_argument("--scatter-mode", type=int, default=0)
ap.add_argument("--ctas", type=int, default=2)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--mat", type=int, default=0)
ap.add_argument("--no-reorder", action="store_true")
ap.add_argument("--force-only", action="store_true")
ap.add_argument("--cell", default="hex8", choices=["hex8", "tet10", "hex27"])
a = ap.parse_args()
PAR = {0: [200.0, 0.25], 1: [1.0, 1000.0], 2: [1000.0, 0.3]}
if a.cell == "tet10":
    X, conn = b200.mesh.unstructured_tet10(a.n)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
else:
    rm = b200.partition.box_partition(getattr(b200.mesh, a.cell.upper()), (a.n,) * 3, 0, 1)
u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / a.n)
ctx = capi.Context(0)
ctx.set_option("reorder", 0 if a.no_reorder else 1)
ctx.set_mesh(rm.celltype, rm.X, rm.conn)
ctx.set_material(a.mat, PAR[a.mat])
ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
ctx.set_option("scatter_mode", a.scatter_mode)
ctx.set_option("ctas_per_sm", a.ctas)
ctx.evaluate(capi.ACTION_NLNSTIFF, np.ascontiguousarray(u.ravel()), None, None)
act = capi.ACTION_INTERNALFORCE if a.force_only else capi.ACTION_NLNSTIFF
ts = []
for _ in range(a.iters):
    ctx.evaluate_async(act)
    ctx.synchronize()
    ts.append(ctx.timers()["kernel"])
print("kernel ms:", [round(t, 4) for t in ts])
ctx.close()

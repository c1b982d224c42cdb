"""Development probe: time the element kernel of several builds of lib4c_b200_solid.so on the same mesh, using
only entry points that exist since round 1 (bisecting a kernel-time regression).
    python tools/lib_probe.py tet10 87 tools/_libs/*.so"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("4c_b200")
cell, n = sys.argv[1], int(sys.argv[2])
if cell == "tet10":
    X, conn = b200.mesh.unstructured_tet10(n)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    mat, par, h = 1, [1.0, 1000.0], 1.0 / n
elif cell in ("tet4", "hex20", "wedge6", "pyramid5", "hex18"):  # cells of the table-driven general kernel
    gen = {"tet4": b200.mesh.unstructured_tet4, "hex20": b200.mesh.structured_hex20, "wedge6": b200.mesh.structured_wedge6,
           "pyramid5": b200.mesh.structured_pyramid5, "hex18": b200.mesh.structured_hex18}[cell]
    X, conn = gen(n) if cell == "tet4" else gen((n,) * 3)
    rm = b200.partition.serial_mesh(getattr(b200.mesh, cell.upper()), X, conn)
    mat, par, h = 0, [200.0, 0.25], 1.0 / n
else:
    rm = b200.partition.box_partition(getattr(b200.mesh, cell.upper()), (n,) * 3, 0, 1)
    mat, par, h = 0, [200.0, 0.25], 0.01
u = np.ascontiguousarray(b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h).ravel())
Xc = np.ascontiguousarray(rm.X, np.float64)
cc = np.ascontiguousarray(rm.conn, np.int32)
rl = np.ascontiguousarray(rm.node_row_lid, np.int32)
cl = np.ascontiguousarray(rm.node_col_lid, np.int32)
p = np.ascontiguousarray(par, np.float64)
vp, i64 = C.c_void_p, C.c_int64
for path in sys.argv[3:]:
    L = C.CDLL(os.path.abspath(path))
    L.b200s_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.b200s_set_mesh.argtypes = [vp, C.c_int, i64, vp, i64, vp]
    L.b200s_set_material.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.b200s_build_graph.argtypes = [vp, i64, vp, vp, vp, vp, C.POINTER(i64)]
    L.b200s_evaluate.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.b200s_evaluate_async.argtypes = [vp, C.c_int]
    L.b200s_synchronize.argtypes = [vp, vp]
    L.b200s_kernel_time.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64), C.c_int]
    L.b200s_destroy.argtypes = [vp]
    L.b200s_last_error.argtypes = [vp]
    L.b200s_last_error.restype = C.c_char_p
    hdl = vp()
    assert L.b200s_create(0, 0, 1, C.byref(hdl)) == 0
    L.b200s_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    for kv in os.environ.get("LIBPROBE_OPTS", "").split():  # e.g. LIBPROBE_OPTS="scatter_mode=1"
        k, v = kv.split("=")
        assert L.b200s_set_option(hdl, k.encode(), float(v)) == 0, kv
    import time
    t_setup = time.perf_counter()
    assert L.b200s_set_mesh(hdl, rm.celltype, len(Xc), Xc.ctypes.data, len(cc), cc.ctypes.data) == 0
    assert L.b200s_set_material(hdl, mat, p.ctypes.data, len(p), None) == 0
    nnz = i64(0)
    rc = L.b200s_build_graph(hdl, rm.n_row_nodes, rl.ctypes.data, cl.ctypes.data, None, None, C.byref(nnz))
    assert rc == 0, L.b200s_last_error(hdl)
    t_setup = time.perf_counter() - t_setup  # set_mesh (H2D, locality order) + set_material + build_graph (graph, row phases, scatter map); synchronous calls
    assert L.b200s_evaluate(hdl, 1, u.ctypes.data, None, None, None) == 0
    for _ in range(3):
        L.b200s_evaluate_async(hdl, 1)
    L.b200s_synchronize(hdl, None)
    s, c = C.c_double(0), i64(0)
    L.b200s_kernel_time(hdl, C.byref(s), C.byref(c), 1)
    for _ in range(8):
        L.b200s_evaluate_async(hdl, 1)
    L.b200s_synchronize(hdl, None)
    L.b200s_kernel_time(hdl, C.byref(s), C.byref(c), 1)
    print(f"LIBPROBE {os.path.basename(path):16s} {os.environ.get('LIBPROBE_OPTS', ''):16s} {cell} n={n} nnz={nnz.value} kernel {s.value / max(c.value, 1):.3f} ms over {c.value} evaluates, {len(cc)} elements, {len(cc) / (s.value / max(c.value, 1)) / 1e3:.2f} M elements/s, set-up {t_setup * 1e3:.0f} ms", flush=True)
    L.b200s_destroy(hdl)

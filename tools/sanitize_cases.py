"""Small evaluates of every element kernel family for compute-sanitizer (racecheck / synccheck / memcheck):
the kernels alias shared memory across phases separated by __syncwarp() / named barriers, which is what these
tools check.  Run as  compute-sanitizer --tool racecheck python tools/sanitize_cases.py  on the GPU box."""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("4c_b200")
capi = b200.capi

CASES = [("hex8", "HEX8", (7, 6, 5), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], {}),
         ("hex8 tma", "HEX8", (7, 6, 5), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], {}),   # option tma_gather = 1
         ("hex8 fbar", "HEX8", (6, 5, 4), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(tech=capi.TECH_FBAR)),
         ("hex8 eas_mild", "HEX8", (5, 4, 4), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(tech=capi.TECH_EAS_MILD)),
         ("hex8 eas_full", "HEX8", (5, 4, 3), capi.MAT_SVK, [1000.0, 0.3], dict(tech=capi.TECH_EAS_FULL)),
         ("hex8 mulf", "HEX8", (5, 4, 3), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(prestress=capi.PRESTRESS_MULF)),
         ("hex27", "HEX27", (3, 3, 2), capi.MAT_SVK, [1000.0, 0.3], {}),
         ("tet10", "TET10", 4, capi.MAT_ISO_VOL, [1.0, 1000.0], {}),
         # round 2, later additions: general kernel with F-bar (+ MULF), plastic material on the hex8 / EAS / general kernels,
         # linear-kinematics EAS
         ("general fbar_mulf", "HEX8", (4, 3, 3), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25],
          dict(tech=capi.TECH_FBAR, prestress=capi.PRESTRESS_MULF)),
         ("hex8 plastic", "HEX8", (5, 4, 3), capi.MAT_PLASTIC_NLNLOG, [206.9, 0.29, 0.0, 0.45, 0.13, 0.715, 16.93, 1.0, 0.1, 1e-12], {}),
         ("eas_mild plastic", "HEX8", (4, 3, 3), capi.MAT_PLASTIC_NLNLOG, [206.9, 0.29, 0.0, 0.45, 0.13, 0.715, 16.93, 1.0, 0.1, 1e-12],
          dict(tech=capi.TECH_EAS_MILD)),
         ("hex27 plastic", "HEX27", (2, 2, 2), capi.MAT_PLASTIC_NLNLOG, [206.9, 0.29, 0.0, 0.45, 0.13, 0.715, 16.93, 1.0, 0.1, 1e-12], {}),
         ("eas_full linear", "HEX8", (4, 3, 3), capi.MAT_SVK, [1000.0, 0.3], dict(tech=capi.TECH_EAS_FULL, kinem=capi.KINEM_LINEAR))]
only = sys.argv[1:] or None
for name, cell, n, mk, par, kw in CASES:
    if only and name.split()[-1] not in only and name not in only:
        continue
    if cell == "TET10":
        X, conn = b200.mesh.unstructured_tet10(n)
        rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
        h = 1.0 / n
    else:
        rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
        h = 1.0 / max(n)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    asm = b200.DeviceAssembler(rm, mk, par, **kw)
    if mk == capi.MAT_PLASTIC_NLNLOG:
        asm.ctx.set_time_step_size(0.175)
    if name.endswith("tma"):
        asm.ctx.set_option("tma_gather", 1)
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    print(f"SANITIZE_CASE {name}: {rm.n_col_eles} elements, |K|max {np.abs(K).max():.4g}, |f|max {np.abs(f).max():.4g}, "
          f"force-only diff {np.abs(f - f2).max():.2g}", flush=True)
    asm.close()
print("SANITIZE_DONE", flush=True)

"""Development probe: device element order with and without the segment ordering (tet10, random numbering)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("4c_b200")
capi = b200.capi
X, conn = b200.mesh.unstructured_tet10(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
orders = {}
for seg in (1, 0):
    asm = b200.DeviceAssembler(rm, capi.MAT_ISO_VOL, [1.0, 1000.0], options={"segments": seg})
    orders[seg] = asm.ctx.element_order()
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.05)
    asm.set_state("displacement", u.ravel())
    asm.evaluate_resident()
    print("segments", seg, asm.ctx.last_evaluate_info(), orders[seg][:12])
    asm.close()
print("identical order:", np.array_equal(orders[0], orders[1]), "differing:", int((orders[0] != orders[1]).sum()))
c = X[conn[:, :4]].mean(axis=1)
for seg in (1, 0):
    d = np.linalg.norm(np.diff(c[orders[seg]], axis=0), axis=1)
    print("segments", seg, "mean centroid distance of consecutive elements", d.mean())

"""Throughput floors of the element kernels at (near) BASELINE sizes, so that the numbers of the workloads bench.py only
measures on request (--workload ...) are also produced -- and guarded -- by the GPU suite the driver runs.  Kernel time =
CUDA events the library records around every element-kernel launch (b200s_kernel_time), device-resident state, 3 warm-up
+ 6 timed evaluates.  The floors are ~70 % of what profiles/r02z_summary.md records on this pool's B200s: a regression
of a kernel trips them, box-to-box variation does not.  The measured rates are printed (pytest -s / the -rA summary)."""
import importlib

import numpy as np
import pytest

b200 = importlib.import_module("4c_b200")
capi = b200.capi
pytestmark = pytest.mark.gpu

# name, cell, mesh size, material, parameters, assembler keywords, floor in M elements/s (kernel time)
CASES = [
    ("hex8 CoupNeoHooke (config 2)", "HEX8", (100, 100, 100), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], {}, 320.0),
    ("hex8 F-bar", "HEX8", (80, 80, 80), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(tech=capi.TECH_FBAR), 170.0),
    ("hex8 EAS mild (tensor-core kernel)", "HEX8", (64, 64, 64), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(tech=capi.TECH_EAS_MILD), 70.0),
    ("hex8 EAS full (tensor-core kernel)", "HEX8", (48, 48, 48), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], dict(tech=capi.TECH_EAS_FULL), 40.0),
    ("tet10 iso/vol (config 4 mesh family)", "TET10", 60, capi.MAT_ISO_VOL, [1.0, 1000.0], {}, 180.0),
    ("hex27 SVK (config 3 mesh family)", "HEX27", (48, 48, 48), capi.MAT_SVK, [1000.0, 0.3], {}, 12.0),
    # hex20 through the hex27 kernel (padded tables); the general kernel runs this block at 2.6 M elements/s
    ("hex20 CoupNeoHooke (through the hex27 kernel)", "HEX20", (40, 40, 40), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], {}, 12.0),
    # hex18 likewise (nodes and Gauss points padded); general kernel: 4.4 M elements/s
    ("hex18 CoupNeoHooke (through the hex27 kernel)", "HEX18", (40, 40, 40), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], {}, 12.0),
]


@pytest.mark.parametrize("name,cell,n,mk,par,kw,floor", CASES, ids=[c[0].split(" (")[0].replace(" ", "_") for c in CASES])
def test_kernel_throughput_floor(name, cell, n, mk, par, kw, floor):
    if cell == "TET10":
        X, conn = b200.mesh.unstructured_tet10(n)
        rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
        h = 1.0 / n
    elif cell in ("HEX20", "HEX18"):
        X, conn = (b200.mesh.structured_hex20 if cell == "HEX20" else b200.mesh.structured_hex18)(n)
        rm = b200.partition.serial_mesh(getattr(b200.mesh, cell), X, conn)
        h = 1.0 / max(n)
    else:
        rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
        h = 1.0 / max(n)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    asm = b200.DeviceAssembler(rm, mk, par, **kw)
    asm.set_state("displacement", u.ravel())
    asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None)   # state on the device, first evaluate checked for errors
    for _ in range(3):
        asm.ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    asm.ctx.synchronize()
    asm.ctx.kernel_time(reset=True)
    for _ in range(6):
        asm.ctx.evaluate_async(capi.ACTION_NLNSTIFF)
    asm.ctx.synchronize()
    ms, count = asm.ctx.kernel_time()
    rate = rm.n_col_eles / (ms / max(count, 1) * 1e-3) / 1e6
    print(f"THROUGHPUT {name}: {rm.n_col_eles} elements, kernel {ms / max(count, 1):.3f} ms, {rate:.1f} M elements/s (floor {floor})")
    asm.close()
    assert count == 6 and rate >= floor, (name, rate, floor)

"""4c_b200 -- B200-native (sm_100a, FP64) solid element evaluate + assemble hot path of 4C.

The package name starts with a digit, so import it with
    import importlib; b200 = importlib.import_module("4c_b200")

Layout (only what the path needs):
  csrc/            hand-written CUDA kernels + the C ABI (include/b200_solid.h)
  capi.py          ctypes binding of lib4c_b200_solid.so
  mesh.py          synthetic hex8 / hex27 / tet10 meshes and displacement states
  partition.py     rank-local discretization views (row/column maps, halo lists)
  fe_export.py     row-element mode: export rows, extended graph and the export plan (FE_MATRIX assembly)
  assembler.py     host-side mirror of Discretization::evaluate + AssembleStrategy for one rank
"""
from . import capi, mesh, partition  # noqa: F401
from .assembler import DeviceAssembler, RowExportAssembler, parity_report  # noqa: F401
from . import fe_export  # noqa: F401

"""GPU parity at BASELINE.json's own sizes (configs 3, 4 and 5: more than 2^31 non-zeros per rank for 3 and 5,
64-bit CSR offsets) through row sampling, plus the library features added in round 2: the host-destination
pipeline (rows of finished phases travel while later element segments are evaluated), several element blocks
in one system matrix (b200s_attach_block), call-order independence of set_element_properties / set_mesh, and
struct_calc_reset_istep before the first update.  Tolerance 1e-12 (north_star); CSR pattern bit-exact."""
import ctypes
import importlib

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import sampled_rows

pytestmark = pytest.mark.gpu
b200 = importlib.import_module("4c_b200")
capi = b200.capi
RTOL = 1e-12


def _cudart():
    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    import torch  # torch ships its own copy
    import glob
    import os
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    return ctypes.CDLL(cands[0])


def _write_device(ptr, arr):
    """Host array -> device memory owned by the library (harness only: stands in for the NCCL halo when ONE
    rank of a partition is tested alone)."""
    arr = np.ascontiguousarray(arr, np.float64)
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rc = rt.cudaMemcpy(ctypes.c_void_p(ptr), ctypes.c_void_p(arr.ctypes.data), arr.nbytes, 1)
    assert rc == 0, rc


def _report(tag, r):
    print(f"{tag}: sampled-row parity worst block-row ratio {r['worst_blockrow_ratio']:.3g}, worst pure relative "
          f"error {r['worst_entry_rel']:.3g} (entries > {r['entry_floor']:g} x row scale), f ratio {r['worst_f_ratio']:.3g}, "
          f"{r['nodes']} nodes, {r['entries']} entries")


# ---- BASELINE config 3: 80^3 hex27 St.Venant-Kirchhoff, nnz = 2.36e9 > 2^31 ------------------------------
def test_config3_hex27_sampled_rows_at_full_size():
    n = (80, 80, 80)
    rm = b200.partition.box_partition(b200.mesh.HEX27, n, 0, 1, box=(0.8, 0.8, 0.8))
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.01)
    par = [1000.0, 0.3]
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, par)
    assert asm.nnz > 2 ** 31
    asm.set_state("displacement", u.ravel())
    _, f, err = asm.evaluate_resident()
    assert err.n_neg_jacobian == 0
    assert np.abs(f.reshape(-1, 3).sum(axis=0)).max() < 1e-9 * np.abs(f).max() * np.sqrt(len(f))
    # nodes spread over the whole row range, incl. the last rows (offsets beyond 2^31)
    rng = np.random.default_rng(3)
    nodes = np.unique(np.concatenate([rng.choice(rm.n_row_nodes, 120, replace=False),
                                      np.arange(rm.n_row_nodes - 30, rm.n_row_nodes)]))
    r = sampled_rows.check(asm.ctx, rm, orc.HEX27, orc.MAT_SVK, par, u, nodes, RTOL, nthreads=8)
    _report("config 3", r)
    assert r["ok"], r
    rp, _, _ = asm.ctx.get_value_rows(asm.n_rows - 3, asm.n_rows, want_values=False)
    assert rp[-1] > 0
    asm.close()


# ---- BASELINE config 4: 3.95 M unstructured tet10, nearly incompressible iso/vol Neo-Hooke -----------------
def test_config4_tet10_sampled_rows_at_full_size():
    X, conn = b200.mesh.unstructured_tet10(87)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    assert rm.n_col_eles == 6 * 87 ** 3
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / 87)
    par = [1.0, 1000.0]
    asm = b200.DeviceAssembler(rm, capi.MAT_ISO_VOL, par)
    asm.set_state("displacement", u.ravel())
    _, f, err = asm.evaluate_resident()
    assert err.n_neg_jacobian == 0 and 0.7 < err.min_detF < 1.3
    assert np.abs(f.reshape(-1, 3).sum(axis=0)).max() < 1e-9 * np.abs(f).max() * np.sqrt(len(f))
    nodes = sampled_rows.pick_nodes(rm, 200, 0, seed=4)
    r = sampled_rows.check(asm.ctx, rm, orc.TET10, orc.MAT_ISO_VOL, par, u, nodes, RTOL, nthreads=8)
    _report("config 4", r)
    assert r["ok"], r
    asm.close()


# ---- BASELINE config 5: one 200^3 rank of the 400^3 = 64 M element hex8 mesh on 8 ranks ---------------------
def test_config5_one_rank_of_8_sampled_rows():
    """Rank 0 of the 8-way box partition of 400^3 hex8 elements: 8.12 M column elements (one ghost layer on three
    sides), nnz 1.97e9.  The NCCL halo is replaced by writing the exact ghost displacements into the column
    vector (the halo itself is covered by tests/test_gpu_multirank.py and by bench.py's parity record at N > 1)."""
    n = (400, 400, 400)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 8, box=(4.0, 4.0, 4.0))
    assert rm.n_row_eles == 200 ** 3 and rm.n_col_eles == 201 ** 3
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 4.0, 0.01)
    par = [200.0, 0.25]
    ctx = capi.Context(0)   # a lone context: no communicator, the ghosts are written below
    ctx.set_mesh(rm.celltype, rm.X, rm.conn)
    ctx.set_material(capi.MAT_COUP_NEOHOOKE, par)
    nnz = ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
    assert nnz > 1.9e9
    _write_device(ctx.device_ptrs()["u_col"], u.ravel())
    u_row = np.ascontiguousarray(u[:rm.n_row_nodes].ravel())
    rc, err = ctx.evaluate(capi.ACTION_NLNSTIFF, u_row, None, None)
    assert err.n_neg_jacobian == 0 and 0.7 < err.min_detF < 1.3
    info = ctx.last_evaluate_info()
    assert info["ghost_elements"] == 201 ** 3 - 200 ** 3, info
    nodes = sampled_rows.pick_nodes(rm, 150, 60, seed=5)
    r = sampled_rows.check(ctx, rm, orc.HEX8, orc.MAT_COUP_NEOHOOKE, par, u, nodes, RTOL, nthreads=8)
    _report("config 5 (rank 0 of 8)", r)
    assert r["ok"], r
    ctx.close()


# ---- host-destination pipeline --------------------------------------------------------------------------------
@pytest.mark.parametrize("cell,n,mk,par", [("HEX8", (40, 36, 30), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]),
                                           ("HEX27", (12, 10, 9), capi.MAT_SVK, [1000.0, 0.3])])
def test_pipelined_host_copy_matches_resident_result(cell, n, mk, par):
    rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    asm = b200.DeviceAssembler(rm, mk, par)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate()                       # host K: pipelined copy
    info = asm.ctx.last_evaluate_info()
    assert info["row_phases"] > 1 and info["pipeline_chunks"] == info["row_phases"], info
    assert info["kernel_launches"] >= info["row_phases"]
    rK, rf, nbad = orc.evaluate_assemble(getattr(orc, cell), mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind, nthreads=8)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    # the monolithic launch (K stays on the device) gives the same matrix
    asm.evaluate_resident()
    assert asm.ctx.last_evaluate_info()["kernel_launches"] == 1
    K2 = asm.ctx.get_values()
    ok, wK, _ = b200.parity_report(rowptr, K2, rK, f, rf, RTOL)
    assert ok, wK
    # one phase: no overlap, same numbers
    ctx = capi.Context(0)
    ctx.set_option("pipeline_phases", 1)
    ctx.set_mesh(rm.celltype, rm.X, rm.conn)
    ctx.set_material(mk, par)
    ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
    K3, f3 = np.empty(asm.nnz), np.empty(asm.n_rows)
    ctx.evaluate(capi.ACTION_NLNSTIFF, np.ascontiguousarray(u.ravel()), K3, f3)
    assert ctx.last_evaluate_info()["pipeline_chunks"] == 0
    ok, wK, wf = b200.parity_report(rowptr, K3, rK, f3, rf, RTOL)
    assert ok, (wK, wf)
    ctx.close()
    asm.close()


def test_random_row_numbering_gets_one_phase():
    """Randomly renumbered tet10 nodes: row ranges have no spatial locality, the pipeline must not reorder by them."""
    X, conn = b200.mesh.unstructured_tet10(16)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_ISO_VOL, [1.0, 1000.0])
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / 16)
    asm.set_state("displacement", u.ravel())
    asm.evaluate()
    assert asm.ctx.last_evaluate_info()["row_phases"] == 1
    asm.close()


# ---- several element blocks in one system matrix ---------------------------------------------------------------
def _union_graph(blocks, n_row_nodes, n_col_nodes):
    """FillComplete of the matrix all blocks assemble into: node-blocked CSR (rows of a node share the pattern,
    columns ascending) from the node adjacency of every block."""
    import scipy.sparse as sp
    A = sp.csr_matrix((n_row_nodes, n_col_nodes), dtype=np.int8)
    for conn in blocks:
        nn = conn.shape[1]
        r = np.repeat(conn, nn, axis=1).ravel()
        c = np.tile(conn, (1, nn)).ravel()
        keep = r < n_row_nodes
        A = A + sp.csr_matrix((np.ones(keep.sum(), np.int8), (r[keep], c[keep])), shape=A.shape)
    A.sum_duplicates()
    A.sort_indices()
    cnt = np.diff(A.indptr)
    rowptr = np.zeros(3 * n_row_nodes + 1, np.int64)
    rowptr[1:] = np.cumsum(np.repeat(3 * cnt, 3))
    colind = np.empty(rowptr[-1], np.int32)
    for r in range(n_row_nodes):
        cols = (3 * A.indices[A.indptr[r]:A.indptr[r + 1]][:, None] + np.arange(3)).ravel()
        for d in range(3):
            colind[rowptr[3 * r + d]:rowptr[3 * r + d + 1]] = cols
    return rowptr, colind


def test_two_cell_types_and_two_materials_in_one_matrix():
    """hex27 elements (SVK) in the lower half, hex8 elements (two MAT ids: CoupNeoHooke / iso-vol split) on the fine
    grid of the upper half, all on the node set of the hex27 grid: the three blocks assemble into ONE CSR like the
    reference's element loop does for any mix of element types and MAT ids
    (4C_fem_discretization_evaluate.cpp:85-117)."""
    n = (4, 3, 4)
    X, conn27 = b200.mesh.structured_hex(b200.mesh.HEX27, n)
    nx, ny, nz = n
    e27 = np.arange(len(conn27))
    lower = (e27 // (nx * ny)) < nz // 2
    blkA = conn27[lower]
    # fine hex8 cells of the upper half on the (2n+1)^3 node grid
    gx, gy, gz = 2 * nx + 1, 2 * ny + 1, 2 * nz + 1
    cells = [(i, j, k) for k in range(nz, 2 * nz) for j in range(2 * ny) for i in range(2 * nx)]
    off = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    c8 = np.array([[((k + c) * gy + (j + b)) * gx + (i + a) for a, b, c in off] for i, j, k in cells], np.int32)
    blkB, blkC = c8[0::2], c8[1::2]
    nnodes = len(X)
    ids = np.arange(nnodes, dtype=np.int32)
    rowptr, colind = _union_graph([blkA, blkB, blkC], nnodes, nnodes)
    u = 2.0 * b200.mesh.displacement_field(X, ids, 1.0, 0.1)
    mats = [(capi.HEX27, blkA, capi.MAT_SVK, [1000.0, 0.3]), (capi.HEX8, blkB, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]),
            (capi.HEX8, blkC, capi.MAT_ISO_VOL, [1.0, 1000.0])]
    d = np.arange(3, dtype=np.int32)
    dof = (3 * ids[:, None] + d).astype(np.int32).ravel()
    ctxs = []
    for k, (ct, conn, mk, par) in enumerate(mats):
        c = capi.Context(0)
        c.set_mesh(ct, X, conn)
        c.set_material(mk, par)
        if k == 0:
            c.set_graph(rowptr, colind, dof, dof)
        else:
            c.attach_block(ctxs[0])
        ctxs.append(c)
    u_row = np.ascontiguousarray(u.ravel())
    K, f = np.empty(rowptr[-1]), np.empty(3 * nnodes)
    ctxs[0].evaluate(capi.ACTION_NLNSTIFF, u_row, None, None)
    ctxs[1].evaluate(capi.ACTION_NLNSTIFF, None, None, None)
    ctxs[2].evaluate(capi.ACTION_NLNSTIFF, None, K, f)       # the last block hands out the sum
    rK, rf = np.zeros_like(K), np.zeros_like(f)
    for ct, conn, mk, par in mats:
        k_, f_, nbad = orc.evaluate_assemble(ct, mk, par, conn, X.ravel(), u.ravel(), ids, ids, nnodes, rowptr, colind)
        assert nbad == 0
        rK += k_
        rf += f_
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok, (wK, wf)
    # a block whose couplings are missing from the primary's graph is refused
    far = capi.Context(0)
    far.set_mesh(capi.HEX8, X, np.array([[0, 1, 2, 3, nnodes - 4, nnodes - 3, nnodes - 2, nnodes - 1]], np.int32))
    far.set_material(capi.MAT_SVK, [1.0, 0.3])
    with pytest.raises(capi.B200SolidError) as ei:
        far.attach_block(ctxs[0])
    assert ei.value.code == capi.ERR_UNSUPPORTED
    far.close()
    for c in reversed(ctxs):
        c.close()


# ---- call order: element properties before the mesh (device element order is final only after set_mesh) --------
@pytest.mark.parametrize("tech,prestress", [(capi.TECH_EAS_MILD, capi.PRESTRESS_NONE), (capi.TECH_NONE, capi.PRESTRESS_MULF)])
def test_properties_before_mesh_on_a_graded_mesh(tech, prestress):
    n = (7, 6, 5)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    X = rm.X.copy()
    X[:, 0] = X[:, 0] ** 1.7 * 1.3            # graded and stretched: every element has its own Jacobian
    X[:, 1] = X[:, 1] * (1.0 + 0.4 * X[:, 2])
    rm.X = X
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.1)
    par = [200.0, 0.25]
    res = []
    for order in ("mesh_first", "properties_first", "mesh_twice"):
        ctx = capi.Context(0)
        if order == "properties_first":
            ctx.set_element_properties(tech=tech, prestress=prestress)
            ctx.set_mesh(rm.celltype, rm.X, rm.conn)
        else:
            ctx.set_mesh(rm.celltype, rm.X[::-1].copy() if order == "mesh_twice" else rm.X, rm.conn)
            ctx.set_element_properties(tech=tech, prestress=prestress)
            if order == "mesh_twice":
                ctx.set_mesh(rm.celltype, rm.X, rm.conn)
        ctx.set_material(capi.MAT_COUP_NEOHOOKE, par)
        nnz = ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
        K, f = np.empty(nnz), np.empty(3 * rm.n_row_nodes)
        ctx.evaluate(capi.ACTION_NLNSTIFF, np.ascontiguousarray(u.ravel()), K, f)
        res.append((K, f))
        ctx.close()
    for K, f in res[1:]:
        assert np.abs(K - res[0][0]).max() <= 1e-12 * np.abs(res[0][0]).max()
        assert np.abs(f - res[0][1]).max() <= 1e-12 * np.abs(res[0][1]).max()


def test_reset_step_before_first_update_restores_fresh_history():
    n = (4, 3, 3)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], tech=capi.TECH_EAS_MILD)
    asm.set_state("displacement", u.ravel())
    K0, f0, _ = asm.evaluate()
    asm.recover(0.3 * u[:rm.n_row_nodes].ravel(), 0.5)   # alpha moves away from zero, old_step_length = 0.5
    assert asm.ctx.eas_old_step_length() == 0.5
    assert np.abs(asm.ctx.get_element_history()).max() > 0
    asm.ctx.reset_step()                                   # diverged first step: back to the fresh history
    assert asm.ctx.eas_old_step_length() == 1.0
    h = asm.ctx.get_element_history().reshape(rm.n_col_eles, -1)
    assert np.abs(h[:, :18]).max() == 0.0                   # alpha_inc, alpha
    K1, f1, _ = asm.evaluate()
    assert np.abs(K1 - K0).max() <= 1e-12 * np.abs(K0).max() and np.abs(f1 - f0).max() <= 1e-12 * np.abs(f0).max()
    # snapshot / restore carries old_step_length
    asm.recover(0.3 * u[:rm.n_row_nodes].ravel(), 0.25)
    asm.ctx.update_step()
    asm.ctx.recover_linesearch(0.75)
    asm.ctx.reset_step()
    assert asm.ctx.eas_old_step_length() == 0.25
    asm.close()


def test_dirichlet_diagonal_with_permuted_column_map():
    """Matrix-local column ids that do not list the owned dofs in row order: the 1.0 of apply_dirichlet must still
    land on the diagonal entry."""
    import scipy.sparse as sp
    n = (3, 3, 2)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    nn = rm.n_col_nodes
    perm = np.random.default_rng(7).permutation(nn).astype(np.int32)      # column-node id of node k
    rowptr0, colind0 = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, perm, rm.n_row_nodes)
    d = np.arange(3, dtype=np.int32)
    dof_row = (3 * np.arange(nn, dtype=np.int32)[:, None] + d).ravel()
    dof_col = (3 * perm[:, None] + d).astype(np.int32).ravel()
    ctx = capi.Context(0)
    ctx.set_mesh(rm.celltype, rm.X, rm.conn)
    ctx.set_material(capi.MAT_SVK, [1000.0, 0.3])
    ctx.set_graph(rowptr0, colind0, dof_row, dof_col)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    ctx.evaluate(capi.ACTION_NLNSTIFF, np.ascontiguousarray(u.ravel()), None, None)
    tog = np.zeros(3 * nn)
    tog[[0, 7, 20, 3 * nn - 1]] = 1.0
    ctx.apply_dirichlet(tog, True)
    K = ctx.get_values()
    A = sp.csr_matrix((K, colind0, rowptr0), shape=(3 * nn, 3 * nn)).toarray()
    for r in np.nonzero(tog)[0]:
        node, dd = divmod(int(r), 3)
        row = A[r]
        assert row[3 * perm[node] + dd] == 1.0 and np.count_nonzero(row) == 1
    ctx.close()


# ---- MAT_VISCONEOHOOKE: finite-strain history material on the device Gauss-point state store ---------------------
VISCO = dict(YOUNGS_SLOW=100.0, POISSON=0.3, DENS=1e-10, YOUNGS_FAST=250.0, RELAX=10.0, THETA=0.5)
VISCO6 = [VISCO[k] for k in ("YOUNGS_SLOW", "POISSON", "DENS", "YOUNGS_FAST", "RELAX", "THETA")]


@pytest.mark.parametrize("cell,n,tech", [("HEX8", (6, 5, 4), "none"), ("HEX27", (3, 2, 2), "none"), ("TET10", 3, "none"),
                                         ("HEX8", (5, 4, 3), "eas_mild"), ("HEX8", (4, 3, 3), "eas_full")])
def test_visco_neohooke_matches_oracle_over_steps(cell, n, tech):
    """Two time steps with a state commit in between and a changed displacement: K, f and the stored state of the
    CUDA path against the oracle (pinned to the reference's creep known answer, tests/test_visco_material.py)."""
    ct = getattr(orc, cell)
    if cell == "TET10":
        X, conn = b200.mesh.unstructured_tet10(n)
        rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
        h = 1.0 / n
    else:
        rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
        h = 1.0 / max(n)
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    dt = 0.2
    par = orc.visco_neohooke_params(dt=dt, **VISCO)
    otech = {"none": orc.TECH_NONE, "eas_mild": orc.TECH_EAS_MILD, "eas_full": orc.TECH_EAS_FULL}[tech]
    dtech = {"none": capi.TECH_NONE, "eas_mild": capi.TECH_EAS_MILD, "eas_full": capi.TECH_EAS_FULL}[tech]
    asm = b200.DeviceAssembler(rm, capi.MAT_VISCONEOHOOKE, VISCO6, tech=dtech)
    rowptr, colind = asm.graph()
    ngp = orc.num_gp(ct)
    mh = np.zeros((rm.n_col_eles, ngp, orc.VISCO_HIST))
    eas = orc.eas_new_history(otech, rm.n_col_eles) if tech != "none" else None
    # without a time step size the evaluate fails like the reference's assertion
    asm.set_state("displacement", u.ravel())
    rc, _ = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
    assert rc == capi.ERR_INVALID
    asm.ctx.set_time_step_size(dt)
    assert asm.ctx.L.b200s_material_history_size(asm.ctx.h) == mh.size
    for step, fac in enumerate((1.0, 1.4)):
        asm.set_state("displacement", (fac * u).ravel())
        K, f, err = asm.evaluate()
        orc.set_material_history(mh)
        rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_VISCONEOHOOKE, par, rm.conn, rm.X.ravel(), (fac * u).ravel(),
                                             rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind, tech=otech,
                                             hist=eas)
        orc.set_material_history(None)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12 if tech == "none" else 1e-11)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (step, wK, wf)
        dh = asm.ctx.get_material_history().reshape(mh.shape)
        assert np.abs(dh - mh).max() <= 1e-12 * np.abs(mh).max()
        asm.ctx.update_step()
        orc.material_update(orc.MAT_VISCONEOHOOKE, mh)
        dh = asm.ctx.get_material_history().reshape(mh.shape)
        assert np.abs(dh[..., 12:]).max() == 0.0 and np.abs(dh[..., :12] - mh[..., :12]).max() <= 1e-12 * np.abs(mh).max()
    # restart: the state store round-trips through get / set (caller's element numbering)
    blob = asm.ctx.get_material_history()
    asm2 = b200.DeviceAssembler(rm, capi.MAT_VISCONEOHOOKE, VISCO6, tech=dtech)
    asm2.ctx.set_time_step_size(dt)
    asm2.ctx.set_material_history(blob)
    if tech != "none":
        asm2.ctx.set_element_history(asm.ctx.get_element_history())
    for a in (asm, asm2):
        a.set_state("displacement", (1.7 * u).ravel())
    K1, f1, _ = asm.evaluate()
    K2, f2, _ = asm2.evaluate()
    assert np.abs(K1 - K2).max() <= 1e-12 * np.abs(K1).max() and np.abs(f1 - f2).max() <= 1e-12 * np.abs(f1).max()
    asm.close()
    asm2.close()


def test_reference_visco_creep_known_answer_through_cuda_path():
    """tests/input_files/viscoNeoHooke_creep.4C.yaml with the CUDA path inside the generalised-alpha loop: device EAS-mild
    kernel with MAT_VISCONEOHOOKE, consistent mass matrix from the device mass action, struct_calc_recover and the
    state commit on the device; RESULT DESCRIPTION displacements of node 122 (tolerance 1e-9)."""
    import json
    import os
    from tests.helpers.newton import solve_genalpha_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "integration_hex8_visco_creep.json")))
    X = np.array(case["coords"])
    conn = np.array(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    m = case["material"]
    in6 = [m[k] for k in ("YOUNGS_SLOW", "POISSON", "DENS", "YOUNGS_FAST", "RELAX", "THETA")]
    asm = b200.DeviceAssembler(rm, capi.MAT_VISCONEOHOOKE, in6, tech=capi.TECH_EAS_MILD, density=m["DENS"])
    asm.ctx.set_time_step_size(case["timestep"])
    rowptr, colind = asm.graph()
    asm.set_state("displacement", np.zeros(asm.n_rows))
    asm.evaluate("struct_calc_nlnstiffmass")
    Mv, _ = asm.mass(want_matrix=True)
    Mo, _ = orc.assemble_mass(orc.HEX8, m["DENS"], conn, X.ravel(), rm.node_row_lid, rm.node_col_lid, len(X), rowptr, colind)
    assert np.abs(Mv - Mo).max() <= 1e-12 * np.abs(Mo).max()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_genalpha_case(case, evaluate, Mv, recover=lambda du: asm.recover(du),
                            update_step=lambda u: asm.ctx.update_step())
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], (e, u[e["node"], e["dof"]])
    assert abs(u[121, 0] - 0.0006320013656913447) < 1e-11
    asm.close()


# ---- block-upper transfer of a symmetric K with the threaded host mirror -------------------------------------------
@pytest.mark.parametrize("cell,n,mk,par,tech", [("HEX8", (40, 36, 30), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], capi.TECH_NONE),
                                                ("HEX27", (12, 10, 9), capi.MAT_SVK, [1000.0, 0.3], capi.TECH_NONE),
                                                ("HEX8", (30, 28, 26), capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], capi.TECH_EAS_MILD)])
@pytest.mark.parametrize("nt_stores,threads", [(1, 0), (0, 3)])
def test_symmetric_transfer_matches_oracle(cell, n, mk, par, tech, nt_stores, threads):
    rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    asm = b200.DeviceAssembler(rm, mk, par, tech=tech,
                               options={"symmetric_transfer": 1, "nt_stores": nt_stores, "host_threads": threads})
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K = np.full(asm.nnz, np.nan)
    f = np.empty(asm.n_rows)
    asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, K, f)
    assert asm.ctx.symmetric_transfer_status() == "in use", asm.ctx.symmetric_transfer_status()
    moved = asm.ctx.last_d2h_value_bytes()
    assert 0.5 * 8 * asm.nnz < moved < 0.6 * 8 * asm.nnz, (moved, 8 * asm.nnz)
    otech = {capi.TECH_NONE: orc.TECH_NONE, capi.TECH_EAS_MILD: orc.TECH_EAS_MILD}[tech]
    hist = orc.eas_new_history(otech, rm.n_col_eles) if tech != capi.TECH_NONE else None
    rK, rf, nbad = orc.evaluate_assemble(getattr(orc, cell), mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind, tech=otech, hist=hist,
                                         nthreads=1 if hist is not None else 8)
    assert np.isfinite(K).all()                                   # every entry of the full CSR was written
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL if tech == capi.TECH_NONE else 1e-10)
    assert ok and nbad == 0, (wK, wf)
    # the mirrored matrix against the one resident on the device (same evaluate): upper blocks identical, lower
    # blocks differ only by the summation order of the atomics
    Kd = asm.ctx.get_values()
    ok, wK, _ = b200.parity_report(rowptr, K, Kd, f, f, 1e-13)
    assert ok, wK
    # off-diagonal node blocks are exact transposes by construction (the diagonal 3x3 blocks travel whole)
    import scipy.sparse as sp
    A = sp.csr_matrix((K, colind, rowptr), shape=(asm.n_rows, asm.n_rows))
    D = abs(A - A.T).tocoo()
    off = (D.row // 3) != (D.col // 3)
    assert D.data[off].max(initial=0.0) == 0.0
    assert D.data.max(initial=0.0) <= 1e-13 * np.abs(K).max()
    asm.close()


def test_symmetric_transfer_is_refused_for_fbar():
    """F-bar has a non-symmetric K_e: the option is ignored, the whole matrix travels."""
    n = (40, 36, 30)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], tech=capi.TECH_FBAR, options={"symmetric_transfer": 1})
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate()
    assert asm.ctx.last_d2h_value_bytes() == 8 * asm.nnz
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(),
                                         rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind, tech=orc.TECH_FBAR,
                                         nthreads=8)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    asm.close()


# ---- tet4 and hex20 blocks (general kernel; src/solid_ele/4C_solid_ele_factory.hpp:36-38 instantiates them) ----
GEN_MATS = {capi.MAT_COUP_NEOHOOKE: [200.0, 0.25], capi.MAT_ISO_VOL: [1.0, 1000.0], capi.MAT_SVK: [1000.0, 0.3]}


def _general_mesh(cell, n=None):
    if cell == "tet4":
        X, conn = b200.mesh.unstructured_tet4(n or 5)
        return orc.TET4, b200.partition.serial_mesh(b200.mesh.TET4, X, conn), 1.0 / (n or 5)
    if cell == "wedge6":
        X, conn = b200.mesh.structured_wedge6(n or (5, 4, 3))
        return orc.WEDGE6, b200.partition.serial_mesh(b200.mesh.WEDGE6, X, conn), 0.25
    if cell == "pyramid5":
        X, conn = b200.mesh.structured_pyramid5(n or (4, 3, 3))
        return orc.PYRAMID5, b200.partition.serial_mesh(b200.mesh.PYRAMID5, X, conn), 0.25
    if cell == "hex18":
        X, conn = b200.mesh.structured_hex18(n or (4, 3, 3))
        return orc.HEX18, b200.partition.serial_mesh(b200.mesh.HEX18, X, conn), 0.3
    X, conn = b200.mesh.structured_hex20(n or (4, 3, 3))
    return orc.HEX20, b200.partition.serial_mesh(b200.mesh.HEX20, X, conn), 0.3


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
@pytest.mark.parametrize("mk", list(GEN_MATS))
@pytest.mark.parametrize("kinem", [capi.KINEM_NONLINEAR_TOTLAG, capi.KINEM_LINEAR])
def test_tet4_hex20_match_oracle(cell, mk, kinem):
    ct, rm, h = _general_mesh(cell)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    asm = b200.DeviceAssembler(rm, mk, GEN_MATS[mk], kinem=kinem)
    rowptr, colind = asm.graph()
    o = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert rowptr.tobytes() == o[0].tobytes() and colind.tobytes() == o[1].tobytes()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    rK, rf, nbad = orc.evaluate_assemble(ct, mk, GEN_MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind, nthreads=4,
                                         tech=orc.TECH_LINEAR if kinem == capi.KINEM_LINEAR else orc.TECH_NONE)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0 and err.n_neg_jacobian == 0, (wK, wf)
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_tet4_hex20_mass_stress_energy(cell):
    ct, rm, h = _general_mesh(cell)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3], density=7.8)
    rowptr, colind = asm.graph()
    acc = np.random.default_rng(2).uniform(-1, 1, 3 * rm.n_col_nodes)
    asm.set_state("displacement", u.ravel())
    asm.set_state("acceleration", acc)
    for action, lumped in (("struct_calc_nlnstiffmass", False), ("struct_calc_nlnstifflmass", True)):
        K, f, _ = asm.evaluate(action)
        rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_SVK, [1000.0, 0.3], rm.conn, rm.X.ravel(), u.ravel(),
                                          rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (wK, wf)
        rM, rfi = orc.assemble_mass(ct, 7.8, rm.conn, rm.X.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes,
                                    rowptr, colind, lumped=lumped, acc_col=acc)
        M, fi = asm.mass()
        assert np.abs(M - rM).max() <= RTOL * np.abs(rM).max()
        assert np.abs(fi - rfi).max() <= RTOL * np.abs(rfi).max()
    st, en = asm.calc_stress("cauchy", "ea")
    rst, ren = orc.calc_stress(ct, orc.MAT_SVK, [1000.0, 0.3], rm.conn, rm.X.ravel(), u.ravel(), orc.STRESS_CAUCHY,
                               orc.STRAIN_EA)
    assert st.shape == rst.shape
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max() and np.abs(en - ren).max() <= RTOL * np.abs(ren).max()
    e = asm.calc_energy()
    re_ = orc.calc_energy(ct, orc.MAT_SVK, [1000.0, 0.3], rm.conn, rm.X.ravel(), u.ravel())
    assert abs(e - re_) <= 1e-11 * abs(re_)
    asm.close()


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_reference_tet4_hex20_known_answers_through_cuda_path(cell):
    """RESULT DESCRIPTION displacements of tests/input_files/solid_ele_{tet4,hex20,wedge6,pyramid5}_Standard_stressout.4C.yaml."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_{cell}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(getattr(b200.mesh, cell.upper()), X, conn)
    mat = case["material"]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [mat["YOUNG"], mat["NUE"]])
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


def test_tet4_inverted_element_is_reported():
    ct, rm, h = _general_mesh("tet4", 2)
    X = rm.X.copy()
    c = rm.conn[3]
    X[c[0]], X[c[1]] = X[c[1]].copy(), X[c[0]].copy()   # swaps orientation of every tet at these two nodes
    rm2 = b200.partition.serial_mesh(b200.mesh.TET4, X, rm.conn)
    asm = b200.DeviceAssembler(rm2, capi.MAT_SVK, [1000.0, 0.3])
    asm.set_state("displacement", np.zeros(3 * rm2.n_row_nodes))
    rowptr, colind = asm.graph()
    K, f, err = asm.evaluate(check=False)
    _, _, nbad = orc.evaluate_assemble(ct, orc.MAT_SVK, [1000.0, 0.3], rm2.conn, rm2.X.ravel(), np.zeros(3 * rm2.n_col_nodes),
                                       rm2.node_row_lid, rm2.node_col_lid, rm2.n_row_nodes, rowptr, colind)
    assert nbad > 0 and err.n_neg_jacobian == nbad
    asm.close()


# ---- hex8 EAS: tensor-core (DMMA) kernel against the warp-per-element FMA kernel and the oracle ----
@pytest.mark.parametrize("tech,otech", [(capi.TECH_EAS_MILD, orc.TECH_EAS_MILD), (capi.TECH_EAS_FULL, orc.TECH_EAS_FULL)])
@pytest.mark.parametrize("mk", [capi.MAT_COUP_NEOHOOKE, capi.MAT_SVK])
def test_eas_tensor_core_kernel_matches_fma_kernel_and_oracle(tech, otech, mk):
    rm = b200.partition.box_partition(b200.mesh.HEX8, (7, 6, 5), 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    par = GEN_MATS[mk]
    res = {}
    for mma in (1, 0):
        asm = b200.DeviceAssembler(rm, mk, par, tech=tech, options={"eas_mma": mma})
        rowptr, colind = asm.graph()
        asm.set_state("displacement", u.ravel())
        K, f, err = asm.evaluate()
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert err.n_neg_jacobian == 0
        assert np.abs(f2 - f).max() <= 1e-13 * np.abs(f).max()
        res[mma] = (K, f, asm.ctx.get_element_history().copy())
        asm.close()
    ohist = orc.eas_new_history(otech, rm.n_col_eles)
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                         rm.n_row_nodes, rowptr, colind, tech=otech, hist=ohist)
    for mma in (1, 0):
        ok, wK, wf = b200.parity_report(rowptr, res[mma][0], rK, res[mma][1], rf, 1e-10)
        print("eas_mma", mma, "worst K", wK * 1e-10, "worst f", wf * 1e-10)
        assert ok and nbad == 0, (mma, wK, wf)
    ok, wK, wf = b200.parity_report(rowptr, res[1][0], res[0][0], res[1][1], res[0][1], 1e-10)
    assert ok, (wK, wf)
    h1, h0 = res[1][2].reshape(ohist.shape), res[0][2].reshape(ohist.shape)
    assert np.abs(h1 - h0).max() <= 1e-8 * np.abs(h0).max()


def test_wedge6_pyramid5_surface_loads_are_refused():
    """Cells with two surface shapes have no face table on the device: the library says so instead of integrating a
    wrong shape (volume loads work)."""
    ct, rm, h = _general_mesh("wedge6", (2, 2, 2))
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    with pytest.raises(capi.B200SolidError) as ei:
        asm.add_surface_neumann("live", np.zeros((1, 4), np.int32), [1, 0, 0], [1.0, 0.0, 0.0])
    assert ei.value.code == capi.ERR_UNSUPPORTED
    c = asm.add_volume_neumann([1, 0, 0], [1.0, 0.0, 0.0])
    asm.ctx.set_neumann_factors(c, [1.0, 1.0, 1.0])
    f = asm.evaluate_neumann()
    assert abs(f.reshape(-1, 3)[:, 0].sum() - 1.0) < 1e-12   # unit cube, unit body force
    asm.close()


@pytest.mark.parametrize("cell", ["hex8", "hex20", "hex27", "tet4", "pyramid5"])
def test_reference_linear_kinematics_known_answers_through_cuda_path(cell):
    """KINEM linear + MAT_Struct_StVenantKirchhoff: RESULT DESCRIPTION displacements (1e-12) of the reference's
    solid_ele_<cell>_Standard_linear.4C.yaml (surface + volume load) with K and f from the device kernels."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_{cell}_linear.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(getattr(b200.mesh, cell.upper()), X, conn)
    mat = case["material"]
    assert mat["kind"] == "MAT_Struct_StVenantKirchhoff" and case["kinem"] == "linear"
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [mat["YOUNG"], mat["NUE"]], kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


# ---- PRESTRESS_TECH mulf on the general cells (MulfFormulation: src/solid_ele/4C_solid_ele_factory.hpp:59-60) ----
@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_general_cells_mulf_evaluate_update_and_history_match_oracle(cell):
    ct, rm, h = _general_mesh(cell)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    mk, par = capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]
    asm = b200.DeviceAssembler(rm, mk, par, prestress=capi.PRESTRESS_MULF)
    rowptr, colind = asm.graph()
    ohist = orc.mulf_new_history(ct, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    ngp = orc.num_gp(ct)
    for scale in (1.0, -0.6, 0.3):
        us = scale * u.ravel()
        asm.set_state("displacement", us[:3 * rm.n_row_nodes])
        K, f, _ = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), us, *maps, tech=orc.TECH_MULF, hist=ohist)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok and nbad == 0, (scale, wK, wf)
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
        asm.update_prestress()
        orc.update_prestress(ct, rm.conn, rm.X.ravel(), us, ohist)
        hh = asm.ctx.get_gp_history().reshape(rm.n_col_eles, ngp, 18)
        assert np.abs(hh - ohist[:, :, :18]).max() <= 1e-13 * np.abs(ohist).max()
    asm.close()


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5"])
def test_reference_general_cells_mulf_known_answers_through_cuda_path(cell):
    """RESULT DESCRIPTION displacements of tests/input_files/solid_ele_{tet4,hex20,wedge6,pyramid5}_Standard_mulf.4C.yaml
    (MAT_ElastHyper{CoupNeoHooke, CoupAnisoExpo} on the element fibres, two prestress steps + eight load steps) with the
    CUDA path doing every evaluate and prestress update."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_{cell}_mulf.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(getattr(b200.mesh, cell.upper()), X, conn)
    nh, fib = case["summands"][0]["params"], case["summands"][1]["params"]
    mat = [("COUP_NEOHOOKE", [nh["YOUNG"], nh["NUE"]]), ("COUP_ANISOEXPO", [fib["K1"], fib["K2"], fib["K1COMP"], fib["K2COMP"]])]
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, mat, prestress=capi.PRESTRESS_MULF, fibers=np.asarray(case["fibers"], float))
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    def up(u):
        asm.set_state("displacement", u)
        asm.update_prestress()

    u = solve_case(case, ev, update_prestress=up)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, u[e["node"], e["dof"]])
    asm.close()


# ---- TECH fbar on pyramid5 and TECH fbar + PRESTRESS_TECH mulf (hex8, pyramid5) through the general kernel -----------
def _fbar_mesh(cell):
    if cell == "hex8":
        return orc.HEX8, b200.partition.box_partition(b200.mesh.HEX8, (4, 3, 3), 0, 1), 0.25
    return _general_mesh("pyramid5")


@pytest.mark.parametrize("cell", ["hex8", "pyramid5"])
@pytest.mark.parametrize("mulf", [False, True])
def test_general_kernel_fbar_and_mulf_fbar_match_oracle(cell, mulf):
    """FbarEvaluators / FBarMulfEvaluators = hex8, pyramid5 (src/solid_ele/4C_solid_ele_factory.hpp:59-60,79-81): K (not
    symmetric), f, the histories after update_prestress and the stress output against the oracle; for hex8 without MULF the
    general kernel (option general_kernel) also against the fast F-bar kernel."""
    ct, rm, h = _fbar_mesh(cell)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    mk, par = capi.MAT_ISO_VOL, [1.0, 1000.0]
    asm = b200.DeviceAssembler(rm, mk, par, tech=capi.TECH_FBAR, prestress=capi.PRESTRESS_MULF if mulf else capi.PRESTRESS_NONE,
                               options={"general_kernel": 1})
    rowptr, colind = asm.graph()
    ohist = orc.mulf_new_history(ct, rm.n_col_eles, fbar=True) if mulf else None
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    for scale in (1.0, -0.6, 0.3):
        us = scale * u.ravel()
        asm.set_state("displacement", us[:3 * rm.n_row_nodes])
        K, f, err = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), us, *maps,
                                             tech=orc.TECH_MULF_FBAR if mulf else orc.TECH_FBAR, hist=ohist)
        # after a prestress update the two histories differ by round-off (1e-13 below) and the nearly incompressible
        # material (kappa / mue = 1000) amplifies that in p = kappa (J - 1): 1e-11 from the second state on
        tol = RTOL if (not mulf or scale == 1.0) else 1e-11
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, tol)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (scale, wK, wf)
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= tol * np.abs(rf).max()
        if mulf:
            asm.update_prestress()
            orc.update_prestress(ct, rm.conn, rm.X.ravel(), us, ohist)
            hh = asm.ctx.get_gp_history().reshape(rm.n_col_eles, orc.num_gp(ct) + 1, 18)
            assert np.abs(hh - ohist[:, :, :18]).max() <= 1e-13 * np.abs(ohist).max()
    if cell == "hex8" and not mulf:
        fast = b200.DeviceAssembler(rm, mk, par, tech=capi.TECH_FBAR)
        fast.set_state("displacement", us[:3 * rm.n_row_nodes])
        Kf, ff, _ = fast.evaluate()
        ok, wK, wf = b200.parity_report(rowptr, K, Kf, f, ff, RTOL)
        assert ok, (wK, wf)
        fast.close()
    asm.close()


@pytest.mark.parametrize("name", ["pyramid5_fbar", "hex8_mulf_fbar", "hex8_fbar"])
def test_reference_fbar_known_answers_through_general_kernel(name):
    """RESULT DESCRIPTION displacements of tests/input_files/solid_ele_pyramid5_Standard_fbar.4C.yaml,
    solid_ele_hex8_Standard_mulf_fbar.4C.yaml and solid_ele_hex8_Standard_fbar.4C.yaml (fibre material) with the CUDA path
    doing every evaluate and prestress update."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_{name}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(getattr(b200.mesh, case["celltype"].upper()), X, conn)
    nh, fib = case["summands"][0]["params"], case["summands"][1]["params"]
    mat = [("COUP_NEOHOOKE", [nh["YOUNG"], nh["NUE"]]), ("COUP_ANISOEXPO", [fib["K1"], fib["K2"], fib["K1COMP"], fib["K2COMP"]])]
    mulf = case["prestresstime"] > 0
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, mat, tech=capi.TECH_FBAR,
                               prestress=capi.PRESTRESS_MULF if mulf else capi.PRESTRESS_NONE,
                               fibers=np.asarray(case["fibers"], float), options={"general_kernel": 1})
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    def up(u):
        asm.set_state("displacement", u)
        asm.update_prestress()

    u = solve_case(case, ev, update_prestress=up if mulf else None)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, u[e["node"], e["dof"]])
    asm.close()


# ---- MAT_Struct_PlasticNlnLogNeoHooke: finite-strain plasticity on the device Gauss-point state store -----------------
PLASTIC = dict(YOUNG=206.9, NUE=0.29, DENS=0.0, YIELD=0.45, ISOHARD=0.12924, SATHARDENING=0.715, HARDEXPO=16.93, VISC=1.0,
               RATE_DEPENDENCY=0.1, TOL=1e-12)
PLASTIC10 = [PLASTIC[k] for k in ("YOUNG", "NUE", "DENS", "YIELD", "ISOHARD", "SATHARDENING", "HARDEXPO", "VISC",
                                  "RATE_DEPENDENCY", "TOL")]


@pytest.mark.parametrize("cell,n,tech", [("HEX8", (6, 5, 4), "none"), ("HEX27", (3, 2, 2), "none"), ("TET10", 3, "none"),
                                         ("HEX8", (5, 4, 3), "eas_mild"), ("HEX8", (4, 3, 3), "eas_full")])
def test_plastic_nlnlog_matches_oracle_over_steps(cell, n, tech):
    """Two load steps into the plastic range with a state commit in between: K, f and the stored state {accumulated
    plastic strain, Cp^-1} of the CUDA path (evaluation from C with the Cholesky factor; hex8 on the fast kernel, the
    other cells on the general kernel, EAS on the tensor-core kernel) against the oracle's F-based restatement, which is
    pinned to the reference's necking known answer (tests/test_plastic_material.py)."""
    ct = getattr(orc, cell)
    if cell == "TET10":
        X, conn = b200.mesh.unstructured_tet10(n)
        rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
        h = 1.0 / n
    else:
        rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
        h = 1.0 / max(n)
    u = 1.5 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    dt = 0.175
    par = orc.plastic_nlnlog_params(dt=dt, **PLASTIC)
    otech = {"none": orc.TECH_NONE, "eas_mild": orc.TECH_EAS_MILD, "eas_full": orc.TECH_EAS_FULL}[tech]
    dtech = {"none": capi.TECH_NONE, "eas_mild": capi.TECH_EAS_MILD, "eas_full": capi.TECH_EAS_FULL}[tech]
    asm = b200.DeviceAssembler(rm, capi.MAT_PLASTIC_NLNLOG, PLASTIC10, tech=dtech)
    rowptr, colind = asm.graph()
    ngp = orc.num_gp(ct)
    mh = orc.plastic_new_history(rm.n_col_eles, ngp)
    eas = orc.eas_new_history(otech, rm.n_col_eles) if tech != "none" else None
    asm.set_state("displacement", u.ravel())
    rc, _ = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
    assert rc == capi.ERR_INVALID  # no time step size yet
    asm.ctx.set_time_step_size(dt)
    assert asm.ctx.L.b200s_material_history_size(asm.ctx.h) == mh.size
    assert np.array_equal(asm.ctx.get_material_history().reshape(mh.shape), mh)  # fresh: Cp^-1 = I
    # the tangent of a yielding point divides by differences of principal stretches: 1e-10 instead of 1e-12
    tol = 1e-10
    for step, fac in enumerate((1.0, 1.4)):
        asm.set_state("displacement", (fac * u).ravel())
        K, f, err = asm.evaluate()
        orc.set_material_history(mh)
        rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_PLASTIC_NLNLOG, par, rm.conn, rm.X.ravel(), (fac * u).ravel(),
                                             rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind, tech=otech,
                                             hist=eas)
        orc.set_material_history(None)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, tol)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (step, wK, wf)
        assert mh[..., 7].max() > 1e-3  # yielding
        dh = asm.ctx.get_material_history().reshape(mh.shape)
        assert np.abs(dh - mh).max() <= 1e-11
        asm.ctx.update_step()
        orc.material_update(orc.MAT_PLASTIC_NLNLOG, mh)
        dh = asm.ctx.get_material_history().reshape(mh.shape)
        assert np.abs(dh - mh).max() <= 1e-11 and np.array_equal(dh[..., 7:], orc.plastic_new_history(rm.n_col_eles, ngp)[..., 7:])
    # restart: the state store round-trips through get / set (caller's element numbering)
    blob = asm.ctx.get_material_history()
    asm2 = b200.DeviceAssembler(rm, capi.MAT_PLASTIC_NLNLOG, PLASTIC10, tech=dtech)
    asm2.ctx.set_time_step_size(dt)
    asm2.ctx.set_material_history(blob)
    if tech != "none":
        asm2.ctx.set_element_history(asm.ctx.get_element_history())
    for a in (asm, asm2):
        a.set_state("displacement", (1.7 * u).ravel())
    K1, f1, _ = asm.evaluate()
    K2, f2, _ = asm2.evaluate()
    assert np.abs(K1 - K2).max() <= 1e-12 * np.abs(K1).max() and np.abs(f1 - f2).max() <= 1e-12 * np.abs(f1).max()
    asm.close()
    asm2.close()


def test_reference_plastic_necking_known_answer_through_cuda_path():
    """tests/input_files/plastic_necking_eas.4C.yaml (96 EAS-mild hex8, MAT_Struct_PlasticNlnLogNeoHooke, 30
    displacement-controlled steps into the necking regime, PREDICT ConstVel) with the CUDA path doing every evaluate,
    struct_calc_recover and the state commit: RESULT DESCRIPTION displacements."""
    import json
    import os
    from tests.helpers.newton import solve_prescribed_displacement_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "integration_hex8_plastic_necking_eas.json")))
    X = np.array(case["coords"])
    conn = np.array(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    m = case["material"]
    in10 = [m[k] for k in ("YOUNG", "NUE", "DENS", "YIELD", "ISOHARD", "SATHARDENING", "HARDEXPO", "VISC", "RATE_DEPENDENCY", "TOL")]
    asm = b200.DeviceAssembler(rm, capi.MAT_PLASTIC_NLNLOG, in10, tech=capi.TECH_EAS_MILD)
    asm.ctx.set_time_step_size(case["timestep"])
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_prescribed_displacement_case(case, evaluate, recover=lambda du: asm.recover(du),
                                           update_step=lambda u: asm.ctx.update_step(), predictor="constvel", max_iter=40)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-9, (e, u[e["node"], e["dof"]])
    asm.close()


# ---- row-element mode: export of the off-rank rows (FE_MATRIX; 4C_linalg_sparsematrix.cpp:598-630,705-710) ----------
@pytest.mark.parametrize("case,world", [("hex8", 2), ("hex8", 4), ("tet10", 2)])
def test_row_export_on_one_gpu_matches_owner_computes(case, world):
    """All ranks of a partition as contexts on ONE GPU: every rank evaluates its row elements only (element kernels write
    own + export rows), the export rows are carried by hand (b200s_add_exported = the Add of the export without a
    communicator) and every rank's own rows equal the oracle's owner-computes assembly over its column elements.  The
    NCCL flavour of the same exchange: tests/test_gpu_multirank.py."""
    fe = b200.fe_export
    if case == "hex8":
        ct, mk, par = orc.HEX8, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]
        mesh_of = lambda r: b200.partition.box_partition(b200.mesh.HEX8, (9, 7, 6), r, world)
    else:
        X, conn = b200.mesh.unstructured_tet10(4)
        owner = (X[:, 0] > 0.5).astype(int)
        ct, mk, par = orc.TET10, capi.MAT_ISO_VOL, [1.0, 1000.0]
        mesh_of = lambda r: b200.partition.generic_partition(b200.mesh.TET10, X, conn, owner, r, world)
    graphs = {}
    asms = []
    for r in range(world):
        view, rowptr, colind, plan = fe.fe_plan(mesh_of, r, world, graphs)
        asm = b200.RowExportAssembler(view, rowptr, colind, plan, mk, par)
        rm = view.rm
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.1)
        # one GPU, no communicator: the ghost displacements are written into the column state by hand, i.e. the
        # state vector handed in is the exact column state (row part) and the halo lists stay unused
        asm.ctx.set_halo([], [], [])
        asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
        asms.append((asm, view, plan, rowptr, colind, u))
    # ghost part of the column state: the library imports it through the halo; without a communicator the test writes
    # it through the device pointer of u_col
    from cuda.bindings import runtime as cudart
    for asm, view, plan, rowptr, colind, u in asms:
        ptr = asm.ctx.device_ptrs()["u_col"]
        buf = np.ascontiguousarray(u, np.float64)
        err, = cudart.cudaMemcpy(ptr, buf.ctypes.data, buf.nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
        assert int(err) == 0, err
    for asm, view, plan, rowptr, colind, u in asms:
        asm.evaluate("struct_calc_nlnstiff")
    n_moved = 0
    for r, (asm, view, plan, rowptr, colind, u) in enumerate(asms):
        for k, p in enumerate(plan.peers):
            src, _, splan = asms[p][0], asms[p][1], asms[p][2]
            vals, f = src.exported(splan.peers.index(r))
            assert len(vals) == len(plan.recv_pos[k]) and len(f) == len(plan.recv_fpos[k])
            if len(vals) or len(f):
                asm.ctx.add_exported(k, vals, f)
                n_moved += len(vals)
    assert n_moved > 0
    for r, (asm, view, plan, rowptr, colind, u) in enumerate(asms):
        rm = view.rm
        K, f = asm.own_rows()
        o_rowptr, o_colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        assert np.array_equal(rowptr[:plan.n_rows_owned + 1], o_rowptr) and np.array_equal(colind[:plan.nnz_owned], o_colind)
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                             rm.n_row_nodes, o_rowptr, o_colind, nthreads=4)
        ok, wK, wf = b200.parity_report(o_rowptr, K, rK, f, rf, RTOL)
        assert ok and nbad == 0, (r, wK, wf)
    # fewer elements evaluated than on the owner-computes path (its ghost layer is evaluated twice)
    assert sum(len(a[1].ele_ids) for a in asms) < sum(a[1].rm.n_col_eles for a in asms)
    for a in asms:
        a[0].close()


# ---- KINEM linear with TECH eas_mild / eas_full (EASFormulation<hex8, h8_9 | h8_21, linear>) -------------------------
@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
@pytest.mark.parametrize("mma", [1, 0])
def test_linear_kinematics_eas_matches_oracle(tech, mma):
    """Linear B-operator, enhanced strain B u + Mtilde alpha, deformation gradient = identity; K (with the geometric term
    the reference's EAS formulation keeps), f and the element history after recover against the oracle, on the tensor-core
    kernel and on the FMA kernel."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (5, 4, 3), 0, 1)
    u = 0.3 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    otech = {"eas_mild": orc.TECH_EAS_MILD_LINEAR, "eas_full": orc.TECH_EAS_FULL_LINEAR}[tech]
    dtech = {"eas_mild": capi.TECH_EAS_MILD, "eas_full": capi.TECH_EAS_FULL}[tech]
    mk, par = capi.MAT_SVK, [1000.0, 0.3]
    asm = b200.DeviceAssembler(rm, mk, par, tech=dtech, kinem=capi.KINEM_LINEAR, options={"eas_mma": mma})
    rowptr, colind = asm.graph()
    hist = orc.eas_new_history(otech, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    rng = np.random.default_rng(3)
    for it in range(2):
        us = (1.0 + 0.5 * it) * u.ravel()
        asm.set_state("displacement", us)
        K, f, err = asm.evaluate()
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, rm.X.ravel(), us, *maps, tech=otech, hist=hist)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-11)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (it, wK, wf)
        du = 1e-2 * rng.standard_normal(asm.n_rows)
        asm.recover(du)
        orc.eas_recover(otech, rm.conn, du, hist)
        dh = asm.ctx.get_element_history().reshape(hist.shape)
        neas = 9 if tech == "eas_mild" else 21
        assert np.abs(dh[:, :2 * neas] - hist[:, :2 * neas]).max() <= 1e-9 * max(np.abs(hist[:, :2 * neas]).max(), 1e-6)
    asm.close()


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
def test_reference_linear_eas_cooks_known_answer_through_cuda_path(tech):
    """tests/input_files/sohex8_easmild_cooks_lin.4C.yaml / sohex8_easfull_cooks_lin.4C.yaml with the CUDA path doing every
    evaluate and struct_calc_recover: dispy of node 162."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_cooks_{tech}_linear.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    dtech = {"eas_mild": capi.TECH_EAS_MILD, "eas_full": capi.TECH_EAS_FULL}[tech]
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [case["material"]["YOUNG"], case["material"]["NUE"]], tech=dtech,
                               kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, ev, recover=lambda du: asm.recover(du, 1.0))
    (e,) = case["expected_disp"]
    assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-8, (e, u[e["node"], e["dof"]])
    asm.close()


# ---- TECH shell_ans on hex8 (ShellANSFormulation, general kernel) ------------------------------------------------------
@pytest.mark.parametrize("mk", list(GEN_MATS))
def test_shell_ans_matches_oracle(mk):
    rm = b200.partition.box_partition(b200.mesh.HEX8, (5, 4, 3), 0, 1)
    rng = np.random.default_rng(4)
    inner = np.all((rm.X > 1e-9) & (rm.X < 1 - 1e-9), axis=1)
    rm.X[inner] += 0.03 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    asm = b200.DeviceAssembler(rm, mk, GEN_MATS[mk], tech=capi.TECH_SHELL_ANS)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, GEN_MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind, nthreads=4, tech=orc.TECH_SHELL_ANS)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0 and err.n_neg_jacobian == 0, (wK, wf)
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()
    # refused where the device has no kernel: wedge6 solid shell, shell_ans with KINEM linear
    X, conn = b200.mesh.structured_wedge6((2, 2, 2))
    ctx = capi.Context(0, 0, 1)
    ctx.set_mesh(b200.mesh.WEDGE6, X, conn)
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_element_properties(tech=capi.TECH_SHELL_ANS)
    assert ei.value.code == capi.ERR_UNSUPPORTED
    ctx.close()


def test_reference_shell_ans_known_answers_through_cuda_path():
    """RESULT DESCRIPTION displacements of tests/input_files/solid_ele_hex8_Standard_shell_ans.4C.yaml."""
    import json
    import os
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "integration_hex8_shell_ans.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [case["material"]["YOUNG"], case["material"]["NUE"]],
                               tech=capi.TECH_SHELL_ANS)
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, ev)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-11, (e, u[e["node"], e["dof"]])
    asm.close()


# ---- TECH shell_eas on hex8: EAS with the seven solid-shell modes (eastype_sh8_7), the FMA EAS kernel ------------------
def test_shell_eas_matches_oracle_and_known_answers():
    import json
    import os
    from tests.helpers.newton import solve_case
    rm = b200.partition.box_partition(b200.mesh.HEX8, (5, 4, 3), 0, 1)
    u = 0.5 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    mk, par = capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]
    asm = b200.DeviceAssembler(rm, mk, par, tech=capi.TECH_SHELL_EAS)
    rowptr, colind = asm.graph()
    hist = orc.eas_new_history(orc.TECH_SHELL_EAS, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    rng = np.random.default_rng(5)
    for it in range(2):
        us = (1.0 + 0.5 * it) * u.ravel()
        asm.set_state("displacement", us)
        K, f, err = asm.evaluate()
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, rm.X.ravel(), us, *maps, tech=orc.TECH_SHELL_EAS, hist=hist)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-10)
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (it, wK, wf)
        du = 1e-2 * rng.standard_normal(asm.n_rows)
        asm.recover(du)
        orc.eas_recover(orc.TECH_SHELL_EAS, rm.conn, du, hist)
        dh = asm.ctx.get_element_history().reshape(hist.shape)
        assert np.abs(dh[:, :14] - hist[:, :14]).max() <= 1e-8 * max(np.abs(hist[:, :14]).max(), 1e-6)
    asm.close()
    # the reference's known answers (solid_ele_hex8_Standard_shell_eas.4C.yaml) through the CUDA path
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "integration_hex8_shell_eas.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [case["material"]["YOUNG"], case["material"]["NUE"]], tech=capi.TECH_SHELL_EAS)
    rowptr, colind = asm.graph()

    def ev(uu):
        asm.set_state("displacement", uu)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    uu = solve_case(case, ev, recover=lambda du: asm.recover(du, 1.0))
    for e in case["expected_disp"]:
        assert abs(uu[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, uu[e["node"], e["dof"]])
    asm.close()


# ---- TECH shell_eas_ans on hex8: ANS + the seven EAS modes, condensed with the ANS B-operator (general kernel) ---------
def test_shell_eas_ans_matches_oracle_and_known_answers():
    import json
    import os
    from tests.helpers.newton import solve_case
    rm = b200.partition.box_partition(b200.mesh.HEX8, (5, 4, 3), 0, 1)
    u = 0.4 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    mk, par = capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]
    asm = b200.DeviceAssembler(rm, mk, par, tech=capi.TECH_SHELL_EAS_ANS)
    rowptr, colind = asm.graph()
    hist = orc.eas_new_history(orc.TECH_SHELL_EAS_ANS, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    rng = np.random.default_rng(5)
    for it in range(2):
        us = (1.0 + 0.5 * it) * u.ravel()
        asm.set_state("displacement", us)
        for action, oaction in (("struct_calc_nlnstiff", orc.ACTION_FORCE_STIFF), ("struct_calc_internalforce", orc.ACTION_FORCE)):
            K, f, err = asm.evaluate(action)
            rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, rm.X.ravel(), us, *maps, tech=orc.TECH_SHELL_EAS_ANS,
                                                 hist=hist, action=oaction)
            if oaction == orc.ACTION_FORCE:
                assert np.abs(f - rf).max() <= 1e-10 * np.abs(rf).max()
            else:
                ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-10)
                assert ok and nbad == 0 and err.n_neg_jacobian == 0, (it, wK, wf)
        du = 1e-3 * rng.standard_normal(asm.n_rows)
        asm.recover(du)
        orc.eas_recover(orc.TECH_SHELL_EAS_ANS, rm.conn, du, hist)
        dh = asm.ctx.get_element_history().reshape(hist.shape)
        assert np.abs(dh[:, :14] - hist[:, :14]).max() <= 1e-8 * max(np.abs(hist[:, :14]).max(), 1e-6)
        assert np.abs(hist[:, 7:14]).max() > 0
    asm.close()
    # the reference's known answers (solid_ele_hex8_Standard_shell_eas_ans.4C.yaml) through the CUDA path
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "integration_hex8_shell_eas_ans.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [case["material"]["YOUNG"], case["material"]["NUE"]],
                               tech=capi.TECH_SHELL_EAS_ANS)
    rowptr, colind = asm.graph()

    def ev(uu):
        asm.set_state("displacement", uu)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    uu = solve_case(case, ev, recover=lambda du: asm.recover(du, 1.0))
    for e in case["expected_disp"]:
        assert abs(uu[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, uu[e["node"], e["dof"]])
    asm.close()


# ---- struct_calc_stress of the three hex8 solid shells: the reference's nodal Cauchy stresses through the CUDA path -------
@pytest.mark.parametrize("tech", ["shell_ans", "shell_eas", "shell_eas_ans"])
def test_solid_shell_stress_output_known_answers_through_cuda_path(tech):
    import json
    import os
    from tests.helpers.newton import solve_case
    from tests.helpers.postprocess import nodal_stresses
    dtech, otech = {"shell_ans": (capi.TECH_SHELL_ANS, orc.TECH_SHELL_ANS), "shell_eas": (capi.TECH_SHELL_EAS, orc.TECH_SHELL_EAS),
                    "shell_eas_ans": (capi.TECH_SHELL_EAS_ANS, orc.TECH_SHELL_EAS_ANS)}[tech]
    case = json.load(open(os.path.join(os.path.dirname(__file__), "golden", f"integration_hex8_{tech}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, tech=dtech)
    rowptr, colind = asm.graph()

    def ev(uu):
        asm.set_state("displacement", uu)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, ev, recover=None if tech == "shell_ans" else (lambda du: asm.recover(du, 1.0)))
    asm.set_state("displacement", u.ravel())
    st, en = asm.calc_stress("cauchy", "gl")
    ns = nodal_stresses(orc.HEX8, conn, st, len(X))
    assert len(case["expected_stress"]) >= 24
    for e in case["expected_stress"]:
        assert abs(ns[e["node"], e["comp"]] - e["value"]) < 1e-8, (e, ns[e["node"], e["comp"]])
    # and against the oracle at the same state, Gauss point by Gauss point (2PK / GL and Cauchy / Euler-Almansi), energy
    hist = None if tech == "shell_ans" else asm.ctx.get_element_history().reshape(len(conn), -1).copy()
    for names, types in ((("2pk", "gl"), (orc.STRESS_2PK, orc.STRAIN_GL)), (("cauchy", "ea"), (orc.STRESS_CAUCHY, orc.STRAIN_EA))):
        st, en = asm.calc_stress(*names)
        ost, oen = orc.calc_stress(orc.HEX8, orc.MAT_COUP_NEOHOOKE, par, conn, X.ravel(), u.ravel(), *types, tech=otech, hist=hist)
        assert np.abs(st - ost).max() <= 1e-10 * np.abs(ost).max(), (tech, names)
        assert np.abs(en - oen).max() <= 1e-10 * np.abs(oen).max(), (tech, names)
    oe = orc.calc_energy(orc.HEX8, orc.MAT_COUP_NEOHOOKE, par, conn, X.ravel(), u.ravel(), tech=otech, hist=hist)
    assert abs(asm.calc_energy() - oe) <= 1e-10 * abs(oe)
    asm.close()


# ---- hex20 blocks through the hex27 kernel (shape tables padded to 27 nodes, padding masked in the scatter map) -----------------
@pytest.mark.parametrize("cell", ["hex20", "hex18"])
def test_hex20_fast_path_equals_general_kernel_and_oracle_on_partitioned_rows(cell):
    """Rank views of a 3-way node partition (ghost rows next to the padding rows), per-element material parameter, both
    actions: the hex27-kernel route of a hex20 block against the oracle and against the general kernel (option
    general_kernel = 1) on the same context set-up."""
    import ctypes
    import torch
    cudart = ctypes.CDLL("libcudart.so")
    X, conn = (b200.mesh.structured_hex20 if cell == "hex20" else b200.mesh.structured_hex18)((5, 3, 3))
    dcell, ocell = (b200.mesh.HEX20, orc.HEX20) if cell == "hex20" else (b200.mesh.HEX18, orc.HEX18)
    rng = np.random.default_rng(11)
    owner = (X[:, 0] * 2.999).astype(int).clip(0, 2)
    flip = rng.random(len(X)) < 0.05
    owner[flip] = rng.integers(0, 3, flip.sum())
    mk, par = capi.MAT_ISO_VOL, [1.0, 1000.0]
    for rank in range(3):
        rm = b200.partition.generic_partition(dcell, X, conn, owner, rank, 3)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.25)
        mue = 1.0 + 0.5 * rng.random(rm.n_col_eles)
        rm1 = b200.partition.RankMesh(**{**rm.__dict__, "nranks": 1, "rank": 0})
        res = {}
        for general in (0, 1):
            asm = b200.DeviceAssembler(rm1, mk, par, per_ele_param=mue, options={"general_kernel": general})
            rowptr, colind = asm.graph()
            ucol = torch.from_numpy(u.ravel().copy()).cuda()
            cudart.cudaMemcpy(ctypes.c_void_p(asm.ctx.device_ptrs()["u_col"]), ctypes.c_void_p(ucol.data_ptr()),
                              ctypes.c_size_t(ucol.numel() * 8), 3)
            asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
            K, f, err = asm.evaluate()
            _, f2, _ = asm.evaluate("struct_calc_internalforce")
            res[general] = (K, f, f2, err.min_detJ)
            asm.close()
        rK, rf, nbad = orc.evaluate_assemble(ocell, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                             rm.n_row_nodes, rowptr, colind, nthreads=4, mue_per_ele=mue)
        for general in (0, 1):
            K, f, f2, _ = res[general]
            ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
            assert ok and nbad == 0, (rank, general, wK, wf)
            assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
        assert abs(res[0][3] - res[1][3]) <= 1e-12 * abs(res[1][3])  # the same minimal Jacobian determinant from both routes


@pytest.mark.parametrize("cell", ["hex20", "hex18"])
def test_hex20_fast_path_reports_negative_nodal_jacobians_like_the_general_kernel(cell):
    """An element folded at one corner: det J <= 0 at a node (not at the Gauss points of its neighbours): both routes count
    the same elements and name the same first offender in the caller's numbering."""
    X, conn = (b200.mesh.structured_hex20 if cell == "hex20" else b200.mesh.structured_hex18)((3, 2, 2))
    X = X.copy()
    bad_ele = 7
    far = 6 if cell == "hex20" else 11   # the corner opposite to node 0 (hex18 numbers the top corners 9..12)
    X[conn[bad_ele, far]] = X[conn[bad_ele, 0]] - 0.05   # pulled through corner 0
    rm = b200.partition.serial_mesh(b200.mesh.HEX20 if cell == "hex20" else b200.mesh.HEX18, X, conn)
    out = {}
    for general in (0, 1):
        asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], options={"general_kernel": general})
        asm.set_state("displacement", np.zeros(asm.n_rows))
        rc, err = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
        out[general] = (rc, err.n_neg_jacobian, err.first_bad_element)
        asm.close()
    assert out[0][0] == capi.ERR_NEG_JACOBIAN and out[0] == out[1], out
    assert out[0][1] >= 1

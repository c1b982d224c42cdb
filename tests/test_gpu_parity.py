"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs, against the committed golden fixtures, and through size-independent properties at
BASELINE.json's full size.  Tolerance: 1e-12 relative per entry in the block-row norm of
SURVEY.md 8(d) (north_star); CSR pattern bit-exact."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers.newton import solve_case

pytestmark = pytest.mark.gpu
b200 = importlib.import_module("4c_b200")
capi = b200.capi
G = os.path.join(os.path.dirname(__file__), "golden")
MATS = {capi.MAT_COUP_NEOHOOKE: [200.0, 0.25], capi.MAT_ISO_VOL: [1.0, 1000.0], capi.MAT_SVK: [1000.0, 0.3]}
RTOL = 1e-12


def _oracle(rm, mk, par, u, rowptr, colind, action=orc.ACTION_FORCE_STIFF, per_ele=None):
    return orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                 rm.node_col_lid, rm.n_row_nodes, rowptr, colind, action=action,
                                 mue_per_ele=per_ele, nthreads=4)


@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (2, 1, 1), (7, 5, 3), (16, 12, 9)])
def test_hex8_matches_oracle(mk, n):
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    per_ele = None
    if mk == capi.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    asm = b200.DeviceAssembler(rm, mk, MATS[mk], per_ele_param=per_ele)
    rowptr, colind = asm.graph()
    o_rowptr, o_colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert rowptr.tobytes() == o_rowptr.tobytes() and colind.tobytes() == o_colind.tobytes()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    rK, rf, nbad = _oracle(rm, mk, MATS[mk], u, rowptr, colind, per_ele=per_ele)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    assert err.n_neg_jacobian == 0 and err.min_detJ > 0 and 0.5 < err.min_detF < 1.5
    # struct_calc_internalforce: force only, K untouched
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()


def test_hex8_reference_graph_through_set_graph():
    """The drop-in route: CSR graph handed over by the caller (here: the oracle's FillComplete
    restatement standing in for Epetra) instead of the device graph builder."""
    n = (5, 4, 3)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.2)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, MATS[capi.MAT_SVK], graph=(rowptr, colind))
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate()
    rK, rf, _ = _oracle(rm, capi.MAT_SVK, MATS[capi.MAT_SVK], u, rowptr, colind)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok, (wK, wf)
    asm.close()


def test_set_graph_rejects_non_blocked_layout():
    n = (2, 2, 2)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    ctx = capi.Context(0)
    ctx.set_mesh(capi.HEX8, rm.X, rm.conn)
    d = np.arange(3, dtype=np.int32)
    row = (3 * rm.node_row_lid[:, None] + d).ravel().astype(np.int32)
    col = (3 * rm.node_col_lid[:, None] + d).ravel().astype(np.int32)
    bad_col = col.copy()
    bad_col[1] = col[2]
    bad_col[2] = col[1]
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_graph(rowptr, colind, row, bad_col)
    assert ei.value.code == capi.ERR_UNSUPPORTED
    ctx.close()


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_hex8_partitioned_rows_match_oracle(nranks):
    """Every rank's view on ONE GPU (no halo exchange: the column state is supplied whole):
    owner-computes rows equal the serial oracle rows."""
    n = (8, 6, 4)
    mk = capi.MAT_COUP_NEOHOOKE
    for rank in range(nranks):
        rm = b200.partition.box_partition(b200.mesh.HEX8, n, rank, nranks)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / 8)
        rm1 = b200.partition.RankMesh(**{**rm.__dict__, "nranks": 1, "rank": 0})
        asm = b200.DeviceAssembler(rm1, mk, MATS[mk])
        rowptr, colind = asm.graph()
        o = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        assert rowptr.tobytes() == o[0].tobytes() and colind.tobytes() == o[1].tobytes()
        # ghost displacements go straight into the device column vector
        import torch
        ptrs = asm.ctx.device_ptrs()
        ucol = torch.from_numpy(u.ravel().copy()).cuda()
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaMemcpy(ctypes.c_void_p(ptrs["u_col"]), ctypes.c_void_p(ucol.data_ptr()),
                          ctypes.c_size_t(ucol.numel() * 8), 3)
        asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
        K, f, _ = asm.evaluate()
        rK, rf, _ = _oracle(rm, mk, MATS[mk], u, rowptr, colind)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (rank, wK, wf)
        asm.close()


def test_golden_hex8_integration_through_cuda_path():
    """RESULT DESCRIPTION displacements (1e-12) of the reference's hex8 ctest, with the CUDA path as
    the evaluate+assemble inside the Newton loop."""
    case = json.load(open(os.path.join(G, "integration_hex8.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par)
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


def test_negative_jacobian_error_code():
    n = (2, 2, 1)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    X = rm.X.copy()
    X[rm.conn[1, 6]] = X[rm.conn[1, 0]] - 0.2
    rm.X = X
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    asm.set_state("displacement", np.zeros(asm.n_rows))
    rc, err = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
    assert rc == capi.ERR_NEG_JACOBIAN and err.n_neg_jacobian >= 1 and err.min_detJ <= 0
    assert 0 <= err.first_bad_element < rm.n_col_eles
    asm.close()


def test_full_size_properties_1M_hex8():
    """BASELINE config 2 (100^3 hex8 Neo-Hooke): size-independent properties of f and K.
       * sum of internal forces vanishes per direction (self-equilibrated)
       * K * (rigid translation) = 0 row-wise
       * K symmetric: x^T K y = y^T K x for random x, y
       * a sub-block of rows equals the oracle evaluated on the elements touching those rows."""
    import torch
    n = (100, 100, 100)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.01)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    assert asm.nnz == 9 * 301 ** 3
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    assert err.n_neg_jacobian == 0 and 0.7 < err.min_detF < 1.3
    fs = np.abs(f).max()
    assert np.abs(f.reshape(-1, 3).sum(axis=0)).max() < 1e-9 * fs * np.sqrt(len(f))
    rows = np.repeat(np.arange(asm.n_rows), np.diff(rowptr))
    Kt = torch.sparse_csr_tensor(torch.from_numpy(rowptr), torch.from_numpy(colind.astype(np.int64)),
                                 torch.from_numpy(K), size=(asm.n_rows, asm.n_rows))
    scale = np.abs(K).max()
    for d in range(3):
        t = np.zeros(asm.n_rows)
        t[d::3] = 1.0
        r = (Kt @ torch.from_numpy(t)).numpy()
        assert np.abs(r).max() < 1e-10 * scale
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(asm.n_rows), rng.standard_normal(asm.n_rows)
    a = float(torch.from_numpy(x) @ (Kt @ torch.from_numpy(y)))
    bb = float(torch.from_numpy(y) @ (Kt @ torch.from_numpy(x)))
    assert abs(a - bb) < 1e-10 * max(abs(a), abs(bb), scale)
    del rows
    # oracle on a slab: elements with ez in [40,43) -> rows of nodes on planes 41 and 42 are complete
    ez = rm.ele_gid // (100 * 100)
    sel = (ez >= 40) & (ez < 43)
    conn_s = rm.conn[sel]
    iz = rm.node_gid // (101 * 101)
    keep = (iz == 41) | (iz == 42)
    sub_row = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    nsub = int(keep.sum())
    # sub-graph rows in the same column numbering
    o_rowptr, o_colind = orc.build_graph(orc.HEX8, conn_s, sub_row, rm.node_col_lid, nsub)
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], conn_s,
                                         rm.X.ravel(), u.ravel(), sub_row, rm.node_col_lid, nsub, o_rowptr,
                                         o_colind, nthreads=8)
    full_rows = (3 * np.nonzero(keep)[0][:, None] + np.arange(3)).ravel()
    lo, hi = rowptr[full_rows[0]], rowptr[full_rows[-1] + 1]
    assert np.array_equal(colind[lo:hi], o_colind) and hi - lo == len(rK)
    ok, wK, wf = b200.parity_report(o_rowptr, K[lo:hi], rK, f[full_rows], rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    asm.close()


# ---- tet10 (BASELINE config 4: unstructured, jittered, randomly renumbered; iso/vol Neo-Hooke) ----
@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [1, 3, 6])
def test_tet10_matches_oracle(mk, n):
    X, conn = b200.mesh.unstructured_tet10(n)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / n)
    per_ele = None
    if mk == capi.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    asm = b200.DeviceAssembler(rm, mk, MATS[mk], per_ele_param=per_ele)
    rowptr, colind = asm.graph()
    o_rowptr, o_colind = orc.build_graph(orc.TET10, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert rowptr.tobytes() == o_rowptr.tobytes() and colind.tobytes() == o_colind.tobytes()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    rK, rf, nbad = orc.evaluate_assemble(orc.TET10, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                         rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                         mue_per_ele=per_ele, nthreads=4)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    assert err.n_neg_jacobian == 0 and err.min_detJ > 0 and 0.5 < err.min_detF < 1.5
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()


def test_tet10_partitioned_rows_match_oracle():
    """3-way node partition of the unstructured tet10 mesh: each rank's owned rows (ghost state supplied
    whole) equal the oracle's owner-computes rows."""
    X, conn = b200.mesh.unstructured_tet10(4)
    rng = np.random.default_rng(5)
    owner = (X[:, 0] * 2.999).astype(int).clip(0, 2)          # slabs in x
    flip = rng.random(len(X)) < 0.05                           # ragged interfaces
    owner[flip] = rng.integers(0, 3, flip.sum())
    import ctypes
    import torch
    cudart = ctypes.CDLL("libcudart.so")
    for rank in range(3):
        rm = b200.partition.generic_partition(b200.mesh.TET10, X, conn, owner, rank, 3)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.25)
        rm1 = b200.partition.RankMesh(**{**rm.__dict__, "nranks": 1, "rank": 0})
        asm = b200.DeviceAssembler(rm1, capi.MAT_ISO_VOL, MATS[capi.MAT_ISO_VOL])
        rowptr, colind = asm.graph()
        o = orc.build_graph(orc.TET10, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        assert rowptr.tobytes() == o[0].tobytes() and colind.tobytes() == o[1].tobytes()
        ucol = torch.from_numpy(u.ravel().copy()).cuda()
        cudart.cudaMemcpy(ctypes.c_void_p(asm.ctx.device_ptrs()["u_col"]), ctypes.c_void_p(ucol.data_ptr()),
                          ctypes.c_size_t(ucol.numel() * 8), 3)
        asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
        K, f, _ = asm.evaluate()
        rK, rf, _ = orc.evaluate_assemble(orc.TET10, orc.MAT_ISO_VOL, MATS[capi.MAT_ISO_VOL], rm.conn,
                                          rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                          rm.n_row_nodes, rowptr, colind, nthreads=4)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (rank, wK, wf)
        asm.close()


def test_golden_tet10_integration_through_cuda_path():
    case = json.load(open(os.path.join(G, "integration_tet10.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    mat = case["material"]
    assert mat["kind"] == "ELAST_CoupNeoHooke"
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [mat["YOUNG"], mat["NUE"]])
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


# ---- hex27 (BASELINE config 3: structured cube, St.Venant-Kirchhoff) ----------------------------
@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (3, 2, 2), (6, 5, 4)])
def test_hex27_matches_oracle(mk, n):
    rm = b200.partition.box_partition(b200.mesh.HEX27, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.5 / max(n))
    per_ele = None
    if mk == capi.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    asm = b200.DeviceAssembler(rm, mk, MATS[mk], per_ele_param=per_ele)
    rowptr, colind = asm.graph()
    o_rowptr, o_colind = orc.build_graph(orc.HEX27, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert rowptr.tobytes() == o_rowptr.tobytes() and colind.tobytes() == o_colind.tobytes()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX27, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                         rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                         mue_per_ele=per_ele, nthreads=8)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    assert err.n_neg_jacobian == 0 and err.min_detJ > 0 and 0.5 < err.min_detF < 1.5
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()


@pytest.mark.parametrize("nranks", [2, 4])
def test_hex27_partitioned_rows_match_oracle(nranks):
    import ctypes
    import torch
    cudart = ctypes.CDLL("libcudart.so")
    n = (4, 3, 2)
    for rank in range(nranks):
        rm = b200.partition.box_partition(b200.mesh.HEX27, n, rank, nranks)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.1)
        rm1 = b200.partition.RankMesh(**{**rm.__dict__, "nranks": 1, "rank": 0})
        asm = b200.DeviceAssembler(rm1, capi.MAT_SVK, MATS[capi.MAT_SVK])
        rowptr, colind = asm.graph()
        o = orc.build_graph(orc.HEX27, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        assert rowptr.tobytes() == o[0].tobytes() and colind.tobytes() == o[1].tobytes()
        ucol = torch.from_numpy(u.ravel().copy()).cuda()
        cudart.cudaMemcpy(ctypes.c_void_p(asm.ctx.device_ptrs()["u_col"]), ctypes.c_void_p(ucol.data_ptr()),
                          ctypes.c_size_t(ucol.numel() * 8), 3)
        asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
        K, f, _ = asm.evaluate()
        rK, rf, _ = orc.evaluate_assemble(orc.HEX27, orc.MAT_SVK, MATS[capi.MAT_SVK], rm.conn, rm.X.ravel(),
                                          u.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                          colind, nthreads=4)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (rank, wK, wf)
        asm.close()


def test_golden_hex27_integration_through_cuda_path():
    case = json.load(open(os.path.join(G, "integration_hex27.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX27, X, conn)
    mat = case["material"]
    assert mat["kind"] == "ELAST_CoupNeoHooke"
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [mat["YOUNG"], mat["NUE"]])
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


def test_hex27_negative_jacobian_error_code():
    rm = b200.partition.box_partition(b200.mesh.HEX27, (2, 1, 1), 0, 1)
    X = rm.X.copy()
    X[rm.conn[1, 6]] = X[rm.conn[1, 0]] - 0.3   # corner pulled through the element
    rm.X = X
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    asm.set_state("displacement", np.zeros(asm.n_rows))
    rc, err = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
    assert rc == capi.ERR_NEG_JACOBIAN and err.n_neg_jacobian == 1 and err.first_bad_element == 1
    asm.close()


# ---- options of the C ABI ---------------------------------------------------------------------------
@pytest.mark.parametrize("opts", [{"reorder": 0}, {"double_buffer": 1}, {"reorder": 0, "double_buffer": 1}])
def test_options_do_not_change_results(opts):
    """Caller element order instead of the Morton order, and the two-buffer zero() overlap: same K, f
    (to rounding of the atomic sums), error element index reported in the caller's numbering."""
    n = (9, 7, 5)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / 9)
    ctx = capi.Context(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_mesh(capi.HEX8, rm.X, rm.conn)
    ctx.set_material(capi.MAT_COUP_NEOHOOKE, MATS[capi.MAT_COUP_NEOHOOKE])
    nnz, rowptr, colind = ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid, fetch=True)
    rK, rf, _ = _oracle(rm, capi.MAT_COUP_NEOHOOKE, MATS[capi.MAT_COUP_NEOHOOKE], u, rowptr, colind)
    K, f = np.empty(nnz), np.empty(3 * rm.n_row_nodes)
    for it in range(3):   # repeated evaluates exercise the buffer alternation
        ctx.evaluate(capi.ACTION_NLNSTIFF, np.ascontiguousarray(u.ravel()), K, f)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (opts, it, wK, wf)
    ctx.close()


def test_first_bad_element_is_in_caller_numbering_with_reordering():
    n = (6, 5, 4)
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    bad_e = 77
    X = rm.X.copy()
    X[rm.conn[bad_e, 6]] = X[rm.conn[bad_e, 0]] - 0.05
    # the moved node belongs to up to 8 elements; the smallest caller index among the inverted ones is reported
    rm.X = X
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    asm.set_state("displacement", np.zeros(asm.n_rows))
    rc, err = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, asm._u, None, None, check=False)
    nbad_ref = 0
    first = None
    for e in range(rm.n_col_eles):
        _, _, bad = orc.element(orc.HEX8, orc.MAT_SVK, [1.0, 0.1], rm.X[rm.conn[e]], np.zeros((8, 3)))
        if bad:
            nbad_ref += 1
            first = e if first is None else first
    assert rc == capi.ERR_NEG_JACOBIAN and err.n_neg_jacobian == nbad_ref and err.first_bad_element == first
    asm.close()


def test_check_finite_reports_nan():
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 3, 3), 0, 1)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    asm.ctx.set_option("check_finite", 1)
    u = np.zeros(asm.n_rows)
    u[5] = np.nan
    rc, err = asm.ctx.evaluate(capi.ACTION_NLNSTIFF, u, None, None, check=False)
    assert rc == capi.ERR_NONFINITE and err.n_nonfinite > 0
    asm.close()


# ---- size-independent properties at larger tet10 / hex27 sizes ---------------------------------------
def _properties(asm, rowptr, colind, K, f, tol=1e-9):
    import torch
    n = asm.n_rows
    Kt = torch.sparse_csr_tensor(torch.from_numpy(rowptr), torch.from_numpy(colind.astype(np.int64)),
                                 torch.from_numpy(K), size=(n, n))
    scale = np.abs(K).max()
    assert np.abs(f.reshape(-1, 3).sum(axis=0)).max() < tol * np.abs(f).max() * np.sqrt(n)   # sum f = 0
    for d in range(3):                                                                      # K t = 0
        t = np.zeros(n)
        t[d::3] = 1.0
        assert np.abs((Kt @ torch.from_numpy(t)).numpy()).max() < 1e-10 * scale
    rng = np.random.default_rng(0)                                                          # K = K^T
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    a = float(torch.from_numpy(x) @ (Kt @ torch.from_numpy(y)))
    b = float(torch.from_numpy(y) @ (Kt @ torch.from_numpy(x)))
    assert abs(a - b) < 1e-10 * max(abs(a), abs(b), scale)


def test_tet10_properties_and_sampled_rows_at_scale():
    """384k unstructured tet10 elements (config 4 family): properties + the rows of 200 random nodes
    against the oracle evaluated on the elements touching them."""
    X, conn = b200.mesh.unstructured_tet10(40)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / 40)
    par = MATS[capi.MAT_ISO_VOL]
    asm = b200.DeviceAssembler(rm, capi.MAT_ISO_VOL, par)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    assert err.n_neg_jacobian == 0 and 0.7 < err.min_detF < 1.3
    _properties(asm, rowptr, colind, K, f)
    rng = np.random.default_rng(11)
    nodes = np.sort(rng.choice(rm.n_row_nodes, 200, replace=False))
    keep = np.zeros(rm.n_col_nodes, bool)
    keep[nodes] = True
    sel = keep[rm.conn].any(axis=1)
    sub_row = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    o_rowptr, o_colind = orc.build_graph(orc.TET10, rm.conn[sel], sub_row, rm.node_col_lid, len(nodes))
    rK, rf, nbad = orc.evaluate_assemble(orc.TET10, orc.MAT_ISO_VOL, par, rm.conn[sel], rm.X.ravel(), u.ravel(),
                                         sub_row, rm.node_col_lid, len(nodes), o_rowptr, o_colind, nthreads=8)
    rows = (3 * nodes[:, None] + np.arange(3)).ravel()
    sl = np.concatenate([np.arange(rowptr[r], rowptr[r + 1]) for r in rows])
    assert np.array_equal(colind[sl], o_colind)
    ok, wK, wf = b200.parity_report(o_rowptr, K[sl], rK, f[rows], rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    asm.close()


def test_hex27_properties_and_slab_at_scale():
    """24^3 = 13.8k hex27 elements (config 3 family, nnz 6e7): properties + a slab of rows vs the oracle."""
    n = (24, 24, 24)
    rm = b200.partition.box_partition(b200.mesh.HEX27, n, 0, 1, box=(0.24, 0.24, 0.24))
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.01)
    par = MATS[capi.MAT_SVK]
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, par)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate()
    assert err.n_neg_jacobian == 0
    _properties(asm, rowptr, colind, K, f)
    ez = rm.ele_gid // (24 * 24)
    sel = (ez >= 10) & (ez < 12)
    iz = rm.node_gid // (49 * 49)
    keep = (iz >= 21) & (iz <= 23)          # node planes strictly inside the two element layers
    sub_row = np.where(keep, np.cumsum(keep) - 1, -1).astype(np.int32)
    nsub = int(keep.sum())
    o_rowptr, o_colind = orc.build_graph(orc.HEX27, rm.conn[sel], sub_row, rm.node_col_lid, nsub)
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX27, orc.MAT_SVK, par, rm.conn[sel], rm.X.ravel(), u.ravel(), sub_row,
                                         rm.node_col_lid, nsub, o_rowptr, o_colind, nthreads=8)
    full_rows = (3 * np.nonzero(keep)[0][:, None] + np.arange(3)).ravel()
    lo, hi = rowptr[full_rows[0]], rowptr[full_rows[-1] + 1]
    assert np.array_equal(colind[lo:hi], o_colind)
    ok, wK, wf = b200.parity_report(o_rowptr, K[lo:hi], rK, f[full_rows], rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    asm.close()


# ---- F-bar hex8 (TECH fbar) -----------------------------------------------------------------------
@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (7, 5, 3), (16, 12, 9)])
def test_hex8_fbar_matches_oracle(mk, n):
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, 0, 1)
    u = 3.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / max(n))
    per_ele = None
    if mk == capi.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    asm = b200.DeviceAssembler(rm, mk, MATS[mk], per_ele_param=per_ele, tech=capi.TECH_FBAR)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                         rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                         mue_per_ele=per_ele, nthreads=4, tech=orc.TECH_FBAR)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    assert err.n_neg_jacobian == 0 and err.min_detJ > 0
    _, f2, _ = asm.evaluate("struct_calc_internalforce")
    assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
    asm.close()


def test_golden_cooks_membrane_fbar_through_cuda_path():
    """tests/input_files/sohex8fbar_cooks_nl_new_struc.4C.yaml: 64 F-bar hex8, nu = 0.4999, 20 load
    steps; dispy of node 162 to the reference's 1e-8 with the CUDA path inside the Newton loop."""
    case = json.load(open(os.path.join(G, "integration_cooks_fbar.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(b200.mesh.HEX8, X, conn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, tech=capi.TECH_FBAR)
    rowptr, colind = asm.graph()

    def evaluate(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, evaluate)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], e
    asm.close()


def test_unsupported_element_properties_are_rejected():
    ctx = capi.Context(0)
    for kw in (dict(tech=7), dict(prestress=2), dict(kinem=5)):
        with pytest.raises(capi.B200SolidError) as ei:
            ctx.set_element_properties(**kw)
        assert ei.value.code == capi.ERR_UNSUPPORTED
    ctx.close()

"""CPU check of the PRODUCT's hex8 device code (phases, shared-memory indexing, scatter addressing)
against the oracle, by stepping the kernel's lanes on the host (tests/cpp/emulate_hex8.cpp)."""
import importlib

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import emulator as emu

b200 = importlib.import_module("4c_b200")
MATS = {orc.MAT_COUP_NEOHOOKE: [200.0, 0.25], orc.MAT_ISO_VOL: [1.0, 1000.0], orc.MAT_SVK: [1000.0, 0.3]}


def _case(n, rank=0, nranks=1):
    rm = b200.partition.box_partition(b200.mesh.HEX8, n, rank, nranks)
    h = 1.0 / max(n)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    return rm, u, rowptr, colind


@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (3, 2, 2), (5, 3, 2)])
def test_emulated_kernel_matches_oracle(mk, n):
    rm, u, rowptr, colind = _case(n)
    per_ele = None
    if mk == orc.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles  # MUE InputField per element
    ref_K, ref_f, nbad = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                               rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                               colind, mue_per_ele=per_ele)
    assert nbad == 0
    K, f, info = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, per_ele=per_ele)
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)
    assert info["n_bad"] == 0 and info["min_detJ"] > 0 and 0.5 < info["min_detF"] < 1.5


def test_emulated_force_only_action():
    rm, u, rowptr, colind = _case((3, 3, 2))
    _, ref_f, _ = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn,
                                        rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                        rm.n_row_nodes, rowptr, colind, action=orc.ACTION_FORCE)
    K, f, _ = emu.evaluate(orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(), rowptr,
                           colind, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, stiff=False)
    assert np.abs(f - ref_f).max() <= 1e-12 * np.abs(ref_f).max()
    assert not K.any()


@pytest.mark.parametrize("rank", [0, 1, 2, 3])
def test_emulated_kernel_owner_computes_rows(rank):
    """Partitioned view (4 ranks): ghost rows are skipped, owned rows equal the oracle's."""
    rm, u, rowptr, colind = _case((4, 4, 2), rank, 4)
    mk = orc.MAT_COUP_NEOHOOKE
    ref_K, ref_f, _ = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                            rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    K, f, _ = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                           rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)


def test_emulated_negative_jacobian_is_reported():
    rm, u, rowptr, colind = _case((2, 2, 1))
    X = rm.X.copy()
    X[rm.conn[1, 6]] = X[rm.conn[1, 0]] - 0.2  # fold element 1 inside out at one corner
    _, _, nbad = orc.evaluate_assemble(orc.HEX8, 0, [200.0, 0.25], rm.conn, X.ravel(), 0 * u.ravel(),
                                       rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    _, _, info = emu.evaluate(0, [200.0, 0.25], rm.conn, X.ravel(), 0 * u.ravel(), rowptr, colind,
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    assert nbad > 0 and info["n_bad"] == nbad and info["min_detJ"] <= 0


def test_hex8_corner_jacobian_shortcut_is_exact():
    """The kernel evaluates the reference's nodal Jacobian check (calc_lib.hpp:543-562) from the three
    edge vectors of the corner; it must equal the full sum over the 8 shape-function derivatives."""
    rng = np.random.default_rng(7)
    xi_nodes = orc.node_param_coords(orc.HEX8)
    for _ in range(20):
        X = xi_nodes * 0.5 + 0.2 * rng.standard_normal((8, 3))
        for g in range(8):
            _, dN = orc.shape(orc.HEX8, xi_nodes[g])        # dN[a, k]
            J_full = dN @ X                                  # J(a,b) = sum_k dN(a,k) X(k,b)
            sr = 0.5 if ((g ^ (g >> 1)) & 1) else -0.5
            ss = 0.5 if (g & 2) else -0.5
            st = 0.5 if (g & 4) else -0.5
            J_short = np.stack([sr * (X[g] - X[g ^ 1]), ss * (X[g] - X[g ^ 3]), st * (X[g] - X[g ^ 4])])
            assert np.array_equal(J_full, J_short)


def _tet_case(n, renumber=True):
    X, conn = b200.mesh.unstructured_tet10(n, renumber=renumber)
    rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 1.0 / n)
    rowptr, colind = orc.build_graph(orc.TET10, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    return rm, u, rowptr, colind


@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [1, 2])
def test_emulated_tet10_kernel_matches_oracle(mk, n):
    rm, u, rowptr, colind = _tet_case(n)
    per_ele = None
    if mk == orc.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    ref_K, ref_f, nbad = orc.evaluate_assemble(orc.TET10, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                               rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                               colind, mue_per_ele=per_ele)
    assert nbad == 0
    K, f, info = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, per_ele=per_ele, cell="tet10")
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)
    assert info["n_bad"] == 0 and info["min_detJ"] > 0 and 0.5 < info["min_detF"] < 1.5
    # force only
    _, f2, _ = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                            rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, stiff=False, per_ele=per_ele,
                            cell="tet10")
    assert np.abs(f2 - ref_f).max() <= 1e-12 * np.abs(ref_f).max()


@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (2, 2, 1)])
def test_emulated_hex27_kernel_matches_oracle(mk, n):
    rm = b200.partition.box_partition(b200.mesh.HEX27, n, 0, 1)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.5 / max(n))
    rowptr, colind = orc.build_graph(orc.HEX27, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    per_ele = None
    if mk == orc.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles)
    ref_K, ref_f, nbad = orc.evaluate_assemble(orc.HEX27, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                               rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                               colind, mue_per_ele=per_ele)
    assert nbad == 0
    K, f, info = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, per_ele=per_ele, cell="hex27")
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)
    assert info["n_bad"] == 0 and info["min_detJ"] > 0 and 0.5 < info["min_detF"] < 1.5
    _, f2, _ = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                            rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, stiff=False, per_ele=per_ele,
                            cell="hex27")
    assert np.abs(f2 - ref_f).max() <= 1e-12 * np.abs(ref_f).max()


def test_hex27_node_line_jacobian_is_exact():
    """hex27 nodal Jacobian from the three nodes of each parameter line == the full 27-term sum."""
    rng = np.random.default_rng(3)
    xi_nodes = orc.node_param_coords(orc.HEX27)
    X = xi_nodes * 0.5 + 0.05 * rng.standard_normal((27, 3))
    for n in range(27):
        _, dN = orc.shape(orc.HEX27, xi_nodes[n])
        nz = np.abs(dN) > 0
        # per direction a exactly the three nodes that share the other two coordinates with node n
        for a in range(3):
            others = [d for d in range(3) if d != a]
            line = np.all(xi_nodes[:, others] == xi_nodes[n, others], axis=1)
            assert np.array_equal(nz[a], line & (np.abs(dN[a]) > 0)) and not nz[a][~line].any()
        J = dN @ X
        assert np.linalg.det(J) > 0


# ---- F-bar hex8 (4c_b200/csrc/hex8_fbar_kernel.cuh) --------------------------------------------------
@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("n", [(1, 1, 1), (3, 2, 2), (5, 3, 2)])
def test_emulated_fbar_kernel_matches_oracle(mk, n):
    rm, u, rowptr, colind = _case(n)
    u = 3.0 * u  # larger strains: det F varies inside the element, fbar factor away from 1
    per_ele = None
    if mk == orc.MAT_ISO_VOL:
        per_ele = 1.0 + 0.1 * np.arange(rm.n_col_eles) / rm.n_col_eles
    ref_K, ref_f, nbad = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                               rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                               colind, mue_per_ele=per_ele, tech=orc.TECH_FBAR)
    assert nbad == 0
    K, f, info = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                              rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, per_ele=per_ele,
                              cell="hex8fbar")
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)
    assert info["n_bad"] == 0 and info["min_detJ"] > 0
    # the F-bar tangent differs from the displacement-based one and is not symmetric
    std_K, _, _ = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                        rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                        mue_per_ele=per_ele)
    assert np.abs(std_K - ref_K).max() > 1e-6 * np.abs(ref_K).max()
    # force only
    _, f2, _ = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                            rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, stiff=False, per_ele=per_ele,
                            cell="hex8fbar")
    assert np.abs(f2 - ref_f).max() <= 1e-12 * np.abs(ref_f).max()


@pytest.mark.parametrize("rank", [0, 3])
def test_emulated_fbar_owner_computes_rows(rank):
    rm, u, rowptr, colind = _case((4, 4, 2), rank, 4)
    mk = orc.MAT_COUP_NEOHOOKE
    ref_K, ref_f, _ = orc.evaluate_assemble(orc.HEX8, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(),
                                            rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                            tech=orc.TECH_FBAR)
    K, f, _ = emu.evaluate(mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                           rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, cell="hex8fbar")
    ok, rK, rf = b200.parity_report(rowptr, K, ref_K, f, ref_f, 1e-12)
    assert ok, (rK, rf)

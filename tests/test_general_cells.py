"""tet4, hex20, wedge6 and pyramid5 displacement-based blocks (create_solid_calculation_interface instantiates them next to hex8 / hex27 /
tet10, src/solid_ele/4C_solid_ele_factory.hpp:36-38,79-89): on the device they run through the table-driven general
kernel (generic_kernel.cuh).  Oracle pins: shape functions vs the reference's own headers, element matrices vs
oracle/_ref fixtures and the RESULT DESCRIPTION displacements of solid_ele_tet4_Standard_stressout.4C.yaml /
solid_ele_hex20_Standard_stressout.4C.yaml (tests/test_oracle_golden.py).  Here: the PRODUCT's tables and general-kernel
arithmetic stepped on the CPU against the oracle."""
import importlib

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import emulator as emu

b200 = importlib.import_module("4c_b200")
MATS = {orc.MAT_COUP_NEOHOOKE: [200.0, 0.25], orc.MAT_ISO_VOL: [1.0, 1000.0], orc.MAT_SVK: [1000.0, 0.3]}


def _mesh(cell):
    if cell == "tet4":
        X, conn = b200.mesh.unstructured_tet4(3)
        return orc.TET4, b200.partition.serial_mesh(b200.mesh.TET4, X, conn), 1.0 / 3
    if cell == "wedge6":
        X, conn = b200.mesh.structured_wedge6((3, 2, 2))
        return orc.WEDGE6, b200.partition.serial_mesh(b200.mesh.WEDGE6, X, conn), 1.0 / 3
    if cell == "pyramid5":
        X, conn = b200.mesh.structured_pyramid5((2, 2, 2))
        return orc.PYRAMID5, b200.partition.serial_mesh(b200.mesh.PYRAMID5, X, conn), 0.25
    if cell == "hex18":
        X, conn = b200.mesh.structured_hex18((2, 2, 2))
        return orc.HEX18, b200.partition.serial_mesh(b200.mesh.HEX18, X, conn), 0.5
    X, conn = b200.mesh.structured_hex20((2, 2, 1))
    return orc.HEX20, b200.partition.serial_mesh(b200.mesh.HEX20, X, conn), 0.5


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
@pytest.mark.parametrize("mk", list(MATS))
@pytest.mark.parametrize("linear", [False, True])
def test_emulated_general_kernel_matches_oracle(cell, mk, linear):
    ct, rm, h = _mesh(cell)
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    rK, rf, nbad = orc.evaluate_assemble(ct, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                         rm.n_row_nodes, rowptr, colind, tech=orc.TECH_LINEAR if linear else orc.TECH_NONE)
    K, f = emu.generic(rm.celltype, mk, MATS[mk], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind, rm.node_row_lid,
                       rm.node_col_lid, rm.n_row_nodes, linear=linear)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
    assert ok and nbad == 0, (wK, wf)


def test_general_kernel_equals_fast_kernels_on_their_cells():
    """The general kernel on hex8 / tet10 / hex27 gives what the oracle gives there too (same tables as the fast path)."""
    for cell, n in (("HEX8", (2, 2, 2)), ("TET10", 2), ("HEX27", (1, 1, 2))):
        ct = getattr(orc, cell)
        if cell == "TET10":
            X, conn = b200.mesh.unstructured_tet10(n)
            rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
        else:
            rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
        u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.4)
        rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(),
                                          rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
        K, f = emu.generic(rm.celltype, orc.MAT_COUP_NEOHOOKE, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind,
                           rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
        assert ok, (cell, wK, wf)


@pytest.mark.parametrize("cell", ["tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_emulated_general_kernel_mulf_matches_oracle(cell):
    """PRESTRESS_TECH mulf on the general cells (MulfFormulation is instantiated for every displacement-based cell,
    src/solid_ele/4C_solid_ele_factory.hpp:59-60; ctests solid_ele_{tet4,hex20,wedge6,pyramid5}_Standard_mulf): the general
    kernel with the MULF deformation gradient, the history set-up and struct_update_prestress of the product (stepped
    on the CPU) against the oracle over two prestress updates."""
    ct, rm, h = _mesh(cell)
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    par = [200.0, 0.25]
    ohist = orc.mulf_new_history(ct, rm.n_col_eles)
    phist = emu.mulf_history(rm.celltype, rm.conn, rm.X.ravel())
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    for scale in (1.0, -0.6, 0.3):
        us = scale * u.ravel()
        rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_COUP_NEOHOOKE, par, rm.conn, rm.X.ravel(), us, *maps, rowptr, colind,
                                             tech=orc.TECH_MULF, hist=ohist)
        K, f = emu.generic(rm.celltype, orc.MAT_COUP_NEOHOOKE, par, rm.conn, rm.X.ravel(), us, rowptr, colind, *maps,
                           gp_hist=phist)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
        assert ok and nbad == 0, (scale, wK, wf)
        orc.update_prestress(ct, rm.conn, rm.X.ravel(), us, ohist)
        emu.mulf_history(rm.celltype, rm.conn, rm.X.ravel(), us, phist)
        assert np.abs(phist - ohist[:, :, :18]).max() <= 1e-13 * np.abs(ohist).max()


@pytest.mark.parametrize("cell", ["hex8", "pyramid5"])
@pytest.mark.parametrize("mulf", [False, True])
def test_emulated_general_kernel_fbar_matches_oracle(cell, mulf):
    """TECH fbar (FbarEvaluators = hex8, pyramid5) and TECH fbar + PRESTRESS_TECH mulf (FBarMulfEvaluators, the same two cells:
    src/solid_ele/4C_solid_ele_factory.hpp:59-60,79-81) in the general kernel -- centroid mapping, F_bar, the three
    stiffness terms of the non-symmetric F-bar tangent, for MULF the Gauss-point and centroid histories and their
    update -- stepped on the CPU against the oracle."""
    if cell == "hex8":
        rm = b200.partition.box_partition(b200.mesh.HEX8, (2, 2, 2), 0, 1)
        ct, h = orc.HEX8, 0.4
    else:
        ct, rm, h = _mesh(cell)
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    for mk, par in ((orc.MAT_COUP_NEOHOOKE, [200.0, 0.25]), (orc.MAT_ISO_VOL, [1.0, 1000.0])):
        ohist = orc.mulf_new_history(ct, rm.n_col_eles, fbar=True) if mulf else None
        phist = emu.mulf_history(rm.celltype, rm.conn, rm.X.ravel(), fbar=True) if mulf else None
        for scale in (1.0, -0.6, 0.3):
            us = scale * u.ravel()
            rK, rf, nbad = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), us, *maps, rowptr, colind,
                                                 tech=orc.TECH_MULF_FBAR if mulf else orc.TECH_FBAR, hist=ohist)
            K, f = emu.generic(rm.celltype, mk, par, rm.conn, rm.X.ravel(), us, rowptr, colind, *maps, gp_hist=phist, fbar=True)
            ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
            assert ok and nbad == 0, (mk, scale, wK, wf)
            if mulf:
                orc.update_prestress(ct, rm.conn, rm.X.ravel(), us, ohist)
                emu.mulf_history(rm.celltype, rm.conn, rm.X.ravel(), us, phist, fbar=True)
                assert np.abs(phist - ohist[:, :, :18]).max() <= 1e-13 * np.abs(ohist).max()


def test_emulated_general_kernel_shell_eas_ans_matches_oracle():
    """TECH shell_eas_ans on hex8 (ShellEASANSFormulation<hex8, eastype_sh8_7>, 4C_solid_ele_calc_shell_eas_ans.hpp): the ANS block
    row plus the seven EAS modes (strain enhancement, K_aa / K_da / s with the ANS B-operator, condensation) -- the general
    kernel's arithmetic stepped on the CPU against the oracle (pinned to the reference's
    solid_ele_hex8_Standard_shell_eas_ans known answers, tests/test_post_actions.py), over two states with a recover of
    the EAS parameters in between so that alpha != 0 enters the second evaluate."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 2, 2), 0, 1)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    X = rm.X.copy()
    rng = np.random.default_rng(4)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.05 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    u = 0.4 * b200.mesh.displacement_field(X, rm.node_gid, 1.0, 0.4)
    for mk, par in MATS.items():
        ohist = orc.eas_new_history(orc.TECH_SHELL_EAS_ANS, rm.n_col_eles)
        phist = np.zeros_like(ohist)
        for it in range(2):
            us = (1.0 + 0.3 * it) * u.ravel()
            for stiff in (True, False):
                rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, X.ravel(), us, *maps, rowptr, colind,
                                                     tech=orc.TECH_SHELL_EAS_ANS, hist=ohist,
                                                     action=orc.ACTION_FORCE_STIFF if stiff else orc.ACTION_FORCE)
                K, f = emu.generic(rm.celltype, mk, par, rm.conn, X.ravel(), us, rowptr, colind, *maps, ans=True,
                                   eas_hist=phist, stiff=stiff)
                # the nearly incompressible block (kappa / mue = 1000) amplifies round-off once alpha != 0: 4.5e-11 observed,
                # 1e-15 for the other two materials on the same element arithmetic
                tol = 1e-9 if mk == orc.MAT_ISO_VOL else 1e-12
                ok, wK, wf = b200.parity_report(rowptr, K, rK if stiff else 0 * K, f, rf, tol)
                assert ok and nbad == 0, (mk, it, stiff, wK, wf)
            # s, K_aa^-1, K_da of the element history (what recover reads)
            assert np.abs(phist - ohist).max() <= 1e-9 * np.abs(ohist).max()
            du = 1e-3 * rng.standard_normal(3 * rm.n_col_nodes)
            orc.eas_recover(orc.TECH_SHELL_EAS_ANS, rm.conn, du, ohist)
            orc.eas_recover(orc.TECH_SHELL_EAS_ANS, rm.conn, du, phist)
            assert np.abs(ohist[:, 7:14]).max() > 0
            assert np.abs(phist[:, :14] - ohist[:, :14]).max() <= 1e-9 * np.abs(ohist[:, :14]).max()
            phist[:] = ohist  # the same alpha on both sides of the next comparison (the nearly incompressible block amplifies 1e-13)
        # and it is not the plain ANS element
        K1, f1 = emu.generic(rm.celltype, mk, par, rm.conn, X.ravel(), us, rowptr, colind, *maps, ans=True)
        assert np.abs(K - K1).max() > 1e-4 * np.abs(K1).max() or np.abs(f - f1).max() > 1e-4 * np.abs(f1).max()


def test_emulated_stress_output_shell_ans_matches_oracle():
    """struct_calc_stress / internal energy of a TECH shell_ans block: the product's post arithmetic (ANS strain at the
    Gauss point, deformation gradient consistent with it) stepped on the CPU against the oracle (pinned to the reference's
    nodal stresses of solid_ele_hex8_Standard_shell_ans.4C.yaml, tests/test_post_actions.py)."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 2, 2), 0, 1)
    X = rm.X.copy()
    rng = np.random.default_rng(4)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.05 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    u = 1.0 * b200.mesh.displacement_field(X, rm.node_gid, 1.0, 0.4)
    for mk, par in MATS.items():
        for stype, etype in ((orc.STRESS_2PK, orc.STRAIN_GL), (orc.STRESS_CAUCHY, orc.STRAIN_EA)):
            ost, oen = orc.calc_stress(orc.HEX8, mk, par, rm.conn, X.ravel(), u.ravel(), stype, etype, tech=orc.TECH_SHELL_ANS)
            st, en, energy = emu.post(rm.celltype, mk, par, rm.conn, X.ravel(), u.ravel(), stype, etype, fbar=3)
            assert np.abs(st - ost).max() <= 1e-11 * np.abs(ost).max(), (mk, stype)
            assert np.abs(en - oen).max() <= 1e-11 * np.abs(oen).max(), (mk, etype)
        oe = orc.calc_energy(orc.HEX8, mk, par, rm.conn, X.ravel(), u.ravel(), tech=orc.TECH_SHELL_ANS)
        assert abs(energy - oe) <= 1e-11 * abs(oe)
        # and it is not the displacement-based element's output
        st0, _, _ = emu.post(rm.celltype, mk, par, rm.conn, X.ravel(), u.ravel(), stype, etype)
        assert np.abs(st - st0).max() > 1e-4 * np.abs(st0).max()


def test_emulated_general_kernel_shell_ans_matches_oracle():
    """TECH shell_ans on hex8 (ShellANSFormulation, src/solid_ele/4C_solid_ele_calc_shell_ans.hpp): sampling-point
    interpolation of the transverse strain rows, T^-T formed from the inverse Jacobian, ANS geometric stiffness -- the
    product's general-kernel arithmetic stepped on the CPU against the oracle (pinned to the reference's
    solid_ele_hex8_Standard_shell_ans known answers, tests/test_post_actions.py)."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 2, 2), 0, 1)
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    # a distorted mesh: jitter the interior nodes so that the Jacobians differ between the Gauss and sampling points
    X = rm.X.copy()
    rng = np.random.default_rng(4)
    inner = np.all((X > 1e-9) & (X < 1 - 1e-9), axis=1)
    X[inner] += 0.05 * rng.uniform(-1, 1, (int(inner.sum()), 3))
    u = 2.0 * b200.mesh.displacement_field(X, rm.node_gid, 1.0, 0.4)
    for mk, par in MATS.items():
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, mk, par, rm.conn, X.ravel(), u.ravel(), *maps, rowptr, colind,
                                             tech=orc.TECH_SHELL_ANS)
        K, f = emu.generic(rm.celltype, mk, par, rm.conn, X.ravel(), u.ravel(), rowptr, colind, *maps, ans=True)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
        assert ok and nbad == 0, (mk, wK, wf)
        # and it is not the displacement-based element
        K0, f0 = emu.generic(rm.celltype, mk, par, rm.conn, X.ravel(), u.ravel(), rowptr, colind, *maps)
        assert np.abs(K - K0).max() > 1e-3 * np.abs(K0).max()


@pytest.mark.parametrize("cell", ["hex20", "hex18"])
def test_emulated_padded_hex27_route_matches_oracle(cell):
    """hex20 / hex18 blocks through the hex27 kernel (PAD instantiation): the kernel's steps on the CPU with the product's padded
    shape tables (pad_tables_to_hex27: zero derivatives for the padding nodes, zero-weight copies of Gauss point 0 for the
    padding points), the padded connectivity and the masked scatter map, on a rank view with ghost rows, against the oracle."""
    ct, rm0, h = _mesh(cell)
    X, conn = rm0.X, rm0.conn
    owner = (X[:, 0] > 0.45).astype(int)
    rm = b200.partition.generic_partition(rm0.celltype, X, conn, owner, 0, 2)
    assert 0 < rm.n_row_nodes < rm.n_col_nodes
    u = 2.0 * b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    for mk, par in MATS.items():
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), u.ravel(), *maps, rowptr, colind)
        K, f, info = emu.hex27_padded(rm.celltype, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind, *maps)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, 1e-12)
        assert ok and nbad == 0 and info["min_detJ"] > 0, (cell, mk, wK, wf)
        _, f2, _ = emu.hex27_padded(rm.celltype, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rowptr, colind, *maps, stiff=False)
        assert np.abs(f2 - rf).max() <= 1e-12 * np.abs(rf).max()

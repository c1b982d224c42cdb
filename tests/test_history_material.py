"""History material on the device GP-state store: MAT_Struct_DruckerPrager
(src/mat/4C_mat_plasticdruckerprager.cpp) with KINEM linear elements.

Oracle pins: every known answer of the reference's unit test unittests/mat/4C_druckerprager_test.cpp (elastic
step, return to the cone, return to the apex, a generic strain, history + unloading, tangent against finite
differences) and the tip displacements of its ctest inputs tests/input_files/mat_druckerprager_RetToCone.4C.yaml
(consistent, non-symmetric tangent) and ..._elastic_tang.4C.yaml.  The product's arithmetic (history_materials.cuh,
the KINEM linear hex8 kernel with the state store) is stepped on the CPU against the oracle; the GPU counterpart is
in tests/test_gpu_post.py."""
import json
import os

import numpy as np
import pytest

import importlib

from oracle import oracle as orc
from tests.helpers import emulator as emu
from tests.helpers.meshes import serial_maps
from tests.helpers.neumann import point_neumann_load
from tests.helpers.newton import solve_case

b200 = importlib.import_module("4c_b200")
G = os.path.join(os.path.dirname(__file__), "golden")
UNIT = dict(YOUNG=1.0, NUE=0.25, ISOHARD=1.0, TOL=1e-12, C=1.0, ETA=1.0, XI=1.0, ETABAR=1.0, MAXITER=50)
DP = orc.MAT_DRUCKER_PRAGER


def sym(e00, e11, e22, e01, e12, e02):
    return np.array([e00, e11, e22, e01, e12, e02], float)


def C_of(E):
    return 2.0 * E + np.array([1, 1, 1, 0, 0, 0.0])


STRAINS = {"elastic": sym(0.1, 0.1, 0.1, 0.05, 0.05, 0.05), "cone": sym(0, 0, 0, 1.1, 1.1, 1.1),
           "apex": sym(1, 1, 1, 0, 0, 0), "generic": sym(1.1, 2.0, 0.1, 1.25, 0.7, 0.5),
           "cone2": sym(0, 0.1, 0.2, 1.1, 2.2, 3.3)}


def test_unit_known_answers_of_the_reference():
    par = orc.drucker_prager_params(**UNIT)
    d = (1.0 / (1.25 * 0.5)) * (0.75 + 0.25 + 0.25) * .1
    o = (1.0 / (1.25 * 0.5)) * (0.5 / 2.0) * .1
    assert np.abs(orc.material_hist(DP, par, STRAINS["elastic"])[0] - [d, d, d, o, o, o]).max() < 1e-12
    Dg = (2.2 * np.sqrt(3) / 2.5 - 1.0) / 31.0 * 15.0
    dd = -(Dg * (1.0 / (3.0 * 0.5)))
    oo = (1.0 / 2.5) * (1 - ((1.0 / 2.5) * Dg / (2.2 * np.sqrt(3) / 2.5))) * 2.2
    assert np.abs(orc.material_hist(DP, par, STRAINS["cone"])[0] - [dd, dd, dd, oo, oo, oo]).max() < 1e-12
    a = 2.0 - (10. / 15.) * (3. / 5.)
    assert np.abs(orc.material_hist(DP, par, STRAINS["apex"])[0] - [a, a, a, 0, 0, 0]).max() < 1e-12
    ref = [1.3231031817668, 1.7934880206154, 0.8004533608238, 0.6533122761787, 0.3658548746601, 0.2613249104715]
    assert np.abs(orc.material_hist(DP, par, STRAINS["generic"])[0] - ref).max() < 1e-12
    # TestEvaluateHistory: plastic loading, update, back to zero strain: residual stress
    hist = np.zeros(orc.DP_HIST)
    orc.material_hist(DP, par, sym(1.1, 1.1, 1.1, 0.05, 0.05, 0.05), hist)
    assert np.all(hist[:7] == 0) and hist[13] > 0
    orc.material_update(DP, hist)
    assert np.array_equal(hist[:7], hist[7:])
    S, _ = orc.material_hist(DP, par, np.zeros(6), hist)
    assert np.linalg.norm(S) > 0.0


@pytest.mark.parametrize("name", list(STRAINS))
def test_tangent_against_finite_differences_like_the_reference(name):
    """fd_tangent of the unit test (delta 1e-7, tolerance 1e-5), for the associative parameters of the unit test
    and a non-associative set (ETA != ETABAR: the tangent loses its major symmetry)."""
    for over in ({}, dict(ETA=0.4, ETABAR=0.9, XI=0.7)):
        par = orc.drucker_prager_params(**dict(UNIT, **over))
        E = STRAINS[name]
        S0, Cm = orc.material_hist(DP, par, E)
        fd = np.zeros((6, 6))
        h = 1e-7
        for v in range(6):
            Ep = E.copy()
            Ep[v] += h
            fd[:, v] = (orc.material_hist(DP, par, Ep)[0] - S0) / (h if v < 3 else 2 * h)
        assert np.abs(fd - Cm).max() < 1e-5
        if over and name.startswith("cone"):
            assert np.abs(Cm - Cm.T).max() > 1e-3


@pytest.mark.parametrize("name", ["RetToCone", "RetToCone_elastic_tang"])
def test_ctest_known_answers(name):
    """Tip displacements (1e-10) of the reference's Drucker-Prager cantilever: 20 load steps, the history is
    committed after every step; 68 Gauss points yield."""
    case = json.load(open(os.path.join(G, f"integration_hex8_druckerprager_{name}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, row, col, nn)
    par = orc.drucker_prager_params(**case["material"])
    hist = np.zeros((len(conn), 8, orc.DP_HIST))

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(orc.HEX8, DP, par, conn, X.ravel(), u, row, col, nn, rowptr, colind,
                                              tech=orc.TECH_LINEAR, hist=hist)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_case(case, evaluate, max_iter=400, external=lambda u, t, v: (point_neumann_load(case, t), v),
                   update_step=lambda u: orc.material_update(DP, hist))
    assert int((hist[:, :, 6] > 0).sum()) == 68
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, u[e["node"], e["dof"]])


@pytest.mark.parametrize("name", list(STRAINS))
@pytest.mark.parametrize("tang", ["consistent", "elastic"])
def test_product_point_matches_oracle(name, tang):
    in10 = orc.drucker_prager_params(**dict(UNIT, TANG=tang))
    E = STRAINS[name]
    h_o, h_p = np.zeros(orc.DP_HIST), np.zeros(orc.DP_HIST)
    for step, fac in enumerate((1.0, 0.3, 1.4)):           # load, unload (elastic), reload beyond
        S, Cm = orc.material_hist(DP, in10, fac * E, h_o)
        S2, C2 = emu.drucker_prager(in10, C_of(fac * E), h_p)
        assert np.abs(S - S2).max() <= 1e-13 * max(np.abs(S).max(), 1e-3), (step, S, S2)
        assert np.abs(Cm - C2).max() <= 1e-13 * np.abs(Cm).max(), step
        assert np.abs(h_o - h_p).max() <= 1e-14 * max(np.abs(h_o).max(), 1.0)
        orc.material_update(DP, h_o)
        h_p[:7] = h_p[7:]
    # a virgin point without a history slot (padding lanes of the kernel) evaluates like zero history
    S3, _ = emu.drucker_prager(in10, C_of(E), None)
    assert np.abs(S3 - orc.material_hist(DP, in10, E)[0]).max() <= 1e-13 * np.abs(S3).max()


def test_product_unsymmetric_tangent_matches_oracle():
    """ETA != ETABAR with the consistent tangent: no major symmetry (the reference's RetToCone ctest input is of this
    kind).  The product evaluates such blocks with the general kernel's full 6x6 tangent."""
    in10 = orc.drucker_prager_params(**dict(UNIT, ETA=0.4))
    for name in STRAINS:
        E = STRAINS[name]
        h_o, h_p = np.zeros(orc.DP_HIST), np.zeros(orc.DP_HIST)
        S, Cm = orc.material_hist(DP, in10, E, h_o)
        S2, C2 = emu.drucker_prager(in10, C_of(E), h_p)
        assert np.abs(S - S2).max() <= 1e-13 * max(np.abs(S).max(), 1e-3)
        assert np.abs(Cm - C2).max() <= 1e-13 * np.abs(Cm).max(), name
        assert np.abs(h_o - h_p).max() <= 1e-14 * max(np.abs(h_o).max(), 1.0)
    Sc, Cc = emu.drucker_prager(in10, C_of(STRAINS["cone"]), None)
    assert np.abs(Cc - Cc.T).max() > 1e-3 * np.abs(Cc).max()          # really unsymmetric on the cone
    assert emu.drucker_prager(orc.drucker_prager_params(**dict(UNIT, ETA=0.4, TANG="elastic")), C_of(STRAINS["cone"])) is not None


@pytest.mark.parametrize("cell,n", [("HEX8", (3, 2, 2)), ("TET10", 2), ("HEX27", (2, 1, 1))])
def test_general_tangent_kernel_matches_oracle(cell, n):
    """generic_kernel.cuh stepped on the CPU: K (not symmetric), f and the state store for the non-associative
    consistent tangent, against the oracle, over load / commit / reload."""
    ct = getattr(orc, cell)
    if cell == "TET10":
        X, conn = b200.mesh.unstructured_tet10(n)
        rm = b200.partition.serial_mesh(b200.mesh.TET10, X, conn)
    else:
        rm = b200.partition.box_partition(getattr(b200.mesh, cell), n, 0, 1)
    u = np.asarray(b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.4), float).ravel()
    rowptr, colind = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    in10 = orc.drucker_prager_params(YOUNG=10.0, NUE=0.3, ISOHARD=0.5, TOL=1e-12, C=0.05, ETA=0.1, XI=0.25, ETABAR=0.25)
    ngp = orc.num_gp(ct)
    h_o = np.zeros((rm.n_col_eles, ngp, orc.DP_HIST))
    h_p = np.zeros_like(h_o)
    for fac in (1.0, 1.5):
        rK, rf, nbad = orc.evaluate_assemble(ct, DP, in10, rm.conn, rm.X.ravel(), fac * u, rm.node_row_lid, rm.node_col_lid,
                                             rm.n_row_nodes, rowptr, colind, tech=orc.TECH_LINEAR, hist=h_o)
        K, f = emu.generic_drucker_prager(rm.celltype, in10, rm.conn, rm.X.ravel(), fac * u, rowptr, colind, rm.node_row_lid,
                                          rm.node_col_lid, rm.n_row_nodes, h_p)
        assert nbad == 0
        assert np.abs(K - rK).max() <= 1e-12 * np.abs(rK).max()
        assert np.abs(f - rf).max() <= 1e-12 * np.abs(rf).max()
        assert np.abs(h_o - h_p).max() <= 1e-13 * np.abs(h_o).max()
        assert (h_o[:, :, 13] > 0).sum() > 4
        orc.material_update(DP, h_o)
        h_p[:, :, :7] = h_p[:, :, 7:]
    import scipy.sparse as sp
    A = sp.csr_matrix((rK, colind, rowptr), shape=(3 * rm.n_row_nodes,) * 2)
    assert abs(A - A.T).max() > 1e-6 * np.abs(rK).max()                 # the assembled K is not symmetric


def test_product_kernel_with_state_store_matches_oracle():
    """Emulated KINEM linear hex8 kernel: K, f and the Gauss-point history over load / commit / unload steps."""
    rm = b200.partition.box_partition(b200.mesh.HEX8, (3, 2, 2), 0, 1)
    u = np.asarray(b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.4), float).ravel()
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    in10 = orc.drucker_prager_params(YOUNG=10.0, NUE=0.3, ISOHARD=0.5, TOL=1e-12, C=0.05, ETA=0.2, XI=0.25, ETABAR=0.2)
    h_o = np.zeros((rm.n_col_eles, 8, orc.DP_HIST))
    h_p = np.zeros_like(h_o)
    for fac in (1.0, 0.6, 1.5):
        rK, rf, nbad = orc.evaluate_assemble(orc.HEX8, DP, in10, rm.conn, rm.X.ravel(), fac * u, rm.node_row_lid,
                                             rm.node_col_lid, rm.n_row_nodes, rowptr, colind, tech=orc.TECH_LINEAR,
                                             hist=h_o)
        K, f = emu.hex8_drucker_prager(in10, rm.conn, rm.X.ravel(), fac * u, rowptr, colind, rm.node_row_lid,
                                       rm.node_col_lid, rm.n_row_nodes, h_p)
        assert nbad == 0
        assert np.abs(K - rK).max() <= 1e-12 * np.abs(rK).max()
        assert np.abs(f - rf).max() <= 1e-12 * np.abs(rf).max()
        assert np.abs(h_o - h_p).max() <= 1e-13 * np.abs(h_o).max()
        assert (h_o[:, :, 13] > 0).sum() > 8     # the state is in the plastic range at many points
        orc.material_update(DP, h_o)
        h_p[:, :, :7] = h_p[:, :, 7:]

"""Pins the CPU oracle (oracle/solid_oracle.cpp) against every golden vector / known answer the
reference's own tests hold for this path and against oracle/_ref outputs."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers.meshes import serial_maps
from tests.helpers.newton import solve_case

G = os.path.join(os.path.dirname(__file__), "golden")
KA = json.load(open(os.path.join(G, "unit_known_answers.json")))
REFEL = np.load(os.path.join(G, "ref_elements.npz"))
CELLS = {"hex8": orc.HEX8, "hex27": orc.HEX27, "tet10": orc.TET10, "tet4": orc.TET4, "hex20": orc.HEX20,
         "wedge6": orc.WEDGE6, "pyramid5": orc.PYRAMID5, "hex18": orc.HEX18}


def _dp(a):
    import ctypes as C
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_calculate_gamma_delta_known_answers():
    k = KA["gamma_delta"]
    g, d = np.zeros(3), np.zeros(8)
    orc.lib().orc_calculate_gamma_delta(_dp(np.array(k["prinv"])), _dp(np.array(k["dPI"])),
                                        _dp(np.array(k["ddPII"])), _dp(g), _dp(d))
    assert np.abs(g - k["gamma_ref"]).max() < k["tol"]
    assert np.abs(d - k["delta_ref"]).max() < k["tol"]


def test_isoneohooke_invariant_derivatives_known_answers():
    k = KA["isoneohooke_invariant_derivatives"]
    dPI, ddPII = np.zeros(3), np.zeros(6)
    par = np.array([k["MUE"], 0.0])  # kappa = 0: IsoNeoHooke summand alone
    orc.lib().orc_invariant_derivatives(orc.MAT_ISO_VOL, _dp(par), _dp(np.array(k["prinv"])),
                                        _dp(dPI), _dp(ddPII))
    assert np.abs(dPI - k["dPI_ref"]).max() < k["tol"]
    assert np.abs(ddPII - k["ddPII_ref"]).max() < k["tol"]


def test_stvenantkirchhoff_known_answers():
    k = KA["stvenantkirchhoff"]
    par, E = np.array([k["YOUNG"], k["NUE"]]), np.array(k["E"])
    S, _ = orc.material(orc.MAT_SVK, par, E)
    y, nu = k["YOUNG"], k["NUE"]
    normal = (y / ((1 + nu) * (1 - 2 * nu))) * ((1 - nu) + nu + nu)
    shear = (y / ((1 + nu) * (1 - 2 * nu))) * ((1 - 2 * nu) / 2)
    assert np.abs(S[:3] - normal).max() < 1e-4
    assert np.abs(S[3:] - shear).max() < 1e-4
    psi = orc.lib().orc_svk_strain_energy(_dp(par), _dp(E))
    assert abs(psi - k["strain_energy_ref"]) < k["tol"]


def test_right_cauchy_green_from_gl_strain():
    # C = 2E + I is the first step of the oracle's ElastHyper evaluate; checked through S for a
    # material with S = gamma0*I + ... is indirect, so check the identity on the fixture directly.
    k = KA["right_cauchy_green"]
    C = 2 * np.array(k["E"]) + np.array([1, 1, 1, 0, 0, 0])
    assert np.abs(C - k["C_ref"]).max() < k["tol"]


def _unit_test_nodes(ct):
    """get_element_nodes<celltype>() of the reference unit test with the glibc rand() sequence."""
    seq = KA["bfree_stiffness"]["glibc_srand0_rand_over_RAND_MAX"]
    n = orc.num_nodes(ct)
    ref = orc.node_param_coords(ct)
    X, u = np.zeros((n, 3)), np.zeros((n, 3))
    q = 0
    for i in range(n):
        for d in range(3):
            X[i, d] = ref[i, d] + 0.03 * seq[q]
            u[i, d] = 0.02 * seq[q + 1]
            q += 2
    return X, u


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_bfree_equals_boperator_stiffness(cell):
    """unittests/solid_ele/4C_solid_ele_calc_lib_test.cpp:243-343 (BFreeElementStiffnessMatrix)."""
    ct = CELLS[cell]
    k = KA["bfree_stiffness"]
    X, u = _unit_test_nodes(ct)
    S = REFEL["unit_pk2"]
    Cm = REFEL["unit_cmat"]  # einsum_sym<"AC","BD"> + einsum_sym<"AD","BC"> from the reference
    # same cmat from its definition S_ac S_bd + S_ad S_bc (pins our Voigt layout too)
    P = np.array(k["pk2_full"])
    VA, VB = [0, 1, 2, 0, 1, 0], [0, 1, 2, 1, 2, 2]
    Cdef = np.array([[P[VA[I], VA[J]] * P[VB[I], VB[J]] + P[VA[I], VB[J]] * P[VB[I], VA[J]]
                      for J in range(6)] for I in range(6)])
    assert np.abs(Cdef - Cm).max() < 1e-14
    Kf = orc.stiffness_point(ct, X, u, k["xi"], S, Cm, 1.0, "bfree")
    Kb = orc.stiffness_point(ct, X, u, k["xi"], S, Cm, 1.0, "bop")
    assert np.abs(Kf - Kb).max() < k["tol"]
    if orc.ref_lib() is not None:
        Kr = orc.stiffness_point(ct, X, u, k["xi"], S, Cm, 1.0, "ref")
        assert np.abs(Kf - Kr).max() < k["tol"]


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10", "tet4", "hex20", "wedge6", "pyramid5", "hex18"])
@pytest.mark.parametrize("mk", [orc.MAT_COUP_NEOHOOKE, orc.MAT_ISO_VOL, orc.MAT_SVK])
def test_oracle_matches_reference_header_fixtures(cell, mk):
    """oracle == oracle/_ref outputs committed in tests/golden/ref_elements.npz."""
    ct = CELLS[cell]
    X, u, par = REFEL[f"{cell}_X"], REFEL[f"{cell}_u"], REFEL[f"m{mk}_par"]
    Ke, fe, err = orc.element(ct, mk, par, X, u)
    assert err == 0
    Kr, fr = REFEL[f"{cell}_m{mk}_K"], REFEL[f"{cell}_m{mk}_f"]
    assert np.abs(Ke - Kr).max() <= 1e-13 * np.abs(Kr).max()
    assert np.abs(fe - fr).max() <= 1e-13 * np.abs(fr).max()
    S, Cm = orc.material(mk, par, REFEL["mat_E"])
    assert np.abs(S - REFEL[f"mat{mk}_S"]).max() <= 1e-14 * np.abs(REFEL[f"mat{mk}_S"]).max()
    assert np.abs(Cm - REFEL[f"mat{mk}_C"]).max() <= 1e-14 * np.abs(REFEL[f"mat{mk}_C"]).max()


@pytest.mark.skipif(orc.ref_lib() is None, reason="oracle/_ref not built (no /root/reference)")
@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10", "tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_shape_functions_match_reference_headers(cell):
    import ctypes as C
    ct = CELLS[cell]
    n = orc.num_nodes(ct)
    rng = np.random.default_rng(3)
    pts = list(orc.gauss(ct)[0]) + list(rng.uniform(0.05, 0.3, (4, 3))) + list(orc.node_param_coords(ct))  # incl. the pyramid apex
    for xi in pts:
        xi = np.ascontiguousarray(xi)
        N, dN = orc.shape(ct, xi)
        Nr, dNr = np.zeros(n), np.zeros(3 * n)
        orc.ref_lib().ref_shape(ct, _dp(xi), _dp(Nr), _dp(dNr))
        assert np.abs(N - Nr).max() < 1e-15
        assert np.abs(dN - dNr.reshape(n, 3).T).max() < 4e-15
    ref = np.zeros(3 * n)
    orc.ref_lib().ref_node_param_coords(ct, _dp(ref))
    assert np.array_equal(ref.reshape(n, 3), orc.node_param_coords(ct))


def test_element_tangent_is_derivative_of_force():
    """K_e = d f_e / d u (central differences) for all cells/materials: ties f and K together."""
    rng = np.random.default_rng(5)
    for cell, ct in CELLS.items():
        for mk in (orc.MAT_COUP_NEOHOOKE, orc.MAT_ISO_VOL, orc.MAT_SVK):
            X, u, par = REFEL[f"{cell}_X"], REFEL[f"{cell}_u"], REFEL[f"m{mk}_par"]
            Ke, fe, _ = orc.element(ct, mk, par, X, u)
            h = 1e-6
            for col in rng.choice(3 * orc.num_nodes(ct), 4, replace=False):
                du = np.zeros(u.size)
                du[col] = h
                fp = orc.element(ct, mk, par, X, u + du.reshape(u.shape), orc.ACTION_FORCE)[1]
                fm = orc.element(ct, mk, par, X, u - du.reshape(u.shape), orc.ACTION_FORCE)[1]
                fd = (fp - fm) / (2 * h)
                assert np.abs(fd - Ke[:, col]).max() < 2e-6 * np.abs(Ke).max()
            assert np.abs(Ke - Ke.T).max() < 1e-12 * np.abs(Ke).max()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10", "tet4", "hex20", "wedge6", "pyramid5", "hex18"])
def test_integration_known_answers(cell):
    """RESULT DESCRIPTION displacements (1e-12) of the reference's ctest inputs, reproduced by the
    oracle's evaluate+assemble inside a Newton loop (tests/golden/integration_*.json)."""
    case = json.load(open(os.path.join(G, f"integration_{cell}.json")))
    ct = CELLS[cell]
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(ct, conn, row, col, nn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(ct, orc.MAT_COUP_NEOHOOKE, par, conn, X.ravel(), u,
                                              row, col, nn, rowptr, colind)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_case(case, evaluate)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], (e, u[e["node"], e["dof"]])


def test_cooks_membrane_fbar_known_answer():
    """tests/input_files/sohex8fbar_cooks_nl_new_struc.4C.yaml (64 F-bar hex8, CoupNeoHooke nu=0.4999,
    20 load steps): dispy of node 162 to the reference's 1e-8 -- pins the F-bar restatement."""
    case = json.load(open(os.path.join(G, "integration_cooks_fbar.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, row, col, nn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, par, conn, X.ravel(), u,
                                              row, col, nn, rowptr, colind, tech=orc.TECH_FBAR)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_case(case, evaluate)
    assert case["expected_disp"]
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], (e, u[e["node"], e["dof"]])


def test_fbar_tangent_is_derivative_of_force():
    rng = np.random.default_rng(11)
    X, u = REFEL["hex8_X"], 2.0 * REFEL["hex8_u"]
    for mk in (orc.MAT_COUP_NEOHOOKE, orc.MAT_ISO_VOL, orc.MAT_SVK):
        par = REFEL[f"m{mk}_par"]
        Ke, fe, _ = orc.element(orc.HEX8, mk, par, X, u, tech=orc.TECH_FBAR)
        h = 1e-6
        for col in rng.choice(24, 6, replace=False):
            du = np.zeros(24)
            du[col] = h
            fp = orc.element(orc.HEX8, mk, par, X, u + du.reshape(8, 3), orc.ACTION_FORCE, tech=orc.TECH_FBAR)[1]
            fm = orc.element(orc.HEX8, mk, par, X, u - du.reshape(8, 3), orc.ACTION_FORCE, tech=orc.TECH_FBAR)[1]
            assert np.abs((fp - fm) / (2 * h) - Ke[:, col]).max() < 2e-6 * np.abs(Ke).max()
        assert np.abs(Ke - Ke.T).max() > 1e-6 * np.abs(Ke).max()  # the F-bar tangent is not symmetric


def test_assembly_threads_equal_serial():
    """Owner-computes with emulated ranks (threads) gives bitwise the same rows as one rank."""
    from tests.helpers.meshes import serial_maps
    case = json.load(open(os.path.join(G, "integration_tet10.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.TET10, conn, row, col, nn)
    u = 0.01 * np.random.default_rng(1).uniform(-1, 1, 3 * nn)
    a = orc.evaluate_assemble(orc.TET10, orc.MAT_ISO_VOL, [1.0, 1000.0], conn, X.ravel(), u, row,
                              col, nn, rowptr, colind, nthreads=1)
    b = orc.evaluate_assemble(orc.TET10, orc.MAT_ISO_VOL, [1.0, 1000.0], conn, X.ravel(), u, row,
                              col, nn, rowptr, colind, nthreads=3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("cell", ["hex8", "hex20", "hex27", "tet4", "pyramid5"])
def test_linear_kinematics_known_answers(cell):
    """KINEM linear + MAT_Struct_StVenantKirchhoff: RESULT DESCRIPTION displacements (1e-12) of the reference's
    solid_ele_<cell>_Standard_linear.4C.yaml, reproduced by the oracle's small-strain path (tech = TECH_LINEAR)."""
    case = json.load(open(os.path.join(G, f"integration_{cell}_linear.json")))
    assert case["kinem"] == "linear" and case["material"]["kind"] == "MAT_Struct_StVenantKirchhoff"
    ct = CELLS[cell]
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(ct, conn, row, col, nn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(ct, orc.MAT_SVK, par, conn, X.ravel(), u, row, col, nn, rowptr, colind,
                                              tech=orc.TECH_LINEAR)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_case(case, evaluate)
    assert len(case["expected_disp"]) >= 12
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < e["tol"], (e, u[e["node"], e["dof"]])


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
def test_cooks_membrane_linear_eas_known_answer(tech):
    """tests/input_files/sohex8_easmild_cooks_lin.4C.yaml / sohex8_easfull_cooks_lin.4C.yaml (64 hex8, KINEM linear with
    TECH eas_mild / eas_full = EASFormulation<hex8, h8_9 | h8_21, linear>, St.Venant-Kirchhoff nu = 0.4999, one load
    step): dispy of node 162 -- pins the linear-kinematics EAS restatement (linear B-operator, enhanced strain
    B u + Mtilde alpha, deformation gradient = identity)."""
    case = json.load(open(os.path.join(G, f"integration_cooks_{tech}_linear.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, row, col, nn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    otech = {"eas_mild": orc.TECH_EAS_MILD_LINEAR, "eas_full": orc.TECH_EAS_FULL_LINEAR}[tech]
    hist = orc.eas_new_history(otech, len(conn))

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_SVK, par, conn, X.ravel(), u, row, col, nn, rowptr, colind,
                                              tech=otech, hist=hist)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_case(case, evaluate, recover=lambda du: orc.eas_recover(otech, conn, du, hist))
    (e,) = case["expected_disp"]
    assert abs(u[e["node"], e["dof"]] - e["value"]) < max(e["tol"], 1e-10), (e, u[e["node"], e["dof"]])

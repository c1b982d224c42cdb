"""MAT_ElastHyper summands formulated in (modified) principal stretches: ELAST_IsoOgden,
ELAST_CoupVarga, ELAST_IsoVarga through Mat::elast_hyper_add_response_stretches
(src/mat/4C_mat_elasthyper_service.cpp:208-386).

Oracle pins: the reference's unit test of the IsoOgden coefficients
(unittests/mat/elast/4C_elast_isoogden_test.cpp:44-68) and the nodal Cauchy stress known answers of its ctest input
tests/input_files/elasthyper_isoogden.4C.yaml.  The product's arithmetic (summands.cuh, eigen-dyad form of the
elasticity tensor, its own Jacobi eigen solver) is stepped on the CPU against the oracle; GPU counterpart in
tests/test_gpu_post.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import emulator as emu
from tests.helpers.meshes import serial_maps
from tests.helpers.newton import solve_prescribed_displacement_case
from tests.helpers.postprocess import nodal_stresses

G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = {"ELAST_IsoOgden": ("ISO_OGDEN", ["MUE", "ALPHA"]), "ELAST_VolOgden": ("VOL_OGDEN", ["KAPPA", "BETA"])}
BLOCKS = [
    [("ISO_OGDEN", [0.8, -25.0]), ("VOL_OGDEN", [1.0e3, 0.0])],
    [("COUP_VARGA", [2.0, 0.7])],
    [("ISO_VARGA", [2.0, 0.7]), ("VOL_SUSSMANBATHE", [50.0])],
    [("ISO_OGDEN", [0.8, 4.0]), ("COUP_VARGA", [1.0, 0.3]), ("ISO_YEOH", [0.6, 0.1, 0.05]), ("VOL_SUSSMANBATHE", [50.0])],
]
E_STATES = {
    "generic": np.array([0.11, -0.05, 0.07, 0.03, -0.02, 0.04]),
    "identity": np.zeros(6),                                          # three equal stretches
    "two-equal": np.array([0.2, 0.2, -0.1, 0.0, 0.0, 0.0]),           # lambda_1 == lambda_2 exactly
    "sheared two-equal": None,                                        # rotated version of the above
}


def case_material(case):
    return [(NAMES[s["name"]][0], [float(s["params"][k]) for k in NAMES[s["name"]][1]]) for s in case["summands"]]


def _state(name):
    if E_STATES[name] is not None:
        return E_STATES[name]
    # rotate diag(0.2, 0.2, -0.1) by a generic rotation: eigenvalues stay, eigenvectors are generic
    a, b = 0.7, -0.4
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    E = Rz @ Rx @ np.diag([0.2, 0.2, -0.1]) @ (Rz @ Rx).T
    return np.array([E[0, 0], E[1, 1], E[2, 2], E[0, 1], E[1, 2], E[0, 2]])


def _C_of(E):
    return 2.0 * E + np.array([1, 1, 1, 0, 0, 0.0])


def test_isoogden_coefficients_unit_known_answer():
    mg, md = np.zeros(3), np.zeros(6)
    p, ms = np.array([0.8, -25.0]), np.array([0.9, 0.8, 0.7])
    dp = C.POINTER(C.c_double)
    orc.lib().orc_test_stretch_coefficients_modified(orc.SUM["ISO_OGDEN"], p.ctypes.data_as(dp), ms.ctypes.data_as(dp),
                                                     mg.ctypes.data_as(dp), md.ctypes.data_as(dp))
    assert np.abs(mg - [-0.990546182, -21.175823681, -681.759084419]).max() < 1e-9
    assert np.abs(md - [28.615778605, 688.214269644, 25322.480278435, 0.0, 0.0, 0.0]).max() < 1e-9


def test_isoogden_ctest_known_answers():
    """Nodal Cauchy stresses (1e-8 absolute on values of 1e5) of elasthyper_isoogden.4C.yaml: compression,
    extension and shear of three cubes under prescribed displacements."""
    case = json.load(open(os.path.join(G, "integration_hex8_isoogden.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, row, col, nn)
    blk = orc.pack_elasthyper(case_material(case))

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_ELASTHYPER, blk, conn, X.ravel(), u, row, col, nn,
                                              rowptr, colind)
        assert nbad == 0
        return rowptr, colind, vals, f

    u = solve_prescribed_displacement_case(case, evaluate)
    st, _ = orc.calc_stress(orc.HEX8, orc.MAT_ELASTHYPER, blk, conn, X.ravel(), u.ravel(), orc.STRESS_CAUCHY,
                            orc.STRAIN_NONE)
    ns = nodal_stresses(orc.HEX8, conn, st, nn)
    assert len(case["expected_stress"]) == 3
    for e in case["expected_stress"]:
        assert abs(ns[e["node"], e["comp"]] - e["value"]) < 1e-8, (e, ns[e["node"], e["comp"]])


@pytest.mark.parametrize("bi", range(len(BLOCKS)))
def test_stretch_tangent_is_the_derivative_of_the_stress(bi):
    blk = orc.pack_elasthyper(BLOCKS[bi])
    E = E_STATES["generic"]
    S, Cm = orc.material(orc.MAT_ELASTHYPER, blk, E)
    assert np.abs(Cm - Cm.T).max() <= 1e-13 * np.abs(Cm).max()
    h = 1e-6
    w = np.array([1, 1, 1, 2, 2, 2.0])
    for v in range(6):
        d = np.zeros(6)
        d[v] = h
        fd = (orc.material(orc.MAT_ELASTHYPER, blk, E + d)[0] - orc.material(orc.MAT_ELASTHYPER, blk, E - d)[0]) / (2 * h)
        assert np.abs(fd / w[v] - Cm[:, v]).max() <= 2e-7 * np.abs(Cm).max()


@pytest.mark.parametrize("bi", range(len(BLOCKS)))
@pytest.mark.parametrize("state", list(E_STATES))
def test_product_stretch_summands_match_oracle(bi, state):
    """Generic state, the undeformed state (three equal stretches: limit branch of the reference, :346-349) and
    two equal stretches with aligned / generic principal axes."""
    blk = orc.pack_elasthyper(BLOCKS[bi])
    E = _state(state)
    S, Cm = orc.material(orc.MAT_ELASTHYPER, blk, E)
    S2, C2, _ = emu.material(3, blk, _C_of(E))
    # equal stretches: the result is independent of the eigenvectors chosen inside the eigenspace only up to
    # the O(1e-6) switch of the reference's limit formula; both sides take the same branch, round-off differs
    tol = 1e-12 if state == "generic" else 1e-9
    assert np.abs(S - S2).max() <= tol * max(np.abs(S).max(), np.abs(Cm).max() * 1e-3), (S, S2)
    assert np.abs(Cm - C2).max() <= tol * np.abs(Cm).max()


def test_stretch_free_blocks_are_unchanged_and_the_undeformed_state_is_stress_free():
    blk = orc.pack_elasthyper(BLOCKS[0])
    S, Cm = orc.material(orc.MAT_ELASTHYPER, blk, np.zeros(6))
    assert np.abs(S).max() <= 1e-12 * np.abs(Cm).max()
    # isotropic linearisation at the reference state, C = lambda 1(x)1 + 2 mu I: the shear modulus of the
    # reference's Ogden convention (psi = 2 mue / alpha^2 sum(ml^alpha - 1)) is MUE
    mu = 0.5 * (Cm[0, 0] - Cm[0, 1])
    assert abs(Cm[3, 3] - mu) <= 1e-9 * abs(mu)
    assert abs(mu - 0.8) <= 1e-9

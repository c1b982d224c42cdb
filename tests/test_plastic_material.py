"""MAT_Struct_PlasticNlnLogNeoHooke (src/mat/4C_mat_plasticnlnlogneohooke.cpp): finite-strain von Mises plasticity in
principal logarithmic strains, a history material ({accumulated plastic strain, Cp^-1} per Gauss point).

The reference has no unit test for it; its known answer is the ctest input plastic_necking_eas.4C.yaml (96 EAS-mild hex8,
30 displacement-controlled steps into the necking regime).  The oracle restatement (material + EAS + the ConstVel
predictor of the input) reproduces its RESULT DESCRIPTION displacements to 1e-13 -- that pins it.  The product evaluates
the same algorithm from C alone (Cholesky factor instead of F: the response is objective)."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import emulator as emu

b200 = importlib.import_module("4c_b200")
G = os.path.join(os.path.dirname(__file__), "golden")
MAT = dict(YOUNG=206.9, NUE=0.29, DENS=0.0, YIELD=0.45, ISOHARD=0.12924, SATHARDENING=0.715, HARDEXPO=16.93, VISC=1.0,
           RATE_DEPENDENCY=0.1, TOL=1e-12)
IN10 = [MAT[k] for k in ("YOUNG", "NUE", "DENS", "YIELD", "ISOHARD", "SATHARDENING", "HARDEXPO", "VISC", "RATE_DEPENDENCY", "TOL")]
DT = 0.175


def _F(seed, amp):
    return np.eye(3) + amp * np.random.default_rng(seed).standard_normal((3, 3))


def _voigt(T):
    return np.array([T[0, 0], T[1, 1], T[2, 2], T[0, 1], T[1, 2], T[0, 2]])


def _E(F):
    return _voigt(0.5 * (F.T @ F - np.eye(3)))


def test_oracle_tangent_is_derivative_of_stress_and_objective():
    par = orc.plastic_nlnlog_params(dt=DT, **MAT)
    for seed, amp in ((1, 0.02), (2, 0.05), (3, 1e-4)):
        F = _F(seed, amp)
        E0 = _E(F)
        h = orc.plastic_new_history()
        S, Cm = orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, E0, h)
        assert (h[7] > 0) == (amp > 1e-3)  # plastic for the two larger states, elastic for the small one
        num = np.zeros((6, 6))
        d = 1e-7
        for J in range(6):
            e = np.zeros(6)
            e[J] = d
            num[:, J] = (orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, E0 + e, orc.plastic_new_history())[0] -
                         orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, E0 - e, orc.plastic_new_history())[0]) / (2 * d)
        num[:, 3:] *= 0.5
        assert np.abs(num - Cm).max() <= 2e-8 * np.abs(Cm).max()
        assert np.abs(Cm - Cm.T).max() <= 1e-12 * np.abs(Cm).max()
        # the reference's F-based evaluation and the evaluation from C alone (any F with F^T F = C) agree
        S1, C1 = orc.material(orc.MAT_PLASTIC_NLNLOG, par, E0, F.T.ravel())
        S2, C2 = orc.material(orc.MAT_PLASTIC_NLNLOG, par, E0)
        # (stresses of small strains are differences of logarithms: absolute accuracy eps * modulus)
        assert np.abs(S1 - S2).max() <= 1e-13 * max(np.abs(S1).max(), MAT["YOUNG"]) and np.abs(C1 - C2).max() <= 1e-12 * np.abs(C1).max()


def test_oracle_elastic_limit_and_history_update():
    par = orc.plastic_nlnlog_params(dt=DT, **MAT)
    F = _F(5, 1e-6)
    S, _ = orc.material(orc.MAT_PLASTIC_NLNLOG, par, _E(F))
    mu, K = MAT["YOUNG"] / 2 / (1 + MAT["NUE"]), MAT["YOUNG"] / 3 / (1 - 2 * MAT["NUE"])
    E = 0.5 * (F.T @ F - np.eye(3))
    lin = _voigt((K - 2 * mu / 3) * np.trace(E) * np.eye(3) + 2 * mu * E)
    assert np.abs(S - lin).max() <= 1e-5 * np.abs(lin).max()
    # loading into the plastic range, update(), unloading back: elastic response around the new plastic state
    h = orc.plastic_new_history()
    F1 = _F(2, 0.05)
    orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, _E(F1), h)
    acc = h[7]
    assert acc > 0 and abs(np.linalg.det(np.array([[h[8], h[11], h[13]], [h[11], h[9], h[12]], [h[13], h[12], h[10]]])) - 1) < 1e-12
    orc.material_update(orc.MAT_PLASTIC_NLNLOG, h)
    assert h[0] == acc and np.allclose(h[7:], [0, 1, 1, 1, 0, 0, 0])
    # the same state again: the rate factor (VISC Dgamma / dt + 1)^RATE_DEPENDENCY had raised the yield stress of the
    # first step, holding the strain lets a little more plastic strain flow (visco-plastic relaxation)
    orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, _E(F1), h)
    assert 0.0 < h[7] - acc < 1e-3
    orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, _E(np.eye(3) + 0.99 * (F1 - np.eye(3))), h)
    assert h[7] == acc  # elastic unloading


@pytest.mark.parametrize("seed,amp", [(1, 0.02), (2, 0.05), (3, 1e-4), (4, 0.15)])
def test_product_point_matches_oracle(seed, amp):
    """The product's evaluation from C (Cholesky factor) against the oracle's F-based restatement: fresh point, then after
    update() with a second state (loading further), histories compared."""
    par = orc.plastic_nlnlog_params(dt=DT, **MAT)
    ho, hp = orc.plastic_new_history(), orc.plastic_new_history()
    for scale in (1.0, 1.4):
        F = np.eye(3) + scale * (_F(seed, amp) - np.eye(3))
        C = F.T @ F
        So, Co = orc.material_hist(orc.MAT_PLASTIC_NLNLOG, par, _E(F), ho)
        Sp, Cp = emu.plastic_nlnlog(IN10, DT, _voigt(C), hp)
        assert np.abs(Sp - So).max() <= 1e-12 * max(np.abs(So).max(), 0.01 * MAT["YOUNG"])
        assert np.abs(Cp - Co).max() <= 1e-10 * np.abs(Co).max()
        assert np.abs(hp - ho).max() <= 1e-12
        orc.material_update(orc.MAT_PLASTIC_NLNLOG, ho)
        orc.material_update(orc.MAT_PLASTIC_NLNLOG, hp.reshape(1, -1))


def test_undeformed_state_takes_the_equal_stretch_branch():
    par = orc.plastic_nlnlog_params(dt=DT, **MAT)
    So, Co = orc.material(orc.MAT_PLASTIC_NLNLOG, par, np.zeros(6))
    Sp, Cp = emu.plastic_nlnlog(IN10, DT, [1, 1, 1, 0, 0, 0])
    mu, K = MAT["YOUNG"] / 2 / (1 + MAT["NUE"]), MAT["YOUNG"] / 3 / (1 - 2 * MAT["NUE"])
    lam = K - 2 * mu / 3
    hooke = np.diag([2 * mu] * 3 + [mu] * 3)
    hooke[:3, :3] += lam
    assert np.abs(So).max() == 0 and np.abs(Sp).max() == 0
    assert np.abs(Co - hooke).max() <= 1e-12 * K and np.abs(Cp - hooke).max() <= 1e-12 * K


def test_reference_necking_known_answer():
    """tests/input_files/plastic_necking_eas.4C.yaml: RESULT DESCRIPTION displacements (tolerance 1e-12 in the reference)
    with the oracle's material + EAS restatements, displacement control, PREDICT ConstVel."""
    from tests.helpers.newton import solve_prescribed_displacement_case
    case = json.load(open(os.path.join(G, "integration_hex8_plastic_necking_eas.json")))
    X = np.array(case["coords"])
    conn = np.array(case["elements"], np.int32)
    ids = np.arange(len(X), dtype=np.int32)
    nn = len(X)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, ids, ids, nn)
    tech = orc.TECH_EAS_MILD
    par = orc.plastic_nlnlog_params(dt=case["timestep"], **case["material"])
    eas = orc.eas_new_history(tech, len(conn))
    mh = orc.plastic_new_history(len(conn), 8)

    def evaluate(u):
        orc.set_material_history(mh)
        K, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_PLASTIC_NLNLOG, par, conn, X.ravel(), u, ids, ids, nn, rowptr,
                                           colind, tech=tech, hist=eas)
        orc.set_material_history(None)
        assert nbad == 0
        return rowptr, colind, K, f

    u = solve_prescribed_displacement_case(case, evaluate, recover=lambda du: orc.eas_recover(tech, conn, du, eas),
                                           update_step=lambda u: orc.material_update(orc.MAT_PLASTIC_NLNLOG, mh),
                                           predictor="constvel", max_iter=40)
    assert mh[..., 0].max() > 0.5  # deep in the plastic range
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-12, (e, u[e["node"], e["dof"]])

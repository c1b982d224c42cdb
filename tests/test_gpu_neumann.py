"""GPU parity tests of the external loads through the C ABI: surface Neumann
conditions (Live, orthopressure with load linearisation, pseudo_orthopressure) and the volume load, against
the known-answer pinned restatement in tests/helpers/neumann.py (see tests/test_neumann.py)."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import neumann as nm
from tests.test_neumann import CELLS, _volume_gp_factors, mesh, synthetic_case, top

pytestmark = pytest.mark.gpu
b200 = importlib.import_module("4c_b200")
capi = b200.capi
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12
MAT = (capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_follower_pressure_and_load_stiffness_match_restatement(cell):
    rm, u = mesh(cell)
    case = synthetic_case(cell, rm, top)
    faces = nm.loaded_faces(case)
    asm = b200.DeviceAssembler(rm, *MAT)
    rowptr, colind = asm.graph()
    cond = asm.add_surface_neumann("orthopressure", faces, [1, 0, 0], case["neumann"]["VAL"])
    assert cond == 0
    asm.ctx.set_neumann_factors(cond, [nm.time_factor(case, 0.5), 1, 1])
    asm.set_state("displacement", u)
    K0, _, _ = asm.evaluate()
    ref_f, ref_K = nm.orthopressure_load(case, u, 0.5)
    refK = nm.add_to_csr_values(ref_K, rowptr, colind, np.zeros(len(colind)))
    # force only: the stiffness in the context is untouched
    f = asm.evaluate_neumann()
    assert np.abs(f - ref_f).max() <= RTOL * np.abs(ref_f).max()
    assert np.array_equal(asm.ctx.get_values(), K0)
    # with LOADLIN the load stiffness is added to the stiffness of this iteration
    f = asm.evaluate_neumann(loadlin=True)
    assert np.abs(f - ref_f).max() <= RTOL * np.abs(ref_f).max()
    dK = asm.ctx.get_values() - K0
    assert np.abs(refK).max() > 0
    assert np.abs(dK - refK).max() <= 1e-10 * np.abs(refK).max() + 1e-15 * np.abs(K0).max()
    # explicit state argument ("displacement" of the last converged step for pseudo_orthopressure)
    asm.ctx.clear_neumann()
    cond = asm.add_surface_neumann("pseudo_orthopressure", faces, [1, 0, 0], case["neumann"]["VAL"])
    u_old = 0.5 * u
    ref_old, _ = nm.orthopressure_load(case, u_old, 1.0, loadlin=False)
    asm.ctx.set_neumann_factors(cond, [nm.time_factor(case, 1.0), 1, 1])
    K1 = asm.ctx.get_values()
    f = asm.evaluate_neumann(u_row=u_old, loadlin=True)     # no linearisation for this type
    assert np.abs(f - ref_old).max() <= RTOL * np.abs(ref_old).max()
    assert np.array_equal(asm.ctx.get_values(), K1)
    # the resident displacement is not disturbed by the explicit state
    _, fint, _ = asm.evaluate()
    _, fint2, _ = asm.evaluate()
    assert np.abs(fint - fint2).max() <= 1e-13 * np.abs(fint).max()
    # MULF active: the orthopressure acts on the material configuration, not linearised
    asm.ctx.clear_neumann()
    cond = asm.add_surface_neumann("orthopressure", faces, [1, 0, 0], case["neumann"]["VAL"])
    ref_mat, _ = nm.orthopressure_load(case, u, 1.0, material=True)
    asm.ctx.set_neumann_factors(cond, [nm.time_factor(case, 1.0), 1, 1])
    K2 = asm.ctx.get_values()
    f = asm.evaluate_neumann(loadlin=True, mulf_active=True)
    assert np.abs(f - ref_mat).max() <= RTOL * np.abs(ref_mat).max()
    assert np.array_equal(asm.ctx.get_values(), K2)
    asm.close()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_live_load_with_host_evaluated_space_function(cell):
    """DESIGN SURF NEUMANN of the reference's integration inputs: FUNCT1 = t, FUNCT2 = (x+1)(y+2)(z+3)
    evaluated by the host at the Gauss-point positions b200s_neumann_gp_coords returns."""
    case = json.load(open(os.path.join(G, f"integration_{cell}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(CELLS[cell], X, conn)
    asm = b200.DeviceAssembler(rm, *MAT)
    faces = nm.loaded_faces(case)
    n = case["neumann"]
    cond = asm.add_surface_neumann("live", faces, n["ONOFF"][:3], n["VAL"][:3])
    gp = asm.ctx.neumann_gp_coords(cond)
    assert len(gp) == len(faces) * {"hex8": 4, "hex27": 9, "tet10": 6}[cell]
    t = 0.7
    fac = np.ones_like(gp)
    dof_fac = np.ones(3)
    for d in range(3):
        if n["FUNCT"][d] == 1:
            dof_fac[d] = t
        elif n["FUNCT"][d] == 2:
            fac[:, d] = (gp[:, 0] + 1) * (gp[:, 1] + 2) * (gp[:, 2] + 3)
    asm.ctx.set_neumann_factors(cond, dof_fac, fac)
    ref = nm.neumann_load(case, t)
    f = asm.evaluate_neumann()
    assert np.abs(ref).max() > 0
    assert np.abs(f - ref).max() <= RTOL * np.abs(ref).max()
    # factors are replaced, not accumulated; NULL resets them to 1
    asm.ctx.set_neumann_factors(cond, None, None)
    case1 = dict(case, neumann=dict(n, FUNCT=[0, 0, 0]))
    f1 = asm.evaluate_neumann()
    ref1 = nm.neumann_load(case1, t)
    assert np.abs(f1 - ref1).max() <= RTOL * np.abs(ref1).max()
    asm.close()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_volume_load_and_condition_sum(cell):
    rm, u = mesh(cell)
    X, conn = np.asarray(rm.X, float), np.asarray(rm.conn, np.int32)
    asm = b200.DeviceAssembler(rm, *MAT)
    vn = {"ONOFF": [1, 0, 1], "VAL": [2.0, 7.0, -0.3], "FUNCT": [2, 0, 1]}
    t = 0.25
    # every element of the block
    c_all = asm.add_volume_neumann(vn["ONOFF"], vn["VAL"])
    gp = asm.ctx.neumann_gp_coords(c_all)
    case = {"celltype": cell, "coords": X.tolist(), "elements": conn.tolist(), "vol_neumann": vn}
    fac = _volume_gp_factors(cell, case, conn, X, t)
    # the positions handed to the host are the reference positions of the stiffness rule's points, element-major
    # in the caller's element numbering
    gxi, _ = orc.gauss(CELLS[cell])
    for g, xi in enumerate(gxi):
        N, _ = orc.shape(CELLS[cell], xi)
        assert np.abs(gp.reshape(len(conn), len(gxi), 3)[:, g] - np.einsum("a,eak->ek", N, X[conn])).max() < 1e-14
    asm.ctx.set_neumann_factors(c_all, None, fac)
    ref = nm.volume_neumann_load(case, t)
    f = asm.evaluate_neumann()
    assert np.abs(f - ref).max() <= RTOL * np.abs(ref).max()
    # a second condition on a subset of the elements and a surface load: the conditions add up
    ids = np.array([len(conn) - 1, 0, 3], np.int32)
    sub = dict(case, elements=conn[ids].tolist())
    c_sub = asm.add_volume_neumann(vn["ONOFF"], vn["VAL"], ids)
    assert c_sub == 1
    asm.ctx.set_neumann_factors(c_sub, None, _volume_gp_factors(cell, sub, conn[ids], X, t))
    scase = synthetic_case(cell, rm, top)
    c_s = asm.add_surface_neumann("orthopressure", nm.loaded_faces(scase), [1, 0, 0], scase["neumann"]["VAL"])
    asm.ctx.set_neumann_factors(c_s, [nm.time_factor(scase, t), 1, 1])
    asm.set_state("displacement", u)
    ref_tot = ref + nm.volume_neumann_load(sub, t) + nm.orthopressure_load(scase, u, t, loadlin=False)[0]
    f = asm.evaluate_neumann()
    assert np.abs(f - ref_tot).max() <= RTOL * np.abs(ref_tot).max()
    asm.ctx.clear_neumann()
    assert np.all(asm.evaluate_neumann() == 0.0)
    asm.close()


def test_reference_follower_load_known_answer_through_cuda_path():
    """tests/input_files/loadlin.4C.yaml: dispy of node 404 = -22.839151032191484 (1e-8) with the device doing
    every evaluate (EAS full), recover, the follower pressure and its linearisation into the resident K."""
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(G, "integration_hex8_loadlin.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, tech=capi.TECH_EAS_FULL)
    rowptr, colind = asm.graph()
    cond = asm.add_surface_neumann("orthopressure", nm.loaded_faces(case), case["neumann"]["ONOFF"],
                                   case["neumann"]["VAL"])

    def ev(u):
        asm.set_state("displacement", u)
        _, f, _ = asm.evaluate_resident()
        return rowptr, colind, "resident", f

    def external(u, t, values):
        asm.ctx.set_neumann_factors(cond, [nm.time_factor(case, t), 1, 1])
        if values is None:                               # convergence check: force only
            return asm.evaluate_neumann(), None
        fext = asm.evaluate_neumann(loadlin=True)        # K_T = K_int + K_load in the context
        return fext, asm.ctx.get_values()

    u = solve_case(case, ev, recover=lambda du: asm.recover(du, 1.0), external=external)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-8, (e, u[e["node"], e["dof"]])
    asm.close()


def test_error_behaviour_of_the_load_interface():
    rm, u = mesh("hex8")
    case = synthetic_case("hex8", rm, top)
    faces = nm.loaded_faces(case)
    asm = b200.DeviceAssembler(rm, *MAT)
    # "orthopressure on 1st dof only!" (4C_solid_ele_surface_evaluate.cpp:322-324)
    for onoff in ([0, 1, 0], [1, 1, 0]):
        with pytest.raises(capi.B200SolidError) as ei:
            asm.add_surface_neumann("orthopressure", faces, onoff, [1.0, 0, 0])
        assert ei.value.code == capi.ERR_INVALID and "1st dof only" in str(ei.value)
    with pytest.raises(capi.B200SolidError) as ei:
        asm.ctx.add_surface_neumann(7, faces, [1, 0, 0], [1.0, 0, 0])
    assert ei.value.code == capi.ERR_UNSUPPORTED
    bad = faces.copy()
    bad[0, 0] = rm.n_col_nodes
    with pytest.raises(capi.B200SolidError) as ei:
        asm.add_surface_neumann("live", bad, [1, 1, 1], [1.0, 0, 0])
    assert ei.value.code == capi.ERR_INVALID
    with pytest.raises(capi.B200SolidError):
        asm.add_volume_neumann([1, 0, 0], [1.0, 0, 0], [rm.n_col_eles])
    with pytest.raises(capi.B200SolidError):
        asm.ctx.neumann_gp_coords(0)
    # a "face" whose nodes do not share an element: no entry in the stiffness graph for the linearisation
    X = np.asarray(rm.X)
    far = [int(np.argmin(np.abs(X - c).sum(axis=1))) for c in ([0, 0, 0], [1, 0, 0], [1, 1, 1], [0, 1, 1])]
    asm.add_surface_neumann("orthopressure", [far], [1, 0, 0], [1.0, 0, 0])
    asm.set_state("displacement", u)
    asm.evaluate_resident()
    with pytest.raises(capi.B200SolidError) as ei:
        asm.evaluate_neumann(loadlin=True)
    assert ei.value.code == capi.ERR_INVALID and "graph" in str(ei.value)
    asm.close()
    # before any evaluate / set_state the resident displacement is zero: load on the undeformed surface
    asm = b200.DeviceAssembler(rm, *MAT)
    asm.add_surface_neumann("orthopressure", faces, [1, 0, 0], [1.0, 0, 0])
    ref0, _ = nm.orthopressure_load(dict(case, neumann=dict(case["neumann"], VAL=[1.0, 0, 0], FUNCT=[0, 0, 0])),
                                    np.zeros_like(u), 0.0, loadlin=False)
    assert np.abs(asm.evaluate_neumann() - ref0).max() <= RTOL * np.abs(ref0).max()
    asm.close()

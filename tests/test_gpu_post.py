"""GPU parity tests of the actions around the hot path, through the C ABI against the
oracle: generic MAT_ElastHyper inside the evaluate+assemble kernels, struct_calc_stress,
struct_calc_energy, the mass actions, and the device-resident linear algebra (apply_dirichlet,
multiply, CG hand-off).  Tolerance 1e-12 in the norms stated per test."""
import importlib
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as orc
from tests.helpers.neumann import neumann_load

pytestmark = pytest.mark.gpu
b200 = importlib.import_module("4c_b200")
capi = b200.capi
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12
CELLS = {"hex8": orc.HEX8, "hex27": orc.HEX27, "tet10": orc.TET10}
HYPER = [("ISO_YEOH", [0.6, 0.1, 0.05]), ("ISO_MOONEYRIVLIN", [0.2, 0.1]), ("VOL_OGDEN", [300.0, -2.0]),
         ("COUP_1POW", [0.3, 2])]
MATS = [(capi.MAT_COUP_NEOHOOKE, [200.0, 0.25]), (capi.MAT_ISO_VOL, [1.0, 1000.0]), (capi.MAT_SVK, [1000.0, 0.3]),
        (capi.MAT_ELASTHYPER, HYPER)]


def _mesh(cell, n=(5, 4, 3), rank=0, nranks=1):
    ct = CELLS[cell]
    if cell == "tet10":
        X, conn = b200.mesh.unstructured_tet10(3, seed=3)
        rm = b200.partition.serial_mesh(ct, X, conn)
        h = 1.0 / 3
    else:
        rm = b200.partition.box_partition(ct, n, rank, nranks)
        h = 1.0 / max(n)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, h)
    return ct, rm, u


def _opar(mk, par):
    return orc.pack_elasthyper(par) if mk == capi.MAT_ELASTHYPER else par


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_generic_elasthyper_evaluate_matches_oracle(cell):
    ct, rm, u = _mesh(cell)
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, HYPER)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_ELASTHYPER, orc.pack_elasthyper(HYPER), rm.conn, rm.X.ravel(),
                                         u.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr,
                                         colind, nthreads=4)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    asm.close()


def test_generic_elasthyper_fbar_and_specialised_equivalence():
    """F-bar kernel with the summand list; and the list {CoupNeoHooke} gives the specialised kernel's K."""
    ct, rm, u = _mesh("hex8")
    one = [("COUP_NEOHOOKE", [200.0, 0.25])]
    a = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, one, tech=capi.TECH_FBAR)
    b = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], tech=capi.TECH_FBAR)
    rowptr, colind = a.graph()
    for asm in (a, b):
        asm.set_state("displacement", u.ravel())
    Ka, fa, _ = a.evaluate()
    Kb, fb, _ = b.evaluate()
    ok, wK, wf = b200.parity_report(rowptr, Ka, Kb, fa, fb, RTOL)
    assert ok, (wK, wf)
    a.close(), b.close()


def test_unknown_summand_is_rejected():
    ct, rm, u = _mesh("hex8", (2, 2, 2))
    ctx = capi.Context(0)
    ctx.set_mesh(ct, rm.X, rm.conn)
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_material_elasthyper([(99, [1.0])])
    assert ei.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_material_elasthyper([("COUP_1POW", [1.0, 1])])  # the reference throws for D < 2
    assert ei.value.code == capi.ERR_INVALID
    ctx.close()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
@pytest.mark.parametrize("mi", range(len(MATS)))
def test_calc_stress_and_energy_match_oracle(cell, mi):
    ct, rm, u = _mesh(cell)
    mk, par = MATS[mi]
    rm.ele_is_row = (np.arange(rm.n_col_eles) % 3 != 0).astype(np.uint8)
    asm = b200.DeviceAssembler(rm, mk, par)
    asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
    for io, types in ((("2pk", "gl"), (orc.STRESS_2PK, orc.STRAIN_GL)),
                      (("cauchy", "ea"), (orc.STRESS_CAUCHY, orc.STRAIN_EA))):
        st, en = asm.calc_stress(*io)
        rst, ren = orc.calc_stress(ct, mk, _opar(mk, par), rm.conn, rm.X.ravel(), u.ravel(), types[0], types[1])
        assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max()
        assert np.abs(en - ren).max() <= RTOL * np.abs(ren).max()
    energy = asm.calc_energy()
    ref = orc.calc_energy(ct, mk, _opar(mk, par), rm.conn, rm.X.ravel(), u.ravel(), ele_is_row=rm.ele_is_row)
    assert abs(energy - ref) <= 1e-11 * abs(ref), (energy, ref)  # atomic summation order
    asm.close()


def test_calc_stress_fbar_and_per_element_mue():
    ct, rm, u = _mesh("hex8")
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25], tech=capi.TECH_FBAR)
    asm.set_state("displacement", u.ravel())
    st, en = asm.calc_stress("cauchy", "gl")
    rst, ren = orc.calc_stress(ct, 0, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(), orc.STRESS_CAUCHY,
                               orc.STRAIN_GL, tech=orc.TECH_FBAR)
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max() and np.abs(en - ren).max() <= RTOL * np.abs(ren).max()
    e = asm.calc_energy()
    re_ = orc.calc_energy(ct, 0, [200.0, 0.25], rm.conn, rm.X.ravel(), u.ravel(), tech=orc.TECH_FBAR)
    assert abs(e - re_) <= 1e-11 * abs(re_)
    asm.close()
    ct, rm, u = _mesh("tet10")
    mue = 1.0 + 0.2 * np.arange(rm.n_col_eles) / rm.n_col_eles
    asm = b200.DeviceAssembler(rm, capi.MAT_ISO_VOL, [1.0, 1000.0], per_ele_param=mue)
    asm.set_state("displacement", u.ravel())
    st, _ = asm.calc_stress("2pk", "none")
    rst, _ = orc.calc_stress(ct, 1, [1.0, 1000.0], rm.conn, rm.X.ravel(), u.ravel(), orc.STRESS_2PK, orc.STRAIN_NONE,
                             mue_per_ele=mue)
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max()
    asm.close()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
@pytest.mark.parametrize("action", ["struct_calc_nlnstiffmass", "struct_calc_nlnstifflmass",
                                    "struct_calc_internalinertiaforce"])
def test_mass_actions_match_oracle(cell, action):
    ct, rm, u = _mesh(cell)
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3], density=7.8)
    rowptr, colind = asm.graph()
    acc = np.random.default_rng(2).uniform(-1, 1, 3 * rm.n_col_nodes)
    asm.set_state("displacement", u[:rm.n_row_nodes].ravel())
    asm.set_state("acceleration", acc[:3 * rm.n_row_nodes])
    K, f, _ = asm.evaluate(action)
    stiff = action != "struct_calc_internalinertiaforce"
    rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_SVK, [1000.0, 0.3], rm.conn, rm.X.ravel(), u.ravel(),
                                      rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                      action=orc.ACTION_FORCE_STIFF if stiff else orc.ACTION_FORCE)
    if stiff:
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok, (wK, wf)
    else:
        assert K is None and np.abs(f - rf).max() <= RTOL * np.abs(rf).max()
    lumped = action == "struct_calc_nlnstifflmass"
    rM, rfi = orc.assemble_mass(ct, 7.8, rm.conn, rm.X.ravel(), rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes,
                                rowptr, colind, lumped=lumped, acc_col=acc)
    M, fi = asm.mass(want_matrix=stiff)
    if stiff:
        assert np.abs(M - rM).max() <= RTOL * np.abs(rM).max()
    assert np.abs(fi - rfi).max() <= RTOL * np.abs(rfi).max()
    asm.close()


def test_mass_action_needs_density():
    ct, rm, u = _mesh("hex8", (2, 2, 2))
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    asm.set_state("displacement", u.ravel())
    with pytest.raises(capi.B200SolidError) as ei:
        asm.evaluate("struct_calc_nlnstiffmass")
    assert ei.value.code == capi.ERR_INVALID
    asm.close()


def _csr(rowptr, colind, vals, ncols):
    return sp.csr_matrix((vals, colind, rowptr), shape=(len(rowptr) - 1, ncols))


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_apply_dirichlet_and_multiply_match_host_csr(cell):
    """SparseMatrix::apply_dirichlet (in-place branch) and multiply on the device-resident K against
    scipy on the host copy of the same values (bitwise for the Dirichlet rows, 1e-12 for the product)."""
    ct, rm, u = _mesh(cell)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate()
    n = asm.n_rows
    rng = np.random.default_rng(4)
    x = rng.uniform(-1, 1, n)
    A = _csr(rowptr, colind, K, n)
    y = asm.ctx.multiply(x)
    ref = A @ x
    assert np.abs(y - ref).max() <= RTOL * np.abs(ref).max()
    toggle = (rng.uniform(0, 1, n) < 0.2).astype(np.float64)
    asm.ctx.apply_dirichlet(toggle, diagonalblock=True)
    Kd = asm.ctx.get_values()
    ref_K = K.copy()
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    dbc = toggle[rows] == 1.0
    ref_K[dbc] = 0.0
    ref_K[dbc & (colind == rows)] = 1.0
    assert np.array_equal(Kd, ref_K)
    y2 = asm.ctx.multiply(x)
    ref2 = _csr(rowptr, colind, ref_K, n) @ x
    assert np.abs(y2 - ref2).max() <= RTOL * np.abs(ref2).max()
    # apply_dirichlet_to_system on host vectors
    xv, bv, val = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    x0, b0 = xv.copy(), bv.copy()
    asm.ctx.apply_dirichlet_to_system(xv, bv, val, toggle)
    assert np.array_equal(xv, np.where(toggle == 1.0, val, x0)) and np.array_equal(bv, np.where(toggle == 1.0, val, b0))
    asm.close()


def test_multiply_general_csr_route_and_mass_matrix():
    """A caller-supplied graph with an extra, non-blocked column forces the general CSR kernel; the mass
    matrix is addressed with which=MATRIX_MASS."""
    ct, rm, u = _mesh("hex8", (4, 3, 2))
    rowptr0, colind0 = orc.build_graph(ct, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    # one extra (never assembled) column in the three rows of row node 0, e.g. a coupling entry of another
    # element block: the rows are no longer made of column triplets
    extra = max(set(range(asm_cols := 3 * rm.n_col_nodes)) - set(colind0[rowptr0[0]:rowptr0[1]].tolist()))
    rows = []
    for r in range(len(rowptr0) - 1):
        c = colind0[rowptr0[r]:rowptr0[r + 1]]
        rows.append(np.sort(np.append(c, extra)) if r < 3 else c)
    rowptr = np.concatenate([[0], np.cumsum([len(c) for c in rows])]).astype(np.int64)
    colind = np.concatenate(rows).astype(np.int32)
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3], graph=(rowptr, colind), density=2.0)
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate("struct_calc_nlnstiffmass")
    rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_SVK, [1000.0, 0.3], rm.conn, rm.X.ravel(), u.ravel(),
                                      rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok, (wK, wf)
    M, _ = asm.mass()
    x = np.random.default_rng(9).uniform(-1, 1, asm.n_rows)
    yK = asm.ctx.multiply(x)
    refK = _csr(rowptr, colind, K, asm.n_rows) @ x
    assert np.abs(yK - refK).max() <= RTOL * np.abs(refK).max()
    yM = asm.ctx.multiply(x, which=capi.MATRIX_MASS)
    ref = _csr(rowptr, colind, M, asm.n_rows) @ x
    assert np.abs(yM - ref).max() <= RTOL * np.abs(ref).max()
    asm.close()


def test_golden_hex8_newton_with_device_resident_linear_algebra():
    """The reference's hex8 ctest input solved with K never leaving the device: evaluate, apply_dirichlet
    and the CG hand-off all operate on the context's CSR; only vectors cross the ABI.  Displacements
    against the RESULT DESCRIPTION known answers (CG tolerance bounds the error: 1e-9 here)."""
    case = json.load(open(os.path.join(G, "integration_hex8.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    par = [case["material"]["YOUNG"], case["material"]["NUE"]]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par)
    nn = len(X)
    toggle = np.zeros(3 * nn)
    for k in case["dirichlet_nodes"]:
        toggle[3 * k:3 * k + 3] = 1.0
    u = np.zeros(3 * nn)
    t = 0.0
    for _ in range(case["numstep"]):
        t += case["timestep"]
        fext = neumann_load(case, t)
        for it in range(30):
            asm.set_state("displacement", u)
            _, fint, _ = asm.evaluate_resident("struct_calc_nlnstiff")
            rhs = -(fint - fext)
            asm.ctx.apply_dirichlet(toggle, True)
            du0 = np.zeros(3 * nn)
            asm.ctx.apply_dirichlet_to_system(du0, rhs, np.zeros(3 * nn), toggle)
            du, iters, rel = asm.ctx.solve_cg(rhs, du0, rtol=1e-13, maxit=500)
            u += du
            if np.linalg.norm(du) < 1e-11:
                break
        else:
            raise AssertionError("Newton did not converge")
    u = u.reshape(nn, 3)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-9, (e, u[e["node"], e["dof"]])
    asm.close()


def test_cg_handoff_on_a_larger_block():
    """K of a 12x10x8 Neo-Hooke block with one face clamped: CG on the device against scipy's direct solve."""
    import scipy.sparse.linalg as spla
    ct, rm, u = _mesh("hex8", (12, 10, 8))
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, [200.0, 0.25])
    rowptr, colind = asm.graph()
    asm.set_state("displacement", 0.2 * u.ravel())
    K, f, _ = asm.evaluate()
    n = asm.n_rows
    toggle = np.repeat((rm.X[:, 0] < 1e-12).astype(np.float64), 3)
    asm.ctx.apply_dirichlet(toggle, True)
    b = np.random.default_rng(1).uniform(-1, 1, n)
    b[toggle == 1.0] = 0.0
    x, iters, rel = asm.ctx.solve_cg(b, rtol=1e-12, maxit=2000)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    Kd = K.copy()
    dbc = toggle[rows] == 1.0
    Kd[dbc] = 0.0
    Kd[dbc & (colind == rows)] = 1.0
    # the Dirichlet-row-only modification is unsymmetric; with zero rhs on those rows CG stays in the
    # subspace x_dbc = 0 where the operator is symmetric positive definite
    ref = spla.spsolve(_csr(rowptr, colind, Kd, n).tocsc(), b)
    assert rel <= 1e-12 and iters < 2000
    assert np.abs(x - ref).max() <= 1e-8 * np.abs(ref).max(), (iters, rel)
    asm.close()


# ------------------------------------------------------------------------------------------------------
# MULF prestress (PRESTRESS_TECH mulf)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
@pytest.mark.parametrize("mi", [0, 3])
def test_mulf_evaluate_update_and_history_match_oracle(cell, mi):
    ct, rm, u = _mesh(cell)
    mk, par = MATS[mi]
    asm = b200.DeviceAssembler(rm, mk, par, prestress=capi.PRESTRESS_MULF)
    rowptr, colind = asm.graph()
    ohist = orc.mulf_new_history(ct, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    ngp = orc.num_gp(ct)
    for scale in (1.0, -0.6, 0.3):
        us = scale * u.ravel()
        asm.set_state("displacement", us[:3 * rm.n_row_nodes])
        K, f, _ = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, _opar(mk, par), rm.conn, rm.X.ravel(), us, *maps,
                                             tech=orc.TECH_MULF, hist=ohist)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok and nbad == 0, (scale, wK, wf)
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
        asm.update_prestress()
        orc.update_prestress(ct, rm.conn, rm.X.ravel(), us, ohist)
        h = asm.ctx.get_gp_history().reshape(rm.n_col_eles, ngp, 18)
        assert np.abs(h - ohist[:, :, :18]).max() <= 1e-13 * np.abs(ohist).max()
    # restart: a second context that unpacks the history continues identically
    asm2 = b200.DeviceAssembler(rm, mk, par, prestress=capi.PRESTRESS_MULF)
    asm2.ctx.set_gp_history(asm.ctx.get_gp_history())
    us = 0.2 * u.ravel()
    for a in (asm, asm2):
        a.set_state("displacement", us[:3 * rm.n_row_nodes])
    Ka, fa, _ = asm.evaluate()
    Kb, fb, _ = asm2.evaluate()
    # (atomic scatter order differs between runs: equal to round-off, not bitwise)
    assert np.abs(fa - fb).max() <= 1e-13 * np.abs(fa).max() and np.abs(Ka - Kb).max() <= 1e-13 * np.abs(Ka).max()
    # stresses see the prestressed state
    st, _ = asm.calc_stress("2pk", "none")
    asm.set_state("displacement", 0.0 * us[:3 * rm.n_row_nodes])
    st0, _ = asm.calc_stress("2pk", "none")
    assert np.abs(st0).max() > 1e-3 and np.abs(st - st0).max() > 1e-6
    asm.close(), asm2.close()


def test_mulf_prestressing_procedure_through_cuda_path():
    """The reference's MULF ctest procedure (two prestress load steps with update + displacement reset,
    then ordinary load steps) with the CUDA path inside the Newton loop, against the same procedure with
    the oracle (whose MULF restatement is pinned to the reference's known answers with the fibre material
    of those inputs; the device block uses the isotropic part only)."""
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(G, "integration_hex8_mulf.json")))
    case = dict(case, numstep=4)
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    nh = case["summands"][0]["params"]
    par = [nh["YOUNG"], nh["NUE"]]
    nn = len(X)
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, prestress=capi.PRESTRESS_MULF)
    rowptr, colind = asm.graph()
    ohist = orc.mulf_new_history(orc.HEX8, len(conn))

    def ev_dev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    def up_dev(u):
        asm.set_state("displacement", u)
        asm.update_prestress()

    def ev_orc(u):
        K, f, _ = orc.evaluate_assemble(orc.HEX8, 0, par, conn, X.ravel(), u, rm.node_row_lid, rm.node_col_lid, nn,
                                        rowptr, colind, tech=orc.TECH_MULF, hist=ohist)
        return rowptr, colind, K, f

    u_dev = solve_case(case, ev_dev, update_prestress=up_dev)
    u_orc = solve_case(case, ev_orc, update_prestress=lambda u: orc.update_prestress(orc.HEX8, conn, X.ravel(), u, ohist))
    assert np.abs(u_dev - u_orc).max() < 1e-10 and np.abs(u_orc).max() > 0.05
    asm.close()


def test_update_prestress_needs_mulf_and_mulf_eas_is_rejected():
    ct, rm, u = _mesh("hex8", (2, 2, 2))
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    asm.set_state("displacement", u.ravel())
    with pytest.raises(capi.B200SolidError) as ei:
        asm.update_prestress()
    assert ei.value.code == capi.ERR_UNSUPPORTED
    with pytest.raises(capi.B200SolidError) as ei:
        asm.ctx.set_element_properties(tech=capi.TECH_EAS_MILD, prestress=capi.PRESTRESS_MULF)
    assert ei.value.code == capi.ERR_UNSUPPORTED
    asm.close()


# ------------------------------------------------------------------------------------------------------
# EAS hex8 (TECH eas_mild / eas_full)
# ------------------------------------------------------------------------------------------------------
EAS = {"eas_mild": (capi.TECH_EAS_MILD, orc.TECH_EAS_MILD), "eas_full": (capi.TECH_EAS_FULL, orc.TECH_EAS_FULL)}
EAS_RTOL = 1e-9


def _hist_close(h, ref, tech):
    """Element history block by block (alpha_inc, alpha, s, invKaa, Kda), each relative to its own scale."""
    n = 9 if tech == "eas_mild" else 21
    cuts = [0, n, 2 * n, 3 * n, 3 * n + n * n, 3 * n + n * n + 24 * n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        scale = max(np.abs(ref[:, a:b]).max(), 1e-300)
        if np.abs(h[:, a:b] - ref[:, a:b]).max() > 1e-8 * scale:
            print("history block", a, b, np.abs(h[:, a:b] - ref[:, a:b]).max() / scale)
            return False
    return True


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
@pytest.mark.parametrize("mi", [0, 1, 2, 3])
def test_eas_evaluate_recover_and_history_match_oracle(tech, mi):
    """Evaluate (K, f after static condensation), struct_calc_recover with a Newton-like increment, evaluate
    again with the updated alpha; the element history (alpha_inc, alpha, s, Kaa^-1, Kda) against the
    oracle's.  The device inverts Kaa by Gauss-Jordan with partial pivoting like the oracle; the history
    tolerance allows for the conditioning of Kaa."""
    ct, rm, u = _mesh("hex8", (6, 5, 4))
    mk, par = MATS[mi]
    dt, ot = EAS[tech]
    asm = b200.DeviceAssembler(rm, mk, par, tech=dt)
    rowptr, colind = asm.graph()
    ohist = orc.eas_new_history(ot, rm.n_col_eles)
    maps = (rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind)
    rng = np.random.default_rng(8)
    # (the nearly incompressible parameters need a state closer to equilibrium: the first recover moves alpha
    #  by -Kaa^-1 s, which must leave det C_enh > 0)
    us = (0.05 if mi == 1 else 1.0) * u.ravel().copy()
    for it in range(2 if mi == 1 else 3):
        asm.set_state("displacement", us)
        K, f, err = asm.evaluate("struct_calc_nlnstiff")
        rK, rf, nbad = orc.evaluate_assemble(ct, mk, _opar(mk, par), rm.conn, rm.X.ravel(), us, *maps, tech=ot,
                                             hist=ohist)
        # static condensation multiplies round-off by cond(Kaa) (1e3..1e4 for the nearly incompressible
        # ISO_VOL parameters): 1e-9 here instead of the 1e-12 of the uncondensed path
        assert np.isfinite(rK).all() and np.isfinite(rf).all(), "oracle state left the admissible range"
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, EAS_RTOL)
        print(f"EAS parity {tech} mat {mi} it {it}: worst K {wK * EAS_RTOL:.2e} f {wf * EAS_RTOL:.2e}")
        assert ok and nbad == 0 and err.n_neg_jacobian == 0, (it, wK, wf)
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= EAS_RTOL * np.abs(rf).max()
        h = asm.ctx.get_element_history().reshape(ohist.shape)
        assert _hist_close(h, ohist, tech), it
        du = 0.02 / max(rm.n_col_eles, 1) ** (1 / 3) * rng.uniform(-1, 1, us.size) * 0.1
        us = us + du
        asm.recover(du, 1.0)
        orc.eas_recover(ot, rm.conn, du, ohist, 1.0)
        h = asm.ctx.get_element_history().reshape(ohist.shape)
        assert _hist_close(h, ohist, tech), it
    # restart round trip
    asm2 = b200.DeviceAssembler(rm, mk, par, tech=dt)
    asm2.ctx.set_element_history(asm.ctx.get_element_history())
    for a in (asm, asm2):
        a.set_state("displacement", us)
    Ka, fa, _ = asm.evaluate()
    Kb, fb, _ = asm2.evaluate()
    assert np.abs(fa - fb).max() <= 1e-13 * np.abs(fa).max() and np.abs(Ka - Kb).max() <= 1e-13 * np.abs(Ka).max()
    asm.close(), asm2.close()


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
def test_eas_newton_procedure_through_cuda_path(tech):
    """The reference's TECH eas_* ctest procedure (Newton with struct_calc_recover after every update) with
    the CUDA path inside the loop against the oracle (pinned to the reference's known answers with the fibre
    material of those inputs; the device block uses the isotropic part)."""
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(G, f"integration_hex8_{tech}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    nh = case["summands"][0]["params"]
    par = [nh["YOUNG"], nh["NUE"]]
    nn = len(X)
    dt, ot = EAS[tech]
    asm = b200.DeviceAssembler(rm, capi.MAT_COUP_NEOHOOKE, par, tech=dt)
    rowptr, colind = asm.graph()
    ohist = orc.eas_new_history(ot, len(conn))

    def ev_dev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    def ev_orc(u):
        K, f, _ = orc.evaluate_assemble(orc.HEX8, 0, par, conn, X.ravel(), u, rm.node_row_lid, rm.node_col_lid, nn,
                                        rowptr, colind, tech=ot, hist=ohist)
        return rowptr, colind, K, f

    u_dev = solve_case(case, ev_dev, recover=lambda du: asm.recover(du, 1.0))
    u_orc = solve_case(case, ev_orc, recover=lambda du: orc.eas_recover(ot, conn, du, ohist))
    assert np.abs(u_dev - u_orc).max() < 1e-9 and np.abs(u_orc).max() > 0.05
    asm.close()


def test_eas_is_hex8_only_and_recover_is_a_noop_without_eas():
    ct, rm, u = _mesh("hex27", (2, 2, 2))
    ctx = capi.Context(0)
    ctx.set_mesh(ct, rm.X, rm.conn)
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_element_properties(tech=capi.TECH_EAS_FULL)
    assert ei.value.code == capi.ERR_UNSUPPORTED
    ctx.close()
    ct, rm, u = _mesh("hex8", (2, 2, 2))
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    asm.recover(np.zeros(asm.n_rows))
    asm.close()


# ------------------------------------------------------------------------------------------------------
# The reference's own MULF / EAS ctest inputs (fibre material) through the CUDA path
# ------------------------------------------------------------------------------------------------------
def _fibre_material(case):
    nh, fib = case["summands"][0]["params"], case["summands"][1]["params"]
    return [("COUP_NEOHOOKE", [nh["YOUNG"], nh["NUE"]]),
            ("COUP_ANISOEXPO", [fib["K1"], fib["K2"], fib["K1COMP"], fib["K2COMP"]])]


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_reference_mulf_known_answers_through_cuda_path(cell):
    """RESULT DESCRIPTION displacements (reference tolerance 1e-12) of tests/input_files/
    solid_ele_<cell>_Standard_mulf.4C.yaml with the CUDA path doing every evaluate and prestress update:
    MAT_ElastHyper{CoupNeoHooke, CoupAnisoExpo} on the element fibres, PRESTRESS_TECH mulf."""
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(G, f"integration_{cell}_mulf.json")))
    ct = CELLS[cell]
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(ct, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, _fibre_material(case), prestress=capi.PRESTRESS_MULF,
                               fibers=np.asarray(case["fibers"], float))
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


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
def test_reference_eas_known_answers_through_cuda_path(tech):
    """RESULT DESCRIPTION displacements of the reference's TECH eas_mild / eas_full hex8 inputs with the
    CUDA path doing every evaluate and struct_calc_recover (fibre material on the device)."""
    from tests.helpers.newton import solve_case
    case = json.load(open(os.path.join(G, f"integration_hex8_{tech}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, _fibre_material(case), tech=EAS[tech][0],
                               fibers=np.asarray(case["fibers"], float))
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, ev, recover=lambda du: asm.recover(du, 1.0))
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, u[e["node"], e["dof"]])
    # nodal stresses / strains of the same inputs from the device's struct_calc_stress (eas_mild: 2PK + GL,
    # eas_full: Cauchy + Euler-Almansi through the enhanced deformation gradient)
    from tests.helpers.postprocess import nodal_stresses
    asm.set_state("displacement", u.ravel())
    st, en = asm.calc_stress(case["iostress"].lower(), case["iostrain"].lower())
    ns, ne = nodal_stresses(orc.HEX8, conn, st, len(X)), nodal_stresses(orc.HEX8, conn, en, len(X))
    for e in case["expected_stress"]:
        assert abs(ns[e["node"], e["comp"]] - e["value"]) < 1e-8, (e, ns[e["node"], e["comp"]])
    for e in case["expected_strain"]:
        assert abs(ne[e["node"], e["comp"]] - e["value"]) < 1e-9, (e, ne[e["node"], e["comp"]])
    asm.close()


def test_fibre_material_needs_fibres_and_stress_output_matches_oracle():
    ct, rm, u = _mesh("hex8")
    mat = [("COUP_NEOHOOKE", [10.0, 0.25]), ("COUP_ANISOEXPO", [100.0, 1.0, 0.0, 1.0])]
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, mat)
    asm.set_state("displacement", u.ravel())
    with pytest.raises(capi.B200SolidError) as ei:
        asm.evaluate()
    assert ei.value.code == capi.ERR_INVALID
    rng = np.random.default_rng(3)
    fib = rng.uniform(-1, 1, (rm.n_col_eles, 3)) + np.array([2.0, 0, 0])  # not normalised on purpose
    asm.ctx.set_element_fibers(fib)
    rowptr, colind = asm.graph()
    K, f, _ = asm.evaluate()
    opar = orc.pack_elasthyper(mat + [("FIBER", [1.0, 0.0, 0.0])])
    rK, rf, _ = orc.evaluate_assemble(ct, orc.MAT_ELASTHYPER, opar, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                      rm.node_col_lid, rm.n_row_nodes, rowptr, colind, mue_per_ele=fib.ravel())
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok, (wK, wf)
    asm.close()


# ------------------------------------------------------------------------------------------------------
# KINEM linear
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
@pytest.mark.parametrize("mi", [2, 0])
def test_linear_kinematics_match_oracle(cell, mi):
    ct, rm, u = _mesh(cell)
    mk, par = MATS[mi]
    asm = b200.DeviceAssembler(rm, mk, par, kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, _ = asm.evaluate()
    rK, rf, _ = orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid, rm.node_col_lid,
                                      rm.n_row_nodes, rowptr, colind, tech=orc.TECH_LINEAR)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok, (wK, wf)
    st, en = asm.calc_stress("cauchy", "gl")
    rst, ren = orc.calc_stress(ct, mk, par, rm.conn, rm.X.ravel(), u.ravel(), orc.STRESS_CAUCHY, orc.STRAIN_GL,
                               tech=orc.TECH_LINEAR)
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max() and np.abs(en - ren).max() <= RTOL * np.abs(ren).max()
    asm.close()
    ctx = capi.Context(0)
    with pytest.raises(capi.B200SolidError) as ei:
        ctx.set_element_properties(kinem=capi.KINEM_LINEAR, tech=capi.TECH_FBAR)
    assert ei.value.code == capi.ERR_UNSUPPORTED
    ctx.close()


def test_eas_linesearch_correction_and_step_update_reset():
    """correct_condensed_variables_for_linesearch: alpha += (new - old step) * alpha_inc; struct_calc_update_istep /
    struct_calc_reset_istep: snapshot and restore of the element history (EAS) and Gauss-point history (MULF)."""
    ct, rm, u = _mesh("hex8")
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3], tech=capi.TECH_EAS_MILD)
    asm.set_state("displacement", u.ravel())
    asm.evaluate()
    du = 1e-3 * np.random.default_rng(4).uniform(-1, 1, asm.n_rows)
    asm.recover(du, 1.0)
    h1 = asm.ctx.get_element_history().reshape(rm.n_col_eles, -1)
    asm.ctx.update_step()
    asm.ctx.recover_linesearch(0.25)
    h2 = asm.ctx.get_element_history().reshape(rm.n_col_eles, -1)
    assert np.abs(h2[:, 9:18] - (h1[:, 9:18] - 0.75 * h1[:, :9])).max() <= 1e-15 * np.abs(h1[:, 9:18]).max()
    assert np.array_equal(h2[:, :9], h1[:, :9]) and np.array_equal(h2[:, 18:], h1[:, 18:])
    asm.ctx.reset_step()
    assert np.array_equal(asm.ctx.get_element_history().reshape(rm.n_col_eles, -1), h1)
    asm.close()
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3], prestress=capi.PRESTRESS_MULF)
    asm.set_state("displacement", u.ravel())
    asm.update_prestress()
    g1 = asm.ctx.get_gp_history()
    asm.ctx.update_step()
    asm.set_state("displacement", -0.5 * u.ravel())
    asm.update_prestress()
    assert np.abs(asm.ctx.get_gp_history() - g1).max() > 1e-6
    asm.ctx.reset_step()
    assert np.array_equal(asm.ctx.get_gp_history(), g1)
    asm.close()
    plain = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    plain.ctx.update_step(), plain.ctx.reset_step(), plain.ctx.recover_linesearch(0.5)  # no-ops
    plain.close()


@pytest.mark.parametrize("tech", ["eas_mild", "eas_full"])
def test_eas_stress_and_energy_output_match_oracle(tech):
    ct, rm, u = _mesh("hex8", (5, 4, 3))
    dt, ot = EAS[tech]
    mk, par = MATS[0]
    asm = b200.DeviceAssembler(rm, mk, par, tech=dt)
    rowptr, colind = asm.graph()
    ohist = orc.eas_new_history(ot, rm.n_col_eles)
    us = u.ravel().copy()
    asm.set_state("displacement", us)
    asm.evaluate()
    orc.evaluate_assemble(ct, mk, par, rm.conn, rm.X.ravel(), us, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes,
                          rowptr, colind, tech=ot, hist=ohist)
    du = 1e-4 * np.random.default_rng(6).uniform(-1, 1, us.size)
    asm.recover(du, 1.0)
    orc.eas_recover(ot, rm.conn, du, ohist, 1.0)
    us = us + du
    asm.set_state("displacement", us)
    assert np.abs(ohist[:, 9 if tech == "eas_mild" else 21:][:, :5]).max() > 0  # alpha is not trivially zero
    for io, types in ((("2pk", "gl"), (orc.STRESS_2PK, orc.STRAIN_GL)), (("cauchy", "ea"), (orc.STRESS_CAUCHY, orc.STRAIN_EA))):
        st, en = asm.calc_stress(*io)
        rst, ren = orc.calc_stress(ct, mk, par, rm.conn, rm.X.ravel(), us, types[0], types[1], tech=ot, hist=ohist)
        assert np.abs(st - rst).max() <= 1e-10 * np.abs(rst).max(), (io, np.abs(st - rst).max() / np.abs(rst).max())
        assert np.abs(en - ren).max() <= 1e-10 * np.abs(ren).max(), (io, np.abs(en - ren).max() / np.abs(ren).max())
    e_dev = asm.calc_energy()
    e_ref = orc.calc_energy(ct, mk, par, rm.conn, rm.X.ravel(), us, tech=ot, hist=ohist)
    assert abs(e_dev - e_ref) <= 1e-10 * abs(e_ref)
    asm.close()


# ------------------------------------------------------------------------------------------------------
# summands in (modified) principal stretches: ELAST_IsoOgden, ELAST_CoupVarga, ELAST_IsoVarga
# ------------------------------------------------------------------------------------------------------
STRETCH = [("ISO_OGDEN", [0.8, 4.0]), ("COUP_VARGA", [1.0, 0.3]), ("ISO_VARGA", [2.0, 0.7]), ("VOL_SUSSMANBATHE", [50.0])]


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_stretch_summands_evaluate_and_stress_match_oracle(cell):
    ct, rm, u = _mesh(cell)
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, STRETCH)
    rowptr, colind = asm.graph()
    asm.set_state("displacement", u.ravel())
    K, f, err = asm.evaluate("struct_calc_nlnstiff")
    opar = orc.pack_elasthyper(STRETCH)
    rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_ELASTHYPER, opar, rm.conn, rm.X.ravel(), u.ravel(), rm.node_row_lid,
                                         rm.node_col_lid, rm.n_row_nodes, rowptr, colind, nthreads=4)
    ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
    assert ok and nbad == 0, (wK, wf)
    st, en = asm.calc_stress("cauchy", "ea")
    rst, ren = orc.calc_stress(ct, orc.MAT_ELASTHYPER, opar, rm.conn, rm.X.ravel(), u.ravel(), orc.STRESS_CAUCHY,
                               orc.STRAIN_EA)
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max() and np.abs(en - ren).max() <= RTOL * np.abs(ren).max()
    # like the reference: "Summand does not support calculation of strain energy" (4C_mat_elast_summand.hpp:127-139)
    with pytest.raises(capi.B200SolidError) as ei:
        asm.calc_energy()
    assert ei.value.code == capi.ERR_UNSUPPORTED and "strain energy" in str(ei.value)
    # the undeformed state: three equal stretches, the reference's limit branch (a lone Varga term is not
    # stress free there, so the force is compared too)
    asm.set_state("displacement", 0.0 * u.ravel())
    K0, f0, _ = asm.evaluate()
    rK0, rf0, _ = orc.evaluate_assemble(ct, orc.MAT_ELASTHYPER, opar, rm.conn, rm.X.ravel(), 0.0 * u.ravel(),
                                      rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind, nthreads=4)
    assert np.abs(f0 - rf0).max() <= 1e-9 * np.abs(K0).max()
    assert np.abs(K0 - rK0).max() <= 1e-9 * np.abs(rK0).max()
    asm.close()


def test_reference_isoogden_known_answers_through_cuda_path():
    """Nodal Cauchy stresses of tests/input_files/elasthyper_isoogden.4C.yaml (MAT_ElastHyper{IsoOgden, VolOgden},
    prescribed displacements) with the device doing every evaluate and the stress output."""
    from tests.helpers.newton import solve_prescribed_displacement_case
    from tests.helpers.postprocess import nodal_stresses
    from tests.test_stretch_summands import case_material
    case = json.load(open(os.path.join(G, "integration_hex8_isoogden.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_ELASTHYPER, case_material(case))
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_prescribed_displacement_case(case, ev)
    asm.set_state("displacement", u.ravel())
    st, _ = asm.calc_stress("cauchy", "none")
    ns = nodal_stresses(orc.HEX8, conn, st, len(X))
    for e in case["expected_stress"]:
        assert abs(ns[e["node"], e["comp"]] - e["value"]) < 1e-6, (e, ns[e["node"], e["comp"]])
    asm.close()


# ------------------------------------------------------------------------------------------------------
# history material on the device GP-state store: MAT_Struct_DruckerPrager (KINEM linear)
# ------------------------------------------------------------------------------------------------------
DP_MAT = dict(YOUNG=10.0, NUE=0.3, ISOHARD=0.5, TOL=1e-12, C=0.05, ETA=0.2, XI=0.25, ETABAR=0.2)


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
@pytest.mark.parametrize("tang", ["consistent", "elastic"])
def test_history_material_evaluate_and_state_store_match_oracle(cell, tang):
    ct, rm, u = _mesh(cell)
    u = 4.0 * u.ravel()                                   # well into the plastic range
    par = capi.drucker_prager_params(**dict(DP_MAT, TANG=tang))
    asm = b200.DeviceAssembler(rm, capi.MAT_DRUCKER_PRAGER, par, kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()
    ngp = {"hex8": 8, "hex27": 27, "tet10": 4}[cell]
    h_o = np.zeros((rm.n_col_eles, ngp, orc.DP_HIST))
    assert asm.ctx.get_material_history().size == h_o.size and not asm.ctx.get_material_history().any()
    for fac in (1.0, 0.6, 1.5):                            # load, commit, unload, commit, reload beyond
        asm.set_state("displacement", fac * u)
        K, f, _ = asm.evaluate()
        rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_DRUCKER_PRAGER, par, rm.conn, rm.X.ravel(), fac * u,
                                             rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                             tech=orc.TECH_LINEAR, hist=h_o)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok and nbad == 0, (fac, wK, wf)
        h_d = asm.ctx.get_material_history().reshape(h_o.shape)
        assert np.abs(h_d - h_o).max() <= 1e-12 * np.abs(h_o).max()
        assert (h_o[:, :, 13] > 0).mean() > 0.2
        # struct_calc_internalforce leaves the same iterate
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
        asm.ctx.update_step()
        orc.material_update(orc.MAT_DRUCKER_PRAGER, h_o)
    h_conv = asm.ctx.get_material_history().reshape(h_o.shape)
    assert np.array_equal(h_conv[:, :, :7], h_conv[:, :, 7:])
    # reset_step discards the iterate of a rejected step; restart = get / set of the state store
    asm.set_state("displacement", 2.5 * u)
    asm.evaluate()
    assert not np.array_equal(asm.ctx.get_material_history(), h_conv.ravel())
    asm.ctx.reset_step()
    assert np.array_equal(asm.ctx.get_material_history(), h_conv.ravel())
    asm2 = b200.DeviceAssembler(rm, capi.MAT_DRUCKER_PRAGER, par, kinem=capi.KINEM_LINEAR)
    asm2.ctx.set_material_history(h_conv)
    asm2.set_state("displacement", 0.9 * u)
    asm.set_state("displacement", 0.9 * u)
    Ka, fa, _ = asm.evaluate()
    Kb, fb, _ = asm2.evaluate()
    assert np.abs(fa - fb).max() <= 1e-13 * np.abs(fa).max() and np.abs(Ka - Kb).max() <= 1e-13 * np.abs(Ka).max()
    # stress output of the current iterate (linear kinematics: Cauchy == 2PK of the small-strain model)
    st, en = asm.calc_stress("2pk", "gl")
    S_ref = np.zeros_like(st)
    Xc, uc = rm.X[rm.conn], (0.9 * u).reshape(-1, 3)[rm.conn]
    gxi, _ = orc.gauss(ct)
    for g, xi in enumerate(gxi):
        _, dN = orc.shape(ct, xi)
        for e in range(rm.n_col_eles):
            Gx = np.linalg.solve(dN @ Xc[e], dN)          # N_XYZ = J^-1 dN, (3, nn)
            H = uc[e].T @ Gx.T                            # displacement gradient
            E = 0.5 * (H + H.T)
            h = h_conv[e, g].copy()
            S_ref[e, g], _ = orc.material_hist(orc.MAT_DRUCKER_PRAGER, par, [E[0, 0], E[1, 1], E[2, 2], E[0, 1], E[1, 2], E[0, 2]], h)
    assert np.abs(st - S_ref).max() <= 1e-11 * np.abs(S_ref).max()
    with pytest.raises(capi.B200SolidError) as ei:
        asm.calc_energy()
    assert ei.value.code == capi.ERR_UNSUPPORTED
    asm.close()
    asm2.close()


def test_history_material_restrictions():
    ct, rm, u = _mesh("hex8")
    par = capi.drucker_prager_params(**DP_MAT)
    asm = b200.DeviceAssembler(rm, capi.MAT_DRUCKER_PRAGER, par)          # nonlinear kinematics
    asm.set_state("displacement", u.ravel())
    with pytest.raises(capi.B200SolidError) as ei:
        asm.evaluate()
    assert ei.value.code == capi.ERR_UNSUPPORTED and "KINEM linear" in str(ei.value)
    # non-associative consistent tangent: accepted (general-tangent kernel); zero friction parameters are not
    asm.ctx.set_material(capi.MAT_DRUCKER_PRAGER, capi.drucker_prager_params(**dict(DP_MAT, ETABAR=0.4)))
    asm.ctx.set_material(capi.MAT_DRUCKER_PRAGER, capi.drucker_prager_params(**dict(DP_MAT, ETABAR=0.4, TANG="elastic")))
    with pytest.raises(capi.B200SolidError) as ei:
        asm.ctx.set_material(capi.MAT_DRUCKER_PRAGER, capi.drucker_prager_params(**dict(DP_MAT, ETA=0.0)))
    assert ei.value.code == capi.ERR_INVALID and "non-zero" in str(ei.value)
    with pytest.raises(capi.B200SolidError):
        asm.ctx.set_material(capi.MAT_DRUCKER_PRAGER, [10.0, 0.3])      # too few parameters
    # a material without history has no state store
    asm.ctx.set_material(capi.MAT_SVK, [1000.0, 0.3])
    assert asm.ctx.L.b200s_material_history_size(asm.ctx.h) == 0
    with pytest.raises(capi.B200SolidError) as ei:
        asm.ctx.get_material_history()
    assert ei.value.code == capi.ERR_INVALID
    asm.close()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_unsymmetric_tangent_block_matches_oracle(cell):
    """MAT_Struct_DruckerPrager with ETA != ETABAR and the consistent tangent (no major symmetry, K not symmetric):
    the general-tangent kernel against the oracle, incl. the state store over a commit."""
    ct, rm, u = _mesh(cell)
    u = 4.0 * u.ravel()
    par = capi.drucker_prager_params(**dict(DP_MAT, ETA=0.1, ETABAR=0.25))
    asm = b200.DeviceAssembler(rm, capi.MAT_DRUCKER_PRAGER, par, kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()
    ngp = {"hex8": 8, "hex27": 27, "tet10": 4}[cell]
    h_o = np.zeros((rm.n_col_eles, ngp, orc.DP_HIST))
    for fac in (1.0, 1.5):
        asm.set_state("displacement", fac * u)
        K, f, _ = asm.evaluate()
        rK, rf, nbad = orc.evaluate_assemble(ct, orc.MAT_DRUCKER_PRAGER, par, rm.conn, rm.X.ravel(), fac * u,
                                             rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes, rowptr, colind,
                                             tech=orc.TECH_LINEAR, hist=h_o)
        ok, wK, wf = b200.parity_report(rowptr, K, rK, f, rf, RTOL)
        assert ok and nbad == 0, (fac, wK, wf)
        h_d = asm.ctx.get_material_history().reshape(h_o.shape)
        assert np.abs(h_d - h_o).max() <= 1e-12 * np.abs(h_o).max()
        _, f2, _ = asm.evaluate("struct_calc_internalforce")
        assert np.abs(f2 - rf).max() <= RTOL * np.abs(rf).max()
        asm.ctx.update_step()
        orc.material_update(orc.MAT_DRUCKER_PRAGER, h_o)
    import scipy.sparse as sp
    A = sp.csr_matrix((K, colind, rowptr), shape=(asm.n_rows, asm.n_rows))
    assert abs(A - A.T).max() > 1e-6 * np.abs(K).max()
    asm.close()


@pytest.mark.parametrize("name", ["RetToCone_elastic_tang", "RetToCone"])
def test_reference_druckerprager_known_answers_through_cuda_path(name):
    """Tip displacements (1e-10) of tests/input_files/mat_druckerprager_RetToCone[_elastic_tang].4C.yaml with the
    device doing every evaluate and the history commits (20 load steps, 68 yielding Gauss points).  RetToCone
    itself uses the non-associative consistent tangent (ETA 0.1, ETABAR 0.25): general-tangent kernel."""
    from tests.helpers.newton import solve_case
    from tests.helpers.neumann import point_neumann_load
    case = json.load(open(os.path.join(G, f"integration_hex8_druckerprager_{name}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    rm = b200.partition.serial_mesh(orc.HEX8, X, conn)
    asm = b200.DeviceAssembler(rm, capi.MAT_DRUCKER_PRAGER, capi.drucker_prager_params(**case["material"]),
                               kinem=capi.KINEM_LINEAR)
    rowptr, colind = asm.graph()

    def ev(u):
        asm.set_state("displacement", u)
        K, f, _ = asm.evaluate()
        return rowptr, colind, K, f

    u = solve_case(case, ev, max_iter=400, external=lambda u, t, v: (point_neumann_load(case, t), v),
                   update_step=lambda u: asm.ctx.update_step())
    h = asm.ctx.get_material_history().reshape(len(conn), 8, orc.DP_HIST)
    assert int((h[:, :, 6] > 0).sum()) == 68
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-10, (e, u[e["node"], e["dof"]])
    asm.close()


def test_reference_strain_measure_known_answer_through_cuda_path():
    """unittests/solid_ele/4C_solid_ele_utils_test.cpp:26-50: Euler-Almansi strain of a given deformation gradient (1e-13),
    here as struct_calc_stress output of a unit hex8 cube under the homogeneous deformation u = (F - I) X."""
    F = np.array([[1.1, 0.2, 0.5], [0.14, 1.2, 0.3], [0.05, 0.2, 1.3]])
    ea_ref = np.array([0.055233442151184, 0.101134166403205, 0.104112596224498, 0.0913211447369115,
                       0.1073842904312605, 0.157679374545429])
    X = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    u = X @ (F - np.eye(3)).T
    rm = b200.partition.serial_mesh(orc.HEX8, X, np.arange(8, dtype=np.int32)[None, :])
    asm = b200.DeviceAssembler(rm, capi.MAT_SVK, [1000.0, 0.3])
    asm.set_state("displacement", u.ravel())
    st, en = asm.calc_stress("cauchy", "ea")
    assert np.abs(en[0] - ea_ref).max() < 1e-13
    rst, _ = orc.calc_stress(orc.HEX8, orc.MAT_SVK, [1000.0, 0.3], rm.conn, X.ravel(), u.ravel(), orc.STRESS_CAUCHY,
                             orc.STRAIN_NONE)
    assert np.abs(st - rst).max() <= RTOL * np.abs(rst).max()
    asm.close()

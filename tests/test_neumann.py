"""External loads: surface Neumann conditions (Live, orthopressure with load
linearisation, pseudo_orthopressure) and the volume load.

CPU suite: the test-side restatement (tests/helpers/neumann.py) is pinned by the reference's own known
answers (the Live / volume loads by every integration case of the suite, the follower pressure and its
linearisation by tests/input_files/loadlin.4C.yaml), and the PRODUCT's per-thread functions
(4c_b200/csrc/neumann_kernels.cuh, stepped on the CPU by tests/cpp/emulate_neumann.cpp) are compared
with it.  The GPU counterpart is tests/test_gpu_neumann.py.
"""
import json
import os

import numpy as np
import pytest

import importlib

from oracle import oracle as orc
from tests.helpers import emulator as emu
from tests.helpers import neumann as nm
from tests.helpers.meshes import serial_maps
from tests.helpers.newton import solve_case

b200 = importlib.import_module("4c_b200")
G = os.path.join(os.path.dirname(__file__), "golden")
CELLS = {"hex8": orc.HEX8, "hex27": orc.HEX27, "tet10": orc.TET10}


def synthetic_case(cell, rm, pick, val=-0.7):
    """Case dict (the layout of the golden fixtures) of a follower pressure on the nodes pick(X) selects."""
    X = np.asarray(rm.X, float)
    return {"celltype": cell, "coords": X.tolist(), "elements": np.asarray(rm.conn).tolist(),
            "neumann_nodes": np.nonzero(pick(X))[0].tolist(),
            "neumann": {"TYPE": "orthopressure", "ONOFF": [1, 0, 0], "VAL": [val, 0, 0], "FUNCT": [1, 1, 1],
                        "FUNCT1_TIMES": [0.0, 1.0], "FUNCT1_VALUES": [0.0, 2.0]}}


def mesh(cell, rank=0, nranks=1):
    if cell == "tet10":
        X, conn = b200.mesh.unstructured_tet10(2, seed=3)
        rm = b200.partition.serial_mesh(CELLS[cell], X, conn)
    else:
        rm = b200.partition.box_partition(CELLS[cell], (3, 2, 2), rank, nranks)
    u = b200.mesh.displacement_field(rm.X, rm.node_gid, 1.0, 0.4)
    return rm, np.asarray(u, float).ravel()


def top(X):
    return X[:, 2] > X[:, 2].max() - 1e-9


def test_follower_pressure_known_answer_loadlin():
    """tests/input_files/loadlin.4C.yaml (statics, 300 hex8 EAS-full elements, TYPE orthopressure with
    LOADLIN): dispy of node 404 = -22.839151032191484 (1e-8).  Pins the restatement of the follower
    pressure, of its linearisation (Newton only converges quadratically to TOLRES with it) and of the face
    orientation (node order of the parent's surfaces)."""
    case = json.load(open(os.path.join(G, "integration_hex8_loadlin.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(orc.HEX8, conn, row, col, nn)
    par = np.array([case["material"]["YOUNG"], case["material"]["NUE"]])
    hist = orc.eas_new_history(orc.TECH_EAS_FULL, len(conn))
    faces = nm.loaded_faces(case)
    assert faces.shape == (300, 4)

    def evaluate(u):
        vals, f, nbad = orc.evaluate_assemble(orc.HEX8, orc.MAT_COUP_NEOHOOKE, par, conn, X.ravel(), u, row, col,
                                              nn, rowptr, colind, tech=orc.TECH_EAS_FULL, hist=hist)
        assert nbad == 0
        return rowptr, colind, vals, f

    def external(u, t, values):
        f, K = nm.orthopressure_load(case, u, t, loadlin=values is not None, faces=faces)
        return f, (None if values is None else nm.add_to_csr_values(K, rowptr, colind, values))

    u = solve_case(case, evaluate, recover=lambda du: orc.eas_recover(orc.TECH_EAS_FULL, conn, du, hist),
                   external=external)
    for e in case["expected_disp"]:
        assert abs(u[e["node"], e["dof"]] - e["value"]) < 1e-9, (e, u[e["node"], e["dof"]])


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_load_stiffness_is_the_derivative_of_the_follower_load(cell):
    rm, u = mesh(cell)
    case = synthetic_case(cell, rm, top)
    f0, K = nm.orthopressure_load(case, u, 0.5)
    assert np.abs(f0).max() > 0
    K = K.toarray()
    h = 1e-6
    rng = np.random.default_rng(1)
    for dof in rng.choice(np.nonzero(np.abs(K).sum(axis=0))[0], 6, replace=False):
        d = np.zeros_like(u)
        d[dof] = h
        fp, _ = nm.orthopressure_load(case, u + d, 0.5, loadlin=False)
        fm, _ = nm.orthopressure_load(case, u - d, 0.5, loadlin=False)
        # the load stiffness enters K_T = d(f_int - f_ext)/du
        assert np.abs(-(fp - fm) / (2 * h) - K[:, dof]).max() < 1e-8 * max(1.0, np.abs(K).max())


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_product_follower_pressure_matches_restatement(cell):
    rm, u = mesh(cell)
    case = synthetic_case(cell, rm, top)
    faces = nm.loaded_faces(case)
    assert len(faces) > 0
    rowptr, colind = orc.build_graph(CELLS[cell], rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    val = [case["neumann"]["VAL"][0] * nm.time_factor(case, 0.5), 0, 0]
    ref_f, ref_K = nm.orthopressure_load(case, u, 0.5)
    f, K, gp, flag = emu.neumann_surface(CELLS[cell], 2, faces, rm.X.ravel(), u, [1, 0, 0], val, rowptr, colind,
                                         rm.node_row_lid, rm.node_col_lid, 3 * rm.n_row_nodes, loadlin=True,
                                         want_gp=True)
    assert flag == 0
    assert np.abs(f - ref_f).max() <= 1e-13 * np.abs(ref_f).max()
    refK = nm.add_to_csr_values(ref_K, rowptr, colind, np.zeros(rowptr[-1]))
    assert np.abs(refK).max() > 0
    assert np.abs(K - refK).max() <= 1e-13 * np.abs(refK).max()
    # Gauss-point positions handed to the host's function manager: sum_b N_b(xi_g) x_b on the deformed patch
    kind, fn = {"hex8": ("quad4", nm._quad4), "hex27": ("quad9", nm._quad9), "tet10": ("tri6", nm._tri6)}[cell]
    x = (rm.X + u.reshape(-1, 3))[faces]
    for g, (r, s) in enumerate(nm._rule(kind)[0]):
        assert np.abs(gp[:, g] - np.einsum("b,fbk->fk", fn(r, s)[0], x)).max() < 1e-14
    # material configuration (MULF active / pseudo-orthopressure on an undeformed last state): no u
    ref_f0, _ = nm.orthopressure_load(case, u, 0.5, material=True)
    f0, _, _, _ = emu.neumann_surface(CELLS[cell], 1, faces, rm.X.ravel(), None, [1, 0, 0], val, rowptr, colind,
                                      rm.node_row_lid, rm.node_col_lid, 3 * rm.n_row_nodes)
    assert np.abs(f0 - ref_f0).max() <= 1e-13 * np.abs(ref_f0).max()
    # resultant on the flat top face of the unit box: pressure * area along the normal
    if cell != "tet10":
        area = 1.0  # top face of the unit box
        assert abs(f0.reshape(-1, 3).sum(axis=0)[2] - val[0] * area) < 1e-12


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_product_live_load_matches_the_known_answer_pinned_restatement(cell):
    """DESIGN SURF NEUMANN of the reference's integration inputs: VAL * FUNCT with FUNCT1 = t and
    FUNCT2 = (x+1)(y+2)(z+3) evaluated by the HOST at the Gauss-point positions the product returns."""
    case = json.load(open(os.path.join(G, f"integration_{cell}.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, colind = orc.build_graph(CELLS[cell], conn, row, col, nn)
    faces = nm.loaded_faces(case)
    n = case["neumann"]
    t = 0.7
    ref = nm.neumann_load(case, t)
    _, _, gp, _ = emu.neumann_surface(CELLS[cell], 0, faces, X.ravel(), None, n["ONOFF"][:3], n["VAL"][:3], rowptr,
                                      colind, row, col, 3 * nn, want_gp=True)
    fac = np.ones(gp.shape)
    for d in range(3):
        if n["FUNCT"][d] == 1:
            fac[..., d] = t
        elif n["FUNCT"][d] == 2:
            fac[..., d] = (gp[..., 0] + 1) * (gp[..., 1] + 2) * (gp[..., 2] + 3)
    f, _, _, _ = emu.neumann_surface(CELLS[cell], 0, faces, X.ravel(), None, n["ONOFF"][:3], n["VAL"][:3], rowptr,
                                     colind, row, col, 3 * nn, gp_fac=fac)
    assert np.abs(ref).max() > 0
    assert np.abs(f - ref).max() <= 1e-13 * np.abs(ref).max()


def _volume_gp_factors(cell, case, conn, X, t):
    gxi, _ = orc.gauss(CELLS[cell])
    vn = case["vol_neumann"]
    fac = np.ones((len(conn), len(gxi), 3))
    for g, xi in enumerate(gxi):
        N, _ = orc.shape(CELLS[cell], xi)
        p = np.einsum("a,eak->ek", N, X[conn])
        for d in range(3):
            if vn["FUNCT"][d] == 1:
                fac[:, g, d] = t
            elif vn["FUNCT"][d] == 2:
                fac[:, g, d] = (p[:, 0] + 1) * (p[:, 1] + 2) * (p[:, 2] + 3)
    return fac


def test_product_volume_load_matches_the_known_answer_pinned_restatement():
    """DESIGN VOL NEUMANN of the EAS inputs (pinned by their displacement known answers)."""
    case = json.load(open(os.path.join(G, "integration_hex8_eas_full.json")))
    X = np.asarray(case["coords"], float)
    conn = np.asarray(case["elements"], np.int32)
    nn = len(X)
    row, col = serial_maps(nn)
    rowptr, _ = orc.build_graph(orc.HEX8, conn, row, col, nn)
    vn = case["vol_neumann"]
    t = 0.5
    ref = nm.volume_neumann_load(case, t)
    fac = _volume_gp_factors("hex8", case, conn, X, t)
    f = emu.neumann_volume(orc.HEX8, conn, X.ravel(), vn["ONOFF"][:3], vn["VAL"][:3], rowptr, row, col, 3 * nn,
                           gp_fac=fac)
    assert np.abs(ref).max() > 0
    assert np.abs(f - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("cell", ["hex8", "hex27", "tet10"])
def test_product_volume_load_on_an_element_subset(cell):
    """Only the condition's elements contribute; the Gauss-point factors are indexed by the position in the
    condition's element list."""
    rm, _ = mesh(cell)
    X, conn = np.asarray(rm.X, float), np.asarray(rm.conn, np.int32)
    rowptr, _ = orc.build_graph(CELLS[cell], conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    vn = {"ONOFF": [1, 0, 1], "VAL": [2.0, 7.0, -0.3], "FUNCT": [2, 0, 1]}
    ids = np.array([len(conn) - 1, 0, 3], np.int32)
    slot = np.full(len(conn), -1, np.int32)
    slot[ids] = np.arange(len(ids))
    sub = {"celltype": cell, "coords": X.tolist(), "elements": conn[ids].tolist(), "vol_neumann": vn}
    t = 0.25
    ref = nm.volume_neumann_load(sub, t)
    fac = _volume_gp_factors(cell, sub, conn[ids], X, t)
    f = emu.neumann_volume(CELLS[cell], conn, X.ravel(), vn["ONOFF"], vn["VAL"], rowptr, rm.node_row_lid,
                           rm.node_col_lid, 3 * rm.n_row_nodes, ele_slot=slot, gp_fac=fac)
    assert np.abs(ref).max() > 0 and np.all(f[1::3] == 0.0)
    assert np.abs(f - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("cell", ["hex8", "hex27"])
def test_product_loads_owner_computes_on_a_partition(cell):
    """Every rank evaluates the faces of its column elements and keeps the rows it owns
    (Discretization::evaluate_neumann + assemble with lmowner): the rank results add up to the serial one."""
    rm0, u0 = mesh(cell)
    case0 = synthetic_case(cell, rm0, top)
    ref_f, ref_K = nm.orthopressure_load(case0, u0, 0.5)
    ref_K = ref_K.toarray()
    val = [case0["neumann"]["VAL"][0] * nm.time_factor(case0, 0.5), 0, 0]
    tot_f = np.zeros_like(ref_f)
    tot_K = np.zeros_like(ref_K)
    serial_of = {int(g): k for k, g in enumerate(rm0.node_gid)}
    for rank in range(4):
        rm, u = mesh(cell, rank, 4)
        case = synthetic_case(cell, rm, top)
        faces = nm.loaded_faces(case)
        rowptr, colind = orc.build_graph(CELLS[cell], rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
        f, K, _, flag = emu.neumann_surface(CELLS[cell], 2, faces, rm.X.ravel(), u, [1, 0, 0], val, rowptr, colind,
                                            rm.node_row_lid, rm.node_col_lid, 3 * rm.n_row_nodes, loadlin=True)
        assert flag == 0
        gl = np.array([serial_of[int(g)] for g in rm.node_gid])   # column lid -> serial node
        col_node = np.empty(rm.n_col_nodes, np.int64)
        col_node[rm.node_col_lid] = gl
        for k in np.nonzero(rm.node_row_lid >= 0)[0]:
            r0 = 3 * rm.node_row_lid[k]
            tot_f[3 * gl[k]:3 * gl[k] + 3] += f[r0:r0 + 3]
            for d in range(3):
                sl = slice(rowptr[r0 + d], rowptr[r0 + d + 1])
                cols = 3 * col_node[colind[sl] // 3] + colind[sl] % 3
                np.add.at(tot_K[3 * gl[k] + d], cols, K[sl])
    assert np.abs(tot_f - ref_f).max() <= 1e-13 * np.abs(ref_f).max()
    assert np.abs(tot_K - ref_K).max() <= 1e-13 * np.abs(ref_K).max()


def test_face_outside_the_graph_is_flagged():
    rm, u = mesh("hex8")
    rowptr, colind = orc.build_graph(orc.HEX8, rm.conn, rm.node_row_lid, rm.node_col_lid, rm.n_row_nodes)
    X = np.asarray(rm.X)
    far = [int(np.argmin(np.abs(X - c).sum(axis=1))) for c in ([0, 0, 0], [1, 0, 0], [1, 1, 1], [0, 1, 1])]
    _, _, _, flag = emu.neumann_surface(orc.HEX8, 2, [far], rm.X.ravel(), u, [1, 0, 0], [1.0, 0, 0], rowptr, colind,
                                        rm.node_row_lid, rm.node_col_lid, 3 * rm.n_row_nodes, loadlin=True)
    assert flag == 1

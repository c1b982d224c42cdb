"""Test-side restatement of the reference's 'Live' surface Neumann load, needed only to reproduce
the RESULT DESCRIPTION known answers of the reference's integration tests with the oracle.

Follows src/solid_ele/4C_solid_ele_surface_evaluate.cpp:43-306 (TYPE Live: dead load on the
material configuration) with the surface Gauss rules of
src/core/fem/src/general/utils/4C_fem_general_utils_integration.cpp (quad_4point :6151-6183,
quad_9point :6185-6233, tri_6point :6374-6397; truncated decimals taken verbatim) and the face
numbering of 4C_fem_general_utils_local_connectivity_matrices.hpp:51-63.
"""
import numpy as np

HEX8_FACES = [[0, 3, 2, 1], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [4, 5, 6, 7]]
HEX27_FACES = [[0, 3, 2, 1, 11, 10, 9, 8, 20], [0, 1, 5, 4, 8, 13, 16, 12, 21],
               [1, 2, 6, 5, 9, 14, 17, 13, 22], [2, 3, 7, 6, 10, 15, 18, 14, 23],
               [0, 4, 7, 3, 12, 19, 15, 11, 24], [4, 5, 6, 7, 16, 17, 18, 19, 25]]
TET10_FACES = [[0, 1, 3, 4, 8, 7], [1, 2, 3, 5, 9, 8], [0, 3, 2, 7, 9, 6], [0, 2, 1, 6, 5, 4]]
TET4_FACES = [f[:3] for f in TET10_FACES]     # tri3: the corners of the tri6 faces
HEX20_FACES = [f[:8] for f in HEX27_FACES]    # quad8: corners + mid-edge nodes of the quad9 faces
# wedge6: three quad4 then two tri3 surfaces, pyramid5: the quad4 base then four tri3 surfaces
# (4C_fem_general_utils_local_connectivity_matrices.hpp:68-86, .cpp:697-765: first corners of the wedge18 tables)
WEDGE6_FACES = [[0, 1, 4, 3], [1, 2, 5, 4], [2, 0, 3, 5], [2, 1, 0], [5, 3, 4]]
PYRAMID5_FACES = [[0, 3, 2, 1], [0, 1, 4], [1, 2, 4], [2, 3, 4], [0, 4, 3]]
# hex18: two quad9 surfaces (bottom, top), then four quad6 side surfaces (:36-43)
HEX18_FACES = [[0, 3, 2, 1, 7, 6, 5, 4, 8], [9, 10, 11, 12, 13, 14, 15, 16, 17],
               [0, 1, 4, 9, 10, 13], [1, 2, 5, 10, 11, 14], [2, 3, 6, 11, 12, 15], [3, 0, 7, 12, 9, 16]]
FACES = {"hex18": HEX18_FACES, "hex8": HEX8_FACES, "hex27": HEX27_FACES, "tet10": TET10_FACES, "tet4": TET4_FACES, "hex20": HEX20_FACES,
         "wedge6": WEDGE6_FACES, "pyramid5": PYRAMID5_FACES}


def _quad4(r, s):
    sg = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float)
    N = 0.25 * (1 + sg[:, 0] * r) * (1 + sg[:, 1] * s)
    dN = np.stack([0.25 * sg[:, 0] * (1 + sg[:, 1] * s), 0.25 * sg[:, 1] * (1 + sg[:, 0] * r)])
    return N, dN


def _quad9(r, s):
    def lag(x):
        return (np.array([0.5 * x * (x - 1), 1 - x * x, 0.5 * x * (x + 1)]),
                np.array([x - 0.5, -2 * x, x + 0.5]))
    Lr, dLr = lag(r)
    Ls, dLs = lag(s)
    pos = [(-1, -1), (1, -1), (1, 1), (-1, 1), (0, -1), (1, 0), (0, 1), (-1, 0), (0, 0)]
    N = np.array([Lr[a + 1] * Ls[b + 1] for a, b in pos])
    dN = np.array([[dLr[a + 1] * Ls[b + 1] for a, b in pos], [Lr[a + 1] * dLs[b + 1] for a, b in pos]])
    return N, dN


def _tri6(r, s):
    t1, t2, t3 = 1 - r - s, r, s
    N = np.array([t1 * (2 * t1 - 1), t2 * (2 * t2 - 1), t3 * (2 * t3 - 1), 4 * t2 * t1, 4 * t2 * t3,
                  4 * t3 * t1])
    dN = np.array([[-(4 * t1 - 1), 4 * t2 - 1, 0, 4 * (t1 - t2), 4 * t3, -4 * t3],
                   [-(4 * t1 - 1), 0, 4 * t3 - 1, -4 * t2, 4 * t2, 4 * (t1 - t3)]])
    return N, dN


def _tri3(r, s):
    N = np.array([1 - r - s, r, s])
    dN = np.array([[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]])
    return N, dN


def _quad8(r, s):
    """8-node serendipity quadrilateral (4C_fem_general_utils_fem_shapefunctions.hpp:2049-2086): corner k
    1/4 [(1 + r_k r)(1 + s_k s) - (1 - r^2)(1 + s_k s) - (1 - s^2)(1 + r_k r)], mid-side nodes 1/2 (1 - r^2)(1 + s_k s) ..."""
    c = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
    N, d0, d1 = np.zeros(8), np.zeros(8), np.zeros(8)
    r2, s2 = 1 - r * r, 1 - s * s
    for k, (a, b) in enumerate(c):
        fr, fs = 1 + a * r, 1 + b * s
        N[k] = 0.25 * (fr * fs - (r2 * fs + s2 * fr))
        d0[k] = 0.25 * (a * fs - (-2 * r * fs + s2 * a))
        d1[k] = 0.25 * (fr * b - (r2 * b - 2 * s * fr))
    mids = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    for k, (a, b) in enumerate(mids):
        if a == 0:
            fs = 1 + b * s
            N[4 + k], d0[4 + k], d1[4 + k] = 0.5 * r2 * fs, -r * fs, 0.5 * r2 * b
        else:
            fr = 1 + a * r
            N[4 + k], d0[4 + k], d1[4 + k] = 0.5 * s2 * fr, 0.5 * s2 * a, -s * fr
    return N, np.stack([d0, d1])


def _quad6(r, s):
    """6-node quadrilateral, quadratic in r and linear in s (4C_fem_general_utils_fem_shapefunctions.hpp:2038-2047,
    2204-2220): nodes (-1,-1), (1,-1), (0,-1), (-1,1), (1,1), (0,1)."""
    q = np.array([0.5 * r * (r - 1), 0.5 * r * (r + 1), -(r - 1) * (r + 1)])
    dq = np.array([r - 0.5, r + 0.5, -2 * r])
    N = np.concatenate([q * 0.5 * (1 - s), q * 0.5 * (1 + s)])
    dN = np.array([np.concatenate([dq * 0.5 * (1 - s), dq * 0.5 * (1 + s)]), np.concatenate([-0.5 * q, 0.5 * q])])
    return N, dN


def _rule(kind):
    if kind == "quad6":  # quad_6point (4C_fem_general_utils_integration.cpp:6235-6261)
        g, d, w1, w2 = 0.7745966692415, 0.5773502691896, 0.5555555555556, 0.8888888888889
        return (np.array([[-g, -d], [g, -d], [0, -d], [-g, d], [g, d], [0, d]]), np.array([w1, w1, w2, w1, w1, w2]))
    if kind == "tri3":   # tri_3point (4C_fem_general_utils_integration.cpp:6341-6355)
        return np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.full(3, 1 / 6)
    if kind == "quad8":  # get_optimal_gauss_rule(quad8) = quad_9point
        return _rule("quad9")
    if kind == "quad4":
        g = 0.5773502691896
        return np.array([[-g, -g], [g, -g], [g, g], [-g, g]]), np.ones(4)
    if kind == "quad9":
        g, w1, w2 = 0.7745966692415, 0.5555555555556, 0.8888888888889
        x = np.array([[-g, -g], [g, -g], [g, g], [-g, g], [0, -g], [g, 0], [0, g], [-g, 0], [0, 0]])
        w = np.array([w1 * w1] * 4 + [w2 * w1, w1 * w2, w2 * w1, w1 * w2, w2 * w2])
        return x, w
    a, b, c, d = 0.0915762135098, 0.8168475729805, 0.4459484909160, 0.1081030181681
    x = np.array([[a, a], [b, a], [a, b], [c, d], [c, c], [d, c]])
    w = np.array([0.0549758718277] * 3 + [0.1116907948390] * 3)
    return x, w


def _funct(fid, gp, time):
    if fid == 1:
        return time
    if fid == 2:
        return (gp[0] + 1) * (gp[1] + 2) * (gp[2] + 3)
    return 1.0


def volume_neumann_load(case, time):
    """DESIGN VOL NEUMANN CONDITIONS: dead body load integrated with the stiffness Gauss rule, functions
    evaluated at the reference coordinates of the Gauss point
    (src/solid_ele/4C_solid_ele_neumann_evaluator.cpp:43-123)."""
    from oracle import oracle as orc
    ct = {"hex8": orc.HEX8, "hex27": orc.HEX27, "tet10": orc.TET10, "tet4": orc.TET4, "hex20": orc.HEX20,
          "wedge6": orc.WEDGE6, "pyramid5": orc.PYRAMID5, "hex18": orc.HEX18}[case["celltype"]]
    X = np.asarray(case["coords"], float)
    vn = case["vol_neumann"]
    gxi, gw = orc.gauss(ct)
    f = np.zeros(3 * len(X))
    for conn in case["elements"]:
        x = X[conn]
        for xi, w in zip(gxi, gw):
            N, dN = orc.shape(ct, xi)          # dN[a, k]
            detJ = np.linalg.det(dN @ x)
            gp = N @ x
            for dof in range(3):
                if vn["ONOFF"][dof]:
                    fac = vn["VAL"][dof] * _funct(vn["FUNCT"][dof], gp, time) * detJ * w
                    for a, nid in enumerate(conn):
                        f[3 * nid + dof] += N[a] * fac
    return f


def neumann_load(case, time):
    """External force vector (3*nnode) of the golden case at total time `time`."""
    if case.get("vol_neumann"):
        return _surface_neumann_load(case, time) + volume_neumann_load(case, time)
    return _surface_neumann_load(case, time)


def _surface_neumann_load(case, time):
    cell = case["celltype"]
    X = np.asarray(case["coords"], float)
    nset = set(case["neumann_nodes"])
    kinds = {"hex8": ("quad4", _quad4), "hex27": ("quad9", _quad9), "tet10": ("tri6", _tri6),
             "tet4": ("tri3", _tri3), "hex20": ("quad8", _quad8)}
    val, onoff, funct = case["neumann"]["VAL"], case["neumann"]["ONOFF"], case["neumann"]["FUNCT"]
    assert case["neumann"]["FUNCT1"] == "t"
    assert case["neumann"]["FUNCT2"] in ("(x+1)*(y+2)*(z+3)", None)
    f = np.zeros(3 * len(X))
    for conn in case["elements"]:
        for face in FACES[cell]:
            ids = [conn[k] for k in face]
            if not all(i in nset for i in ids):
                continue
            x = X[ids]
            if cell == "hex18":
                kind, fn = ("quad9", _quad9) if len(face) == 9 else ("quad6", _quad6)
            else:
                kind, fn = kinds[cell] if cell in kinds else (("tri3", _tri3) if len(face) == 3 else ("quad4", _quad4))
            gx, gw = _rule(kind)
            for (r, s), w in zip(gx, gw):
                N, dN = fn(r, s)
                dx = dN @ x  # (2,3)
                m = dx @ dx.T
                detA = np.sqrt(m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0])
                gp = N @ x
                for dof in range(3):
                    if not onoff[dof]:
                        continue
                    if funct[dof] == 1:
                        ff = time
                    elif funct[dof] == 2:
                        ff = (gp[0] + 1) * (gp[1] + 2) * (gp[2] + 3)
                    else:
                        ff = 1.0
                    fac = w * detA * val[dof] * ff
                    for a, nid in enumerate(ids):
                        f[3 * nid + dof] += N[a] * fac
    return f


def loaded_faces(case):
    """Faces of the condition geometry: every element surface whose nodes all belong to the condition's
    node set, in the parent's surface node order (the boundary-element factory of the reference)."""
    cell = case["celltype"]
    nset = set(case["neumann_nodes"])
    out = []
    for conn in case["elements"]:
        for face in FACES[cell]:
            ids = [conn[k] for k in face]
            if all(i in nset for i in ids):
                out.append(ids)
    return np.asarray(out, np.int32)


def time_factor(case, time):
    """FUNCT1 of the follower-load case: one linearinterpolation variable."""
    n = case["neumann"]
    return float(np.interp(time, n["FUNCT1_TIMES"], n["FUNCT1_VALUES"]))


def orthopressure_load(case, u, time, loadlin=True, faces=None, material=False):
    """TYPE orthopressure (4C_solid_ele_surface_evaluate.cpp:319-378): follower pressure on the spatial
    configuration x = X + u, force N_a * (g1 x g2) * w * funct * VAL[0], and with LOADLIN the load
    stiffness -N_a * d(g1 x g2)/dx * fac (analytical_d_surface_integration :502-547) as a scipy CSR matrix.
    material=True: the configuration is the material one and there is no linearisation (MULF active,
    :158-166; also what pseudo_orthopressure does on the last converged state when u is that state)."""
    cell = case["celltype"]
    X = np.asarray(case["coords"], float)
    x_all = X if material else X + np.asarray(u, float).reshape(-1, 3)
    kind, fn = {"hex8": ("quad4", _quad4), "hex27": ("quad9", _quad9), "tet10": ("tri6", _tri6),
                "tet4": ("tri3", _tri3), "hex20": ("quad8", _quad8)}[cell]
    gx, gw = _rule(kind)
    n = case["neumann"]
    assert n["ONOFF"][0] == 1 and not any(n["ONOFF"][1:3]), "orthopressure on 1st dof only!"
    val = n["VAL"][0] * (time_factor(case, time) if n["FUNCT"][0] else 1.0)
    faces = loaded_faces(case) if faces is None else faces
    nd = 3 * len(X)
    f = np.zeros(nd)
    lin = loadlin and not material
    ri, ci, vi = [], [], []
    eps = np.zeros((3, 3, 3))  # permutation symbol: (a x b)_i = eps_ijk a_j b_k
    eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1.0
    eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1.0
    for ids in faces:
        x = x_all[ids]
        dofs = np.array([3 * i + d for i in ids for d in range(3)])
        Ke = np.zeros((len(dofs), len(dofs)))
        for (r, s), w in zip(gx, gw):
            N, dN = fn(r, s)
            g1, g2 = dN[0] @ x, dN[1] @ x
            normal = np.cross(g1, g2)
            fac = w * val
            for a, nid in enumerate(ids):
                f[3 * nid:3 * nid + 3] += N[a] * normal * fac
            if lin:
                # d(g1 x g2)_dim / d x_{b,k} = dN0_b (e_k x g2)_dim + dN1_b (g1 x e_k)_dim
                A2 = np.einsum("ikm,m->ik", eps, g2)   # (e_k x g2)_i
                A1 = np.einsum("imk,m->ik", eps, g1)   # (g1 x e_k)_i
                Dn = (dN[0][None, :, None] * A2[:, None, :] + dN[1][None, :, None] * A1[:, None, :])
                Dn = Dn.reshape(3, len(dofs))
                Ke -= np.einsum("a,dj->adj", N, Dn).reshape(len(dofs), len(dofs)) * fac
        if lin:
            ri.append(np.repeat(dofs, len(dofs)))
            ci.append(np.tile(dofs, len(dofs)))
            vi.append(Ke.ravel())
    if not lin:
        return f, None
    import scipy.sparse as sp
    K = sp.coo_matrix((np.concatenate(vi), (np.concatenate(ri), np.concatenate(ci))), shape=(nd, nd)).tocsr()
    return f, K


def add_to_csr_values(K, rowptr, colind, values):
    """values + the entries of the sparse matrix K at the CSR pattern (K's pattern must be contained)."""
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    add = np.asarray(K[rows, colind]).ravel()
    assert abs(add).sum() >= abs(K).sum() * (1 - 1e-12)
    return values + add


def point_neumann_load(case, time):
    """DESIGN POINT NEUMANN CONDITIONS (Discretization::evaluate_neumann, 4C_fem_discretization_evaluate.cpp:
    205-256): VAL * FUNCT(t) on every node of the condition; FUNCT1 = t."""
    pn = case["point_neumann"]
    f = np.zeros(3 * len(case["coords"]))
    for n in pn["nodes"]:
        for d in range(3):
            if pn["ONOFF"][d]:
                f[3 * n + d] += pn["VAL"][d] * (time if pn["FUNCT"][d] == 1 else 1.0)
    return f

"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes/numpy front end of the CPU oracle (oracle/liboracle.so, built from solid_oracle.cpp) and,
when present, of oracle/_ref/libref_solid.so (built on the reference's own headers).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

HEX8, HEX27, TET10, TET4, HEX20, WEDGE6, PYRAMID5, HEX18 = 0, 1, 2, 3, 4, 5, 6, 7
MAT_COUP_NEOHOOKE, MAT_ISO_VOL, MAT_SVK, MAT_ELASTHYPER, MAT_DRUCKER_PRAGER, MAT_VISCONEOHOOKE = 0, 1, 2, 3, 4, 5
MAT_PLASTIC_NLNLOG = 6
PLASTIC_HIST = 14  # {accplstrainlast, invplrcglast[6], accplstraincurr, invplrcgcurr[6]} per Gauss point
DP_HIST = 14  # {strainpllast[6], strainbarpllast, strainplcurr[6], strainbarplcurr} per Gauss point
VISCO_HIST = 24  # {histstresslast[6], artstresslast[6], histstresscurr[6], artstresscurr[6]} per Gauss point
# summand ids of the generic MAT_ElastHyper (ORC_SUM_* in solid_oracle.h)
SUMMANDS = ["COUP_NEOHOOKE", "COUP_MOONEYRIVLIN", "COUP_LOGNEOHOOKE", "COUP_LOGMIXNEOHOOKE", "COUP_SVK",
            "COUP_SIMOPISTER", "COUP_BLATZKO", "COUP_EXPPOL", "COUP_1POW", "COUP_2POW", "COUP_3POW",
            "COUP_13APOW", "ISO_NEOHOOKE", "ISO_MOONEYRIVLIN", "ISO_YEOH", "ISO_1POW", "ISO_2POW",
            "ISO_EXPOPOW", "VOL_SUSSMANBATHE", "VOL_PENALTY", "VOL_OGDEN", "VOL_POW"]
SUM = {name: i for i, name in enumerate(SUMMANDS)}
SUM.update({"ISO_OGDEN": 23, "COUP_VARGA": 24, "ISO_VARGA": 25})  # principal-stretch summands
SUM.update({"COUP_ANISOEXPO": 100, "FIBER": 101})  # oracle only, see solid_oracle.h
STRESS_NONE, STRESS_2PK, STRESS_CAUCHY = 0, 1, 2
STRAIN_NONE, STRAIN_GL, STRAIN_EA = 0, 1, 2
ACTION_FORCE, ACTION_FORCE_STIFF = 0, 1
TECH_NONE, TECH_FBAR, TECH_MULF, TECH_EAS_MILD, TECH_EAS_FULL, TECH_LINEAR, TECH_MULF_FBAR = 0, 1, 2, 3, 4, 5, 6
TECH_EAS_MILD_LINEAR, TECH_EAS_FULL_LINEAR = 7, 8  # hex8 EAS with KINEM linear
TECH_SHELL_ANS = 9  # hex8 / wedge6 solid shell with assumed natural strains
TECH_SHELL_EAS = 10  # hex8 EAS with the 7 solid-shell modes (eastype_sh8_7)
TECH_SHELL_EAS_ANS = 11  # hex8 solid shell: ANS + the 7 EAS modes
MULF_HIST = 19

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(_lp) if a is not None else None


def build(force=False):
    """Compile liboracle.so (always) and _ref/libref_solid.so (only where /root/reference exists)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "solid_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libref_solid.so")
    if os.path.isdir("/root/reference/src"):
        rsrc = os.path.join(_HERE, "ref_driver.cpp")
        if force or not os.path.exists(ref) or os.path.getmtime(ref) < os.path.getmtime(rsrc):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.orc_svk_strain_energy.restype = C.c_double
        L.orc_evaluate_assemble.restype = C.c_int64
        L.orc_evaluate_assemble.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int64, _ip,
                                            C.c_int64, _dp, _dp, _ip, _ip, C.c_int64, _lp, _ip, _dp,
                                            _dp, C.c_int]
        L.orc_evaluate_assemble_tech.restype = C.c_int64
        L.orc_evaluate_assemble_tech.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                                 C.c_int64, _ip, C.c_int64, _dp, _dp, _ip, _ip,
                                                 C.c_int64, _lp, _ip, _dp, _dp, C.c_int]
        L.orc_evaluate_assemble_hist.restype = C.c_int64
        L.orc_evaluate_assemble_hist.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int,
                                                 C.c_int64, _ip, C.c_int64, _dp, _dp, _ip, _ip,
                                                 C.c_int64, _lp, _ip, _dp, _dp, C.c_int, _dp]
        L.orc_set_output_history.argtypes = [_dp]
        L.orc_set_material_history.argtypes = [_dp]
        L.orc_set_element_material_history.argtypes = [_dp]
        L.orc_material_hist.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.orc_material_update.argtypes = [C.c_int, C.c_int64, _dp]
        L.orc_eas_hist_size.restype = C.c_int64
        L.orc_eas_recover.argtypes = [C.c_int, C.c_int64, _ip, _dp, C.c_double, _dp]
        L.orc_mulf_init_history.argtypes = [C.c_int64, _dp]
        L.orc_update_prestress.argtypes = [C.c_int, C.c_int64, _ip, _dp, _dp, _dp]
        L.orc_update_prestress_fbar.argtypes = [C.c_int, C.c_int64, _ip, _dp, _dp, _dp]
        L.orc_element_tech.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_build_graph.argtypes = [C.c_int, C.c_int64, _ip, C.c_int64, _ip, _ip, C.c_int64, _lp,
                                      _ip]
        L.orc_element.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp]
        L.orc_element_stress.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
        L.orc_element_energy.restype = C.c_double
        L.orc_element_energy.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_strain_energy.restype = C.c_double
        L.orc_strain_energy.argtypes = [C.c_int, _dp, _dp]
        L.orc_element_mass.argtypes = [C.c_int, C.c_double, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_calc_stress.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int64, _ip, _dp, _dp,
                                      C.c_int, C.c_int, _dp, _dp]
        L.orc_calc_energy.restype = C.c_double
        L.orc_calc_energy.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int64, _ip, _dp, _dp,
                                      C.POINTER(C.c_uint8)]
        L.orc_assemble_mass.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int64, _ip, _dp, _dp, _ip, _ip,
                                        C.c_int64, _lp, _ip, _dp, _dp]
        L.orc_invariant_derivatives.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        for fn in (L.orc_stiffness_bfree, L.orc_stiffness_bop):
            fn.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp]
        _lib = L
    return _lib


def ref_lib():
    """oracle/_ref (reference headers) or None when it was never built (e.g. no /root/reference)."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libref_solid.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_element.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp, C.c_int, _dp,
                                  _dp]
        R.ref_stiffness_bfree.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp]
        R.ref_evaluate_assemble.restype = C.c_int64
        R.ref_evaluate_assemble.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int64, _ip, _dp,
                                            _dp, _ip, _ip, C.c_int64, _lp, _ip, _dp, _dp, C.c_int,
                                            _dp, _dp, C.c_int]
        _ref = R
    return _ref


def num_nodes(ct):
    return lib().orc_num_nodes(ct)


def num_gp(ct):
    return lib().orc_num_gp(ct)


def shape(ct, xi):
    n = num_nodes(ct)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    N = np.zeros(n)
    dN = np.zeros(3 * n)
    lib().orc_shape(ct, _d(xi), _d(N), _d(dN))
    return N, dN.reshape(n, 3).T.copy()  # dN[a,k]


def gauss(ct):
    g = num_gp(ct)
    xi = np.zeros(3 * g)
    w = np.zeros(g)
    lib().orc_gauss(ct, _d(xi), _d(w))
    return xi.reshape(g, 3), w


def node_param_coords(ct):
    n = num_nodes(ct)
    xi = np.zeros(3 * n)
    lib().orc_node_param_coords(ct, _d(xi))
    return xi.reshape(n, 3)


def material(kind, params, E, F=None):
    params = np.ascontiguousarray(params, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    F = np.eye(3).T.copy().ravel() if F is None else np.ascontiguousarray(F, dtype=np.float64)
    S = np.zeros(6)
    Cm = np.zeros(36)
    lib().orc_material(kind, _d(params), _d(F), _d(E), _d(S), _d(Cm))
    return S, Cm.reshape(6, 6).T.copy()  # Cm[I,J]


def drucker_prager_params(YOUNG, NUE, ISOHARD, TOL, C, ETA, XI, ETABAR, MAXITER=50, TANG="consistent", **_):
    """Parameter block of MAT_DRUCKER_PRAGER from the keys of MAT_Struct_DruckerPrager."""
    return np.array([YOUNG, NUE, ISOHARD, TOL, C, ETA, XI, ETABAR, MAXITER, 0.0 if TANG == "elastic" else 1.0], float)


def material_hist(kind, params, E, hist=None):
    """History material at one point: E tensor components [00,11,22,01,12,02]; hist (DP_HIST,) updated in place."""
    params = np.ascontiguousarray(params, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    S, Cm = np.zeros(6), np.zeros(36)
    lib().orc_material_hist(kind, _d(params), _d(E), _d(hist), _d(S), _d(Cm))
    return S, Cm.reshape(6, 6).T.copy()


def material_update(kind, hist):
    """Mat::...::update(): current -> last for every Gauss point of hist (n, DP_HIST or VISCO_HIST)."""
    per = VISCO_HIST if kind == MAT_VISCONEOHOOKE else (PLASTIC_HIST if kind == MAT_PLASTIC_NLNLOG else DP_HIST)
    lib().orc_material_update(kind, C.c_int64(hist.size // per), _d(hist))


def visco_neohooke_params(YOUNGS_SLOW, POISSON, DENS, YOUNGS_FAST, RELAX, THETA, dt, **_):
    """Parameter block of MAT_VISCONEOHOOKE from the keys of the input line + the time step size."""
    return np.array([YOUNGS_SLOW, POISSON, DENS, YOUNGS_FAST, RELAX, THETA, dt], float)


def plastic_nlnlog_params(YOUNG, NUE, DENS, YIELD, ISOHARD, SATHARDENING, HARDEXPO, VISC, RATE_DEPENDENCY, TOL, dt, **_):
    """Parameter block of MAT_PLASTIC_NLNLOG from the keys of MAT_Struct_PlasticNlnLogNeoHooke + the time step size."""
    return np.array([YOUNG, NUE, DENS, YIELD, ISOHARD, SATHARDENING, HARDEXPO, VISC, RATE_DEPENDENCY, TOL, dt], float)


def plastic_new_history(*shape):
    """Fresh history of MAT_PLASTIC_NLNLOG (PlasticNlnLogNeoHooke::setup): (*shape, PLASTIC_HIST) with Cp^-1 = I."""
    h = np.zeros(tuple(shape) + (PLASTIC_HIST,))
    h[..., 1:4] = 1.0
    h[..., 8:11] = 1.0
    return h


def set_material_history(mat_hist):
    """History store of a finite-strain history material for the element loops: (n_ele, ngp, VISCO_HIST) float64,
    C-contiguous, updated in place; None switches it off.  Keep the array alive while it is installed."""
    if mat_hist is not None:
        assert mat_hist.dtype == np.float64 and mat_hist.flags["C_CONTIGUOUS"]
    lib().orc_set_material_history(_d(mat_hist))


def set_element_material_history(mat_hist):
    """Same for single-element calls (element / element_tech): (ngp, VISCO_HIST)."""
    if mat_hist is not None:
        assert mat_hist.dtype == np.float64 and mat_hist.flags["C_CONTIGUOUS"]
    lib().orc_set_element_material_history(_d(mat_hist))


def element(ct, matkind, params, X, u, action=ACTION_FORCE_STIFF, tech=TECH_NONE):
    """X,u: (n,3). Returns (Ke (nd,nd) with Ke[row,col], fe, err)."""
    n = num_nodes(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 3)
    Ke = np.zeros((3 * n) ** 2)
    fe = np.zeros(3 * n)
    err = lib().orc_element_tech(ct, tech, matkind, _d(params), action, _d(X), _d(u), _d(Ke), _d(fe))
    return Ke.reshape(3 * n, 3 * n).T.copy(), fe, err


def ref_element(ct, matkind, params, X, u, action=ACTION_FORCE_STIFF):
    R = ref_lib()
    n = num_nodes(ct)
    gxi, gw = gauss(ct)
    gxi = np.ascontiguousarray(gxi)
    params = np.ascontiguousarray(params, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 3)
    Ke = np.zeros((3 * n) ** 2)
    fe = np.zeros(3 * n)
    err = R.ref_element(ct, matkind, _d(params), action, _d(X), _d(u), _d(gxi), _d(gw), len(gw),
                        _d(Ke), _d(fe))
    return Ke.reshape(3 * n, 3 * n).T.copy(), fe, err


def stiffness_point(ct, X, u, xi, S, Cm, fac=1.0, which="bfree"):
    n = num_nodes(ct)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 3)
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    S = np.ascontiguousarray(S, dtype=np.float64)
    Cc = np.ascontiguousarray(np.asarray(Cm, dtype=np.float64).T).ravel()  # -> column-major
    Ke = np.zeros((3 * n) ** 2)
    if which == "ref":
        ref_lib().ref_stiffness_bfree(ct, _d(X), _d(u), _d(xi), _d(S), _d(Cc), fac, _d(Ke))
    else:
        fn = lib().orc_stiffness_bfree if which == "bfree" else lib().orc_stiffness_bop
        fn(ct, _d(X), _d(u), _d(xi), _d(S), _d(Cc), fac, _d(Ke))
    return Ke.reshape(3 * n, 3 * n).T.copy()


def build_graph(ct, conn, node_row_lid, node_col_lid, n_row_nodes):
    n = num_nodes(ct)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    node_row_lid = np.ascontiguousarray(node_row_lid, dtype=np.int32)
    node_col_lid = np.ascontiguousarray(node_col_lid, dtype=np.int32)
    rowptr = np.zeros(3 * n_row_nodes + 1, dtype=np.int64)
    lib().orc_build_graph(ct, conn.shape[0], _i(conn), len(node_row_lid), _i(node_row_lid),
                          _i(node_col_lid), n_row_nodes, _l(rowptr), None)
    colind = np.zeros(rowptr[-1], dtype=np.int32)
    lib().orc_build_graph(ct, conn.shape[0], _i(conn), len(node_row_lid), _i(node_row_lid),
                          _i(node_col_lid), n_row_nodes, _l(rowptr), _i(colind))
    return rowptr, colind


def evaluate_assemble(ct, matkind, params, conn, X, u_col, node_row_lid, node_col_lid, n_row_nodes,
                      rowptr, colind, action=ACTION_FORCE_STIFF, mue_per_ele=None, nthreads=1,
                      impl="oracle", tech=TECH_NONE, hist=None):
    """impl="oracle": plain restatement; impl="ref": element arithmetic by the reference's headers."""
    n = num_nodes(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    u_col = np.ascontiguousarray(u_col, dtype=np.float64)
    node_row_lid = np.ascontiguousarray(node_row_lid, dtype=np.int32)
    node_col_lid = np.ascontiguousarray(node_col_lid, dtype=np.int32)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    if mue_per_ele is not None:
        mue_per_ele = np.ascontiguousarray(mue_per_ele, dtype=np.float64)
    values = np.zeros(rowptr[-1]) if action == ACTION_FORCE_STIFF else None
    f = np.zeros(3 * n_row_nodes)
    if impl == "ref":
        assert tech == TECH_NONE, "oracle/_ref covers the displacement-based formulation only"
        # ref_evaluate_assemble sizes its element arrays for these three cells (ref_driver.cpp); the other cells have
        # element-level entries only (ref_element)
        assert ct in (HEX8, HEX27, TET10), "oracle/_ref assembles hex8 / hex27 / tet10 blocks only"
        gxi, gw = gauss(ct)
        gxi = np.ascontiguousarray(gxi)
        nbad = ref_lib().ref_evaluate_assemble(ct, matkind, _d(params), _d(mue_per_ele), action,
                                               conn.shape[0], _i(conn), _d(X), _d(u_col),
                                               _i(node_row_lid), _i(node_col_lid), n_row_nodes,
                                               _l(rowptr), _i(colind), _d(gxi), _d(gw), len(gw),
                                               _d(values), _d(f), nthreads)
        return values, f, int(nbad)
    if hist is not None:
        assert hist.dtype == np.float64 and hist.flags["C_CONTIGUOUS"] and nthreads == 1
    nbad = lib().orc_evaluate_assemble_hist(ct, tech, matkind, _d(params), _d(mue_per_ele), action,
                                            conn.shape[0], _i(conn), len(node_row_lid), _d(X), _d(u_col),
                                            _i(node_row_lid), _i(node_col_lid), n_row_nodes, _l(rowptr),
                                            _i(colind), _d(values), _d(f), nthreads, _d(hist))
    assert nbad >= 0, "ORC_TECH_MULF needs a history array"
    return values, f, int(nbad)


def pack_elasthyper(summands):
    """summands: list of (name or id, [params...]) or (name, [params...], mode) -> packed parameter block of
    MAT_ELASTHYPER (solid_oracle.cpp hyper_generic): {n, n x {kind, mode, p0..p3}}."""
    out = [float(len(summands))]
    for sm in summands:
        kind, par = sm[0], list(sm[1])
        mode = sm[2] if len(sm) > 2 else 0
        kind = SUM[kind] if isinstance(kind, str) else int(kind)
        out += [float(kind), float(mode)] + [float(v) for v in par] + [0.0] * (4 - len(par))
    return np.asarray(out, dtype=np.float64)


def invariant_derivatives(kind, params, prinv):
    params = np.ascontiguousarray(params, dtype=np.float64)
    prinv = np.ascontiguousarray(prinv, dtype=np.float64)
    dPI, ddPII = np.zeros(3), np.zeros(6)
    lib().orc_invariant_derivatives(kind, _d(params), _d(prinv), _d(dPI), _d(ddPII))
    return dPI, ddPII


def strain_energy(kind, params, E):
    params = np.ascontiguousarray(params, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    return lib().orc_strain_energy(kind, _d(params), _d(E))


def element_stress(ct, matkind, params, X, u, stress_type=STRESS_2PK, strain_type=STRAIN_GL, tech=TECH_NONE):
    n, g = num_nodes(ct), num_gp(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 3)
    st, en = np.zeros((g, 6)), np.zeros((g, 6))
    lib().orc_element_stress(ct, tech, matkind, _d(params), _d(X), _d(u), stress_type, strain_type,
                             _d(st), _d(en))
    return st, en


def element_energy(ct, matkind, params, X, u, tech=TECH_NONE):
    n = num_nodes(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(n, 3)
    return lib().orc_element_energy(ct, tech, matkind, _d(params), _d(X), _d(u))


def element_mass(ct, density, X, lumped=False, acc=None):
    """Returns (Me[row, col], f_inertia or None)."""
    n = num_nodes(ct)
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(n, 3)
    Me = np.zeros((3 * n) ** 2)
    fi = np.zeros(3 * n) if acc is not None else None
    acc = None if acc is None else np.ascontiguousarray(acc, dtype=np.float64).reshape(n, 3)
    lib().orc_element_mass(ct, float(density), _d(X), int(lumped), _d(acc), _d(Me), _d(fi))
    return Me.reshape(3 * n, 3 * n).T.copy(), fi


def calc_stress(ct, matkind, params, conn, X, u_col, stress_type=STRESS_2PK, strain_type=STRAIN_GL,
                mue_per_ele=None, tech=TECH_NONE, hist=None):
    n, g = num_nodes(ct), num_gp(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    u_col = np.ascontiguousarray(u_col, dtype=np.float64)
    if mue_per_ele is not None:
        mue_per_ele = np.ascontiguousarray(mue_per_ele, dtype=np.float64)
    st, en = np.zeros((len(conn), g, 6)), np.zeros((len(conn), g, 6))
    lib().orc_set_output_history(_d(hist))
    lib().orc_calc_stress(ct, tech, matkind, _d(params), _d(mue_per_ele), len(conn), _i(conn), _d(X),
                          _d(u_col), stress_type, strain_type, _d(st), _d(en))
    lib().orc_set_output_history(None)
    return st, en


def calc_energy(ct, matkind, params, conn, X, u_col, ele_is_row=None, mue_per_ele=None, tech=TECH_NONE, hist=None):
    n = num_nodes(ct)
    params = np.ascontiguousarray(params, dtype=np.float64)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    u_col = np.ascontiguousarray(u_col, dtype=np.float64)
    if mue_per_ele is not None:
        mue_per_ele = np.ascontiguousarray(mue_per_ele, dtype=np.float64)
    flag = None
    if ele_is_row is not None:
        ele_is_row = np.ascontiguousarray(ele_is_row, dtype=np.uint8)
        flag = ele_is_row.ctypes.data_as(C.POINTER(C.c_uint8))
    lib().orc_set_output_history(_d(hist))
    e = lib().orc_calc_energy(ct, tech, matkind, _d(params), _d(mue_per_ele), len(conn), _i(conn),
                              _d(X), _d(u_col), flag)
    lib().orc_set_output_history(None)
    return e


def assemble_mass(ct, density, conn, X, node_row_lid, node_col_lid, n_row_nodes, rowptr, colind,
                  lumped=False, acc_col=None):
    """Owner-computes assembly of the element mass matrices into the stiffness graph; returns
    (mass_values, f_inertia or None)."""
    n = num_nodes(ct)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    node_row_lid = np.ascontiguousarray(node_row_lid, dtype=np.int32)
    node_col_lid = np.ascontiguousarray(node_col_lid, dtype=np.int32)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    colind = np.ascontiguousarray(colind, dtype=np.int32)
    vals = np.zeros(rowptr[-1])
    fi = np.zeros(3 * n_row_nodes) if acc_col is not None else None
    acc_col = None if acc_col is None else np.ascontiguousarray(acc_col, dtype=np.float64)
    lib().orc_assemble_mass(ct, float(density), int(lumped), len(conn), _i(conn), _d(X), _d(acc_col),
                            _i(node_row_lid), _i(node_col_lid), n_row_nodes, _l(rowptr), _i(colind),
                            _d(vals), _d(fi))
    return vals, fi


def mulf_new_history(ct, n_ele, fbar=False):
    """Fresh MulfHistoryData of every Gauss point: (n_ele, ngp, MULF_HIST); fbar (MulfFBarFormulation): one more per
    element, the global history at the centroid."""
    npt = num_gp(ct) + (1 if fbar else 0)
    h = np.zeros((n_ele, npt, MULF_HIST))
    lib().orc_mulf_init_history(n_ele * npt, _d(h))
    return h


def update_prestress(ct, conn, X, u_col, hist):
    """struct_update_prestress: in place on hist (hist with ngp + 1 points per element: MulfFBarFormulation)."""
    if hist.shape[1] == num_gp(ct) + 1:
        n = num_nodes(ct)
        conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
        lib().orc_update_prestress_fbar(ct, len(conn), _i(conn), _d(np.ascontiguousarray(X, dtype=np.float64)),
                                        _d(np.ascontiguousarray(u_col, dtype=np.float64)), _d(hist))
        return
    n = num_nodes(ct)
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, n)
    X = np.ascontiguousarray(X, dtype=np.float64)
    u_col = np.ascontiguousarray(u_col, dtype=np.float64)
    assert hist.dtype == np.float64 and hist.flags["C_CONTIGUOUS"]
    lib().orc_update_prestress(ct, len(conn), _i(conn), _d(X), _d(u_col), _d(hist))


def eas_new_history(tech, n_ele):
    """Fresh EASHistoryData of every element: zeros (n_ele, orc_eas_hist_size)."""
    return np.zeros((n_ele, lib().orc_eas_hist_size(tech)))


def eas_recover(tech, conn, du_col, hist, step_length=1.0):
    """struct_calc_recover (default step): in place on hist."""
    conn = np.ascontiguousarray(conn, dtype=np.int32).reshape(-1, 8)
    du_col = np.ascontiguousarray(du_col, dtype=np.float64)
    assert hist.dtype == np.float64 and hist.flags["C_CONTIGUOUS"]
    lib().orc_eas_recover(tech, len(conn), _i(conn), _d(du_col), float(step_length), _d(hist))

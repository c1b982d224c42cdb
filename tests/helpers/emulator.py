"""ctypes front end of tests/cpp/emulate_hex8.cpp: the PRODUCT's hex8 device functions stepped on the
CPU (test infrastructure; see the header of that file)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CPP = os.path.join(os.path.dirname(_HERE), "cpp")
_libs = {}


def lib(cell="hex8"):
    cell = {"hex8fbar": "hex8", "hex8mulf": "hex8", "hex8linear": "hex8"}.get(cell, cell)
    if cell not in _libs:
        so = os.path.join(_CPP, f"libemu_{cell}.so")
        srcs = [os.path.join(_CPP, f"emulate_{cell}.cpp")] + [
            os.path.join(_CPP, "..", "..", "4c_b200", "csrc", f)
            for f in ("hex8_kernel.cuh", "hex8_fbar_kernel.cuh", "tet10_kernel.cuh", "hex27_kernel.cuh", "generic_kernel.cuh", "solid_math.cuh",
                      "summands.cuh", "history_materials.cuh", "post_kernels.cuh", "neumann_kernels.cuh", "shape_tables.hpp")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so, srcs[0]])
        _libs[cell] = C.CDLL(so)
    return _libs[cell]


def scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid):
    """numpy version of k_node_rows + k_build_scatter_map (host check of the device set-up)."""
    conn = np.asarray(conn)
    ncn = len(node_row_lid)
    rowstart = np.full(ncn, -1, np.int64)
    rowlen = np.zeros(ncn, np.int32)
    frow = np.full(ncn, -1, np.int32)
    own = node_row_lid >= 0
    r0 = 3 * node_row_lid[own].astype(np.int64)
    rowstart[own] = rowptr[r0]
    rowlen[own] = (rowptr[r0 + 1] - rowptr[r0]).astype(np.int32)
    frow[own] = r0
    ne, nn = conn.shape
    pos = np.full((ne, nn, nn), -1, np.int32)
    for e in range(ne):
        for i in range(nn):
            ni = conn[e, i]
            if rowstart[ni] < 0:
                continue
            idx = colind[rowstart[ni]:rowstart[ni] + rowlen[ni]]
            for j in range(nn):
                col = 3 * node_col_lid[conn[e, j]]
                p = np.searchsorted(idx, col)
                assert idx[p] == col and idx[p + 2] == col + 2
                pos[e, i, j] = p
    return pos, rowstart, rowlen, frow


def evaluate(matkind, params, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes,
             stiff=True, per_ele=None, cell="hex8", gp_hist=None):
    conn = np.ascontiguousarray(conn, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    u_col = np.ascontiguousarray(u_col, np.float64)
    params = np.ascontiguousarray(params, np.float64)
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    vals = np.zeros(rowptr[-1])
    f = np.zeros(3 * n_row_nodes)
    dJ, dF, nb = C.c_double(0), C.c_double(0), C.c_int64(0)
    vp = C.c_void_p
    p = lambda a: None if a is None else a.ctypes.data_as(vp)
    if per_ele is not None:
        per_ele = np.ascontiguousarray(per_ele, np.float64)
    rc = getattr(lib(cell), f"emu_{cell}")(C.c_int(matkind), p(params), p(per_ele), C.c_int(int(stiff)),
                        C.c_int64(conn.shape[0]), p(conn), p(X), p(u_col), p(pos), p(rowstart),
                        p(rowlen), p(frow), C.c_int64(len(rowstart)), p(vals), p(f), C.byref(dJ), C.byref(dF), C.byref(nb),
                        *([p(np.ascontiguousarray(gp_hist, np.float64))] if cell == "hex8mulf" else []))
    assert rc == 0
    return vals, f, dict(min_detJ=dJ.value, min_detF=dF.value, n_bad=nb.value)


_vp = C.c_void_p


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def material(matkind, params, Cv):
    """Product material at one point: C (Voigt) -> S[6], Cm[6,6], psi."""
    params = np.ascontiguousarray(params, np.float64)
    Cv = np.ascontiguousarray(Cv, np.float64)
    S, Cm, psi = np.zeros(6), np.zeros(36), C.c_double(0)
    rc = lib("post").emu_material(C.c_int(matkind), _p(params), _p(Cv), _p(S), _p(Cm), C.byref(psi))
    assert rc == 0, rc
    return S, Cm.reshape(6, 6), psi.value


def mulf_history(celltype, conn, X, u_col=None, hist=None, fbar=False):
    """hist is None: fresh history (n_ele, ngp, 18); else struct_update_prestress in place for state u_col.
    fbar: MULF + F-bar, ngp + 1 histories per element (the centroid's last)."""
    conn = np.ascontiguousarray(conn, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    ngp = {0: 8, 1: 27, 2: 4, 3: 1, 4: 27, 5: 6, 6: 8, 7: 18}[celltype]
    update = hist is not None
    if hist is None:
        hist = np.zeros((len(conn), ngp + (1 if fbar else 0), 18))
    u = None if u_col is None else np.ascontiguousarray(u_col, np.float64)
    assert hist.flags["C_CONTIGUOUS"]
    rc = lib("post").emu_mulf_history(C.c_int(celltype), C.c_int(int(update)), C.c_int64(len(conn)), _p(conn), _p(X),
                                      _p(u), _p(hist), C.c_int(int(fbar)))
    assert rc == 0
    return hist


def post(celltype, matkind, params, conn, X, u_col, stress_type=1, strain_type=1, ele_is_row=None,
         per_ele=None, fbar=False, want_energy=True, gp_hist=None):
    conn = np.ascontiguousarray(conn, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    u_col = np.ascontiguousarray(u_col, np.float64)
    params = np.ascontiguousarray(params, np.float64)
    ngp = {0: 8, 1: 27, 2: 4}[celltype]
    st = np.zeros((len(conn), ngp, 6))
    en = np.zeros((len(conn), ngp, 6))
    energy = C.c_double(0)
    if ele_is_row is not None:
        ele_is_row = np.ascontiguousarray(ele_is_row, np.uint8)
    if per_ele is not None:
        per_ele = np.ascontiguousarray(per_ele, np.float64)
    # fbar: False / True, or 2 = KINEM linear, 3 = TECH shell_ans (hex8)
    rc = lib("post").emu_post(C.c_int(celltype), C.c_int(int(fbar)), C.c_int(matkind), _p(params), _p(per_ele),
                              C.c_int64(len(conn)), _p(conn), _p(X), _p(u_col), _p(ele_is_row),
                              C.c_int(stress_type), C.c_int(strain_type), _p(st), _p(en),
                              C.byref(energy) if want_energy else None,
                              None if gp_hist is None else _p(np.ascontiguousarray(gp_hist, np.float64)))
    assert rc == 0, rc
    return st, en, energy.value


def mass(celltype, density, conn, X, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, lumped=False,
         acc_col=None):
    conn = np.ascontiguousarray(conn, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    vals = np.zeros(rowptr[-1])
    fi = np.zeros(3 * n_row_nodes)
    if acc_col is not None:
        acc_col = np.ascontiguousarray(acc_col, np.float64)
    rc = lib("post").emu_mass(C.c_int(celltype), C.c_double(density), C.c_int(int(lumped)), C.c_int64(len(conn)),
                              _p(conn), _p(X), _p(acc_col), _p(pos), _p(rowstart), _p(rowlen), _p(frow),
                              C.c_int64(len(rowstart)), _p(vals), _p(fi))
    assert rc == 0, rc
    return vals, fi


def _node_rows(rowptr, node_row_lid, node_col_lid):
    """numpy version of k_node_rows: per column node (rowstart, rowlen, frow, colbase); lids per NODE."""
    node_row_lid = np.asarray(node_row_lid)
    ncn = len(node_row_lid)
    rowstart = np.full(ncn, -1, np.int64)
    rowlen = np.zeros(ncn, np.int32)
    frow = np.full(ncn, -1, np.int32)
    own = node_row_lid >= 0
    r0 = 3 * node_row_lid[own].astype(np.int64)
    rowptr = np.asarray(rowptr, np.int64)
    rowstart[own] = rowptr[r0]
    rowlen[own] = (rowptr[r0 + 1] - rowptr[r0]).astype(np.int32)
    frow[own] = r0
    colbase = np.ascontiguousarray(3 * np.asarray(node_col_lid), np.int32)
    return rowstart, rowlen, frow, colbase


def neumann_surface(celltype, kind, faces, X, u_col, onoff, val, rowptr, colind, node_row_lid, node_col_lid, n_rows,
                    gp_fac=None, loadlin=False, want_gp=False):
    """Product surface-load functions on the CPU: (fext, K values or None, gp positions or None, flag)."""
    faces = np.ascontiguousarray(faces, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    rowstart, rowlen, frow, colbase = _node_rows(rowptr, node_row_lid, node_col_lid)
    colind = np.ascontiguousarray(colind, np.int32)
    fext = np.zeros(n_rows)
    K = np.zeros(int(rowptr[-1])) if loadlin else None
    ngp = {0: 4, 1: 9, 2: 6}[celltype]
    gp = np.zeros((len(faces), ngp, 3)) if want_gp else None
    if u_col is not None:
        u_col = np.ascontiguousarray(u_col, np.float64)
    if gp_fac is not None:
        gp_fac = np.ascontiguousarray(gp_fac, np.float64)
    flag = lib("neumann").emu_neumann_surface(
        C.c_int(celltype), C.c_int(kind), C.c_int64(len(faces)), _p(faces), _p(X), _p(u_col),
        _p(np.ascontiguousarray(onoff, np.int32)), _p(np.ascontiguousarray(val, np.float64)), _p(gp_fac),
        C.c_int64(len(rowstart)), _p(rowstart), _p(rowlen), _p(frow), _p(colbase), _p(colind), _p(fext), _p(K), _p(gp))
    return fext, K, gp, flag


def neumann_volume(celltype, conn, X, onoff, val, rowptr, node_row_lid, node_col_lid, n_rows, ele_slot=None,
                   gp_fac=None):
    conn = np.ascontiguousarray(conn, np.int32)
    X = np.ascontiguousarray(X, np.float64)
    rowstart, rowlen, frow, _ = _node_rows(rowptr, node_row_lid, node_col_lid)
    fext = np.zeros(n_rows)
    if ele_slot is not None:
        ele_slot = np.ascontiguousarray(ele_slot, np.int32)
    if gp_fac is not None:
        gp_fac = np.ascontiguousarray(gp_fac, np.float64)
    rc = lib("neumann").emu_neumann_volume(
        C.c_int(celltype), C.c_int64(len(conn)), _p(conn), _p(X), _p(ele_slot),
        _p(np.ascontiguousarray(onoff, np.int32)), _p(np.ascontiguousarray(val, np.float64)), _p(gp_fac),
        C.c_int64(len(rowstart)), _p(rowstart), _p(rowlen), _p(frow), _p(fext))
    assert rc == 0
    return fext


def drucker_prager(in10, Cv, hist=None):
    """Product MAT_Struct_DruckerPrager at one point: (S, Cm) or None if the device rejects the parameters."""
    in10 = np.ascontiguousarray(in10, np.float64)
    Cv = np.ascontiguousarray(Cv, np.float64)
    S, Cm = np.zeros(6), np.zeros(36)
    rc = lib("post").emu_drucker_prager(_p(in10), _p(hist), _p(Cv), _p(S), _p(Cm))
    if rc == -3:
        return None
    assert rc in (0, 1)   # 1: the full (non-symmetric) tangent of the general kernel, stored [I + 6 J]
    return S, (Cm.reshape(6, 6).T.copy() if rc == 1 else Cm.reshape(6, 6))


def hex8_drucker_prager(in10, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, mat_hist,
                        stiff=True):
    """Emulated KINEM linear hex8 kernel with the history material; mat_hist (n_ele, 8, 14) updated in place."""
    conn = np.ascontiguousarray(conn, np.int32)
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    vals, f = np.zeros(rowptr[-1]), np.zeros(3 * n_row_nodes)
    rc = lib("hex8").emu_hex8_drucker_prager(
        _p(np.ascontiguousarray(in10, np.float64)), C.c_int(int(stiff)), C.c_int64(len(conn)), _p(conn),
        _p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(u_col, np.float64)), _p(pos), _p(rowstart),
        _p(rowlen), _p(frow), C.c_int64(len(rowstart)), _p(vals), _p(f), _p(mat_hist))
    assert rc == 0, rc
    return vals, f


def visco_neohooke(in6, dt, Cv, hist=None):
    """Product MAT_VISCONEOHOOKE at one point: (S, Cm); hist (24,) updated in place or None."""
    in6 = np.ascontiguousarray(in6, np.float64)
    Cv = np.ascontiguousarray(Cv, np.float64)
    S, Cm = np.zeros(6), np.zeros(36)
    L = lib("post")
    L.emu_visco_neohooke.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.emu_visco_neohooke(_p(in6), float(dt), _p(hist), _p(Cv), _p(S), _p(Cm))
    if rc == -3:
        return None
    assert rc == 0
    return S, Cm.reshape(6, 6)


def plastic_nlnlog(in10, dt, Cv, hist=None):
    """Product MAT_Struct_PlasticNlnLogNeoHooke at one point: (S, Cm); hist (14,) updated in place or None."""
    in10 = np.ascontiguousarray(in10, np.float64)
    Cv = np.ascontiguousarray(Cv, np.float64)
    S, Cm = np.zeros(6), np.zeros(36)
    L = lib("post")
    L.emu_plastic_nlnlog.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.emu_plastic_nlnlog(_p(in10), float(dt), _p(hist), _p(Cv), _p(S), _p(Cm))
    if rc == -3:
        return None
    assert rc == 0
    return S, Cm.reshape(6, 6)


def hex8_visco(in6, dt, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, mat_hist, stiff=True):
    """Emulated nonlinear hex8 kernel with MAT_VISCONEOHOOKE; mat_hist (n_ele, 8, 24) updated in place."""
    conn = np.ascontiguousarray(conn, np.int32)
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    vals, f = np.zeros(rowptr[-1]), np.zeros(3 * n_row_nodes)
    L = lib("hex8")
    L.emu_hex8_visco.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int64] + [C.c_void_p] * 3
    rc = L.emu_hex8_visco(_p(np.ascontiguousarray(in6, np.float64)), float(dt), int(stiff), len(conn), _p(conn),
                          _p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(u_col, np.float64)), _p(pos),
                          _p(rowstart), _p(rowlen), _p(frow), len(rowstart), _p(vals), _p(f), _p(mat_hist))
    assert rc == 0, rc
    return vals, f


def generic_drucker_prager(celltype, in10, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, mat_hist,
                           stiff=True):
    """Emulated general-tangent kernel (generic_kernel.cuh), KINEM linear, MAT_Struct_DruckerPrager with a tangent
    without major symmetry; mat_hist (n_ele, ngp, 14) updated in place."""
    conn = np.ascontiguousarray(conn, np.int32)
    nn = conn.shape[1]
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    pos = np.ascontiguousarray(pos.reshape(len(conn), -1))
    vals, f = np.zeros(rowptr[-1]), np.zeros(3 * n_row_nodes)
    L = lib("post")
    L.emu_generic_drucker_prager.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 4 + [C.c_int] + \
        [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 3
    rc = L.emu_generic_drucker_prager(int(celltype), _p(np.ascontiguousarray(in10, np.float64)), int(stiff), len(conn), _p(conn),
                                      _p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(u_col, np.float64)),
                                      _p(pos), pos.shape[1], _p(rowstart), _p(rowlen), _p(frow), len(rowstart), _p(vals),
                                      _p(f), _p(mat_hist))
    assert rc == 0, rc
    return vals, f


def generic(celltype, matkind, params, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, linear=False,
            stiff=True, gp_hist=None, fbar=False, ans=False, eas_hist=None):
    """Emulated general kernel (generic_kernel.cuh) for the hyperelastic kinds 0..2 and any cell type of
    shape_tables.hpp (tet4 / hex20 blocks run through it on the device)."""
    conn = np.ascontiguousarray(conn, np.int32)
    pos, rowstart, rowlen, frow = scatter_map(conn, rowptr, colind, node_row_lid, node_col_lid)
    pos = np.ascontiguousarray(pos.reshape(len(conn), -1))
    vals, f = np.zeros(rowptr[-1]), np.zeros(3 * n_row_nodes)
    L = lib("post")
    L.emu_generic.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 4 + [C.c_int] + \
        [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p]
    if eas_hist is not None:  # TECH shell_eas_ans: (n_ele, 238) doubles, updated in place
        assert eas_hist.dtype == np.float64 and eas_hist.flags["C_CONTIGUOUS"]
    rc = L.emu_generic(int(celltype), int(matkind), _p(np.ascontiguousarray(params, np.float64)), int(linear), int(stiff),
                       len(conn), _p(conn), _p(np.ascontiguousarray(X, np.float64)),
                       _p(np.ascontiguousarray(u_col, np.float64)), _p(pos), pos.shape[1], _p(rowstart), _p(rowlen),
                       _p(frow), len(rowstart), _p(vals), _p(f), _p(gp_hist), int(fbar), int(ans), _p(eas_hist))
    assert rc == 0, rc
    return vals, f


def hex27_padded(celltype, matkind, params, conn, X, u_col, rowptr, colind, node_row_lid, node_col_lid, n_row_nodes, stiff=True):
    """hex20 (celltype 4) / hex18 (7) block through the emulated hex27 kernel in its PAD instantiation: connectivity padded to 27
    nodes with the element's node 0 (k_pad_conn), 27 x 27 scatter map with -1 for padding rows / columns
    (k_build_scatter_map(nn_real)), padded shape tables from the product's pad_tables_to_hex27."""
    conn = np.ascontiguousarray(conn, np.int32)
    n_ele, nn = conn.shape
    conn27 = np.concatenate([conn, np.repeat(conn[:, :1], 27 - nn, axis=1)], axis=1).astype(np.int32)
    pos, rowstart, rowlen, frow = scatter_map(conn27, rowptr, colind, node_row_lid, node_col_lid)
    pos = np.ascontiguousarray(pos.reshape(n_ele, 27, 27))
    pos[:, nn:, :] = -1
    pos[:, :, nn:] = -1
    pos = np.ascontiguousarray(pos.reshape(n_ele, 729))
    conn27 = np.ascontiguousarray(conn27)
    vals, f = np.zeros(rowptr[-1]), np.zeros(3 * n_row_nodes)
    dJ, dF = C.c_double(0), C.c_double(0)
    L = lib("hex27")
    L.emu_hex27_padded.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int64] + [C.c_void_p] * 4
    rc = L.emu_hex27_padded(int(celltype), int(matkind), _p(np.ascontiguousarray(params, np.float64)), int(stiff), n_ele, _p(conn27),
                            _p(np.ascontiguousarray(X, np.float64)), _p(np.ascontiguousarray(u_col, np.float64)), _p(pos),
                            _p(rowstart), _p(rowlen), _p(frow), len(rowstart), _p(vals), _p(f), C.byref(dJ), C.byref(dF))
    assert rc == 0, rc
    return vals, f, dict(min_detJ=dJ.value, min_detF=dF.value)

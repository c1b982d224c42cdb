"""ctypes binding of lib4c_b200_solid.so (include/b200_solid.h).

The hot path lives entirely in the CUDA library; this module only marshals pointers.  Loading
fails loudly when the library has not been built -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200S_LIB") or os.path.join(_HERE, "lib4c_b200_solid.so")  # B200S_LIB: A/B builds (tools/)

OK = 0
ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NEG_JACOBIAN, ERR_NONFINITE, ERR_COMM = -1, -2, -3, -4, -5, -6
HEX8, HEX27, TET10, TET4, HEX20, WEDGE6, PYRAMID5, HEX18 = 0, 1, 2, 3, 4, 5, 6, 7
MAT_COUP_NEOHOOKE, MAT_ISO_VOL, MAT_SVK, MAT_ELASTHYPER, MAT_DRUCKER_PRAGER, MAT_VISCONEOHOOKE = 0, 1, 2, 3, 4, 5
MAT_PLASTIC_NLNLOG = 6  # MAT_Struct_PlasticNlnLogNeoHooke
ACTION_INTERNALFORCE, ACTION_NLNSTIFF, ACTION_NLNSTIFFMASS, ACTION_NLNSTIFFLMASS, ACTION_INTERNALINERTIAFORCE = 0, 1, 2, 3, 4
STRESS_NONE, STRESS_2PK, STRESS_CAUCHY = 0, 1, 2
STRAIN_NONE, STRAIN_GL, STRAIN_EA = 0, 1, 2
MATRIX_STIFFNESS, MATRIX_MASS = 0, 1
NEUMANN_LIVE, NEUMANN_PSEUDO_ORTHOPRESSURE, NEUMANN_ORTHOPRESSURE = 0, 1, 2
NEUMANN_LOADLIN, NEUMANN_MULF_ACTIVE = 1, 2
# b200s_summand ids (include/b200_solid.h), by name
SUMMANDS = ["COUP_NEOHOOKE", "COUP_MOONEYRIVLIN", "COUP_LOGNEOHOOKE", "COUP_LOGMIXNEOHOOKE", "COUP_SVK",
            "COUP_SIMOPISTER", "COUP_BLATZKO", "COUP_EXPPOL", "COUP_1POW", "COUP_2POW", "COUP_3POW",
            "COUP_13APOW", "ISO_NEOHOOKE", "ISO_MOONEYRIVLIN", "ISO_YEOH", "ISO_1POW", "ISO_2POW",
            "ISO_EXPOPOW", "VOL_SUSSMANBATHE", "VOL_PENALTY", "VOL_OGDEN", "VOL_POW", "COUP_ANISOEXPO",
            "ISO_OGDEN", "COUP_VARGA", "ISO_VARGA"]
SUM = {name: i for i, name in enumerate(SUMMANDS)}
KINEM_NONLINEAR_TOTLAG, KINEM_LINEAR = 0, 1
TECH_NONE, TECH_FBAR, TECH_EAS_MILD, TECH_EAS_FULL, TECH_SHELL_ANS, TECH_SHELL_EAS, TECH_SHELL_EAS_ANS = 0, 1, 2, 3, 4, 5, 6
PRESTRESS_NONE, PRESTRESS_MULF = 0, 1

# every symbol include/b200_solid.h declares (checked by the CPU test-suite)
SYMBOLS = [
    "b200s_create", "b200s_destroy", "b200s_last_error", "b200s_set_option", "b200s_set_mesh",
    "b200s_set_element_properties", "b200s_set_material", "b200s_set_graph", "b200s_set_halo", "b200s_nccl_unique_id",
    "b200s_set_row_export", "b200s_add_exported",
    "b200s_comm_init_nccl", "b200s_evaluate", "b200s_evaluate_async", "b200s_synchronize",
    "b200s_device_values", "b200s_device_f", "b200s_device_u_row", "b200s_device_u_col",
    "b200s_stream", "b200s_nnz", "b200s_n_rows", "b200s_get_timers", "b200s_launch_count",
    "b200s_kernel_time",
    "b200s_build_graph", "b200s_measure_fp64_peak", "b200s_microbench", "b200s_host_alloc", "b200s_host_free",
    "b200s_version",
    "b200s_set_material_elasthyper", "b200s_set_density", "b200s_set_row_elements", "b200s_set_state",
    "b200s_get_mass", "b200s_device_mass_values", "b200s_device_f_inertia", "b200s_calc_stress",
    "b200s_calc_energy", "b200s_apply_dirichlet", "b200s_apply_dirichlet_to_system", "b200s_multiply",
    "b200s_last_multiply_ms", "b200s_solve_cg", "b200s_get_values",
    "b200s_update_prestress", "b200s_gp_history_size", "b200s_get_gp_history", "b200s_set_gp_history",
    "b200s_recover", "b200s_element_history_size", "b200s_get_element_history", "b200s_set_element_history",
    "b200s_set_element_fibers", "b200s_recover_linesearch", "b200s_update_step", "b200s_reset_step",
    "b200s_add_surface_neumann", "b200s_add_volume_neumann", "b200s_clear_neumann", "b200s_neumann_gp_count",
    "b200s_neumann_gp_coords", "b200s_set_neumann_factors", "b200s_evaluate_neumann", "b200s_device_fext",
    "b200s_material_history_size", "b200s_get_material_history", "b200s_set_material_history",
    "b200s_get_value_rows", "b200s_get_f_rows", "b200s_eas_old_step_length", "b200s_set_eas_old_step_length",
    "b200s_attach_block", "b200s_host_alloc_near", "b200s_last_evaluate_info", "b200s_set_time_step_size",
    "b200s_last_d2h_value_bytes", "b200s_symmetric_transfer_status", "b200s_debug_element_order",
]


class ErrorInfo(C.Structure):
    _fields_ = [("code", C.c_int32), ("n_neg_jacobian", C.c_int32), ("first_bad_element", C.c_int64),
                ("min_detJ", C.c_double), ("min_detF", C.c_double), ("n_nonfinite", C.c_int64)]


class B200SolidError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200s error {code}: {msg}")
        self.code = code


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a). The device path has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.b200s_create.argtypes = [i32, i32, i32, C.POINTER(vp)]
    L.b200s_destroy.argtypes = [vp]
    L.b200s_destroy.restype = None
    L.b200s_last_error.argtypes = [vp]
    L.b200s_last_error.restype = C.c_char_p
    L.b200s_set_option.argtypes = [vp, C.c_char_p, dbl]
    L.b200s_set_mesh.argtypes = [vp, i32, i64, vp, i64, vp]
    L.b200s_set_element_properties.argtypes = [vp, i32, i32, i32]
    L.b200s_set_material.argtypes = [vp, i32, vp, i32, vp]
    L.b200s_set_graph.argtypes = [vp, i64, vp, vp, vp, vp]
    L.b200s_set_halo.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.b200s_nccl_unique_id.argtypes = [vp, vp]
    L.b200s_comm_init_nccl.argtypes = [vp, vp]
    L.b200s_evaluate.argtypes = [vp, i32, vp, vp, vp, C.POINTER(ErrorInfo)]
    L.b200s_evaluate_async.argtypes = [vp, i32]
    L.b200s_synchronize.argtypes = [vp, C.POINTER(ErrorInfo)]
    for fn in ("b200s_device_values", "b200s_device_f", "b200s_device_u_row", "b200s_device_u_col",
               "b200s_stream"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = vp
    L.b200s_nnz.argtypes = [vp]
    L.b200s_nnz.restype = i64
    L.b200s_n_rows.argtypes = [vp]
    L.b200s_n_rows.restype = i64
    L.b200s_get_timers.argtypes = [vp, vp]
    L.b200s_launch_count.argtypes = [vp]
    L.b200s_launch_count.restype = i64
    L.b200s_kernel_time.argtypes = [vp, C.POINTER(dbl), C.POINTER(i64), i32]
    L.b200s_build_graph.argtypes = [vp, i64, vp, vp, vp, vp, C.POINTER(i64)]
    L.b200s_measure_fp64_peak.argtypes = [i32, C.POINTER(dbl), C.POINTER(dbl)]
    L.b200s_microbench.argtypes = [i32, i32, i64, C.POINTER(dbl)]
    L.b200s_host_alloc.argtypes = [i64]
    L.b200s_host_alloc.restype = vp
    L.b200s_host_free.argtypes = [vp]
    L.b200s_host_free.restype = None
    L.b200s_version.restype = C.c_char_p
    L.b200s_set_material_elasthyper.argtypes = [vp, i32, vp, vp, vp]
    L.b200s_set_density.argtypes = [vp, dbl]
    L.b200s_set_element_fibers.argtypes = [vp, vp]
    L.b200s_set_row_elements.argtypes = [vp, vp]
    L.b200s_set_state.argtypes = [vp, C.c_char_p, vp]
    L.b200s_get_mass.argtypes = [vp, vp, vp]
    for fn in ("b200s_device_mass_values", "b200s_device_f_inertia"):
        getattr(L, fn).argtypes = [vp]
        getattr(L, fn).restype = vp
    L.b200s_calc_stress.argtypes = [vp, i32, i32, vp, vp]
    L.b200s_get_values.argtypes = [vp, i32, vp]
    L.b200s_update_prestress.argtypes = [vp]
    L.b200s_recover.argtypes = [vp, vp, dbl]
    L.b200s_recover_linesearch.argtypes = [vp, dbl]
    L.b200s_update_step.argtypes = [vp]
    L.b200s_reset_step.argtypes = [vp]
    L.b200s_element_history_size.argtypes = [vp]
    L.b200s_element_history_size.restype = i64
    L.b200s_get_element_history.argtypes = [vp, vp]
    L.b200s_set_element_history.argtypes = [vp, vp]
    L.b200s_gp_history_size.argtypes = [vp]
    L.b200s_gp_history_size.restype = i64
    L.b200s_get_gp_history.argtypes = [vp, vp]
    L.b200s_set_gp_history.argtypes = [vp, vp]
    L.b200s_calc_energy.argtypes = [vp, C.POINTER(dbl)]
    L.b200s_apply_dirichlet.argtypes = [vp, i32, vp, i32]
    L.b200s_apply_dirichlet_to_system.argtypes = [vp, vp, vp, vp, vp]
    L.b200s_multiply.argtypes = [vp, i32, vp, vp]
    L.b200s_last_multiply_ms.argtypes = [vp]
    L.b200s_last_multiply_ms.restype = dbl
    L.b200s_solve_cg.argtypes = [vp, vp, vp, dbl, i32, C.POINTER(i32), C.POINTER(dbl)]
    L.b200s_add_surface_neumann.argtypes = [vp, i32, i64, vp, vp, vp]
    L.b200s_add_volume_neumann.argtypes = [vp, i64, vp, vp, vp]
    L.b200s_clear_neumann.argtypes = [vp]
    L.b200s_neumann_gp_count.argtypes = [vp, i32]
    L.b200s_neumann_gp_count.restype = i64
    L.b200s_neumann_gp_coords.argtypes = [vp, i32, i32, vp]
    L.b200s_set_neumann_factors.argtypes = [vp, i32, vp, vp]
    L.b200s_evaluate_neumann.argtypes = [vp, vp, i32, vp]
    L.b200s_device_fext.argtypes = [vp]
    L.b200s_device_fext.restype = vp
    L.b200s_material_history_size.argtypes = [vp]
    L.b200s_material_history_size.restype = i64
    L.b200s_get_material_history.argtypes = [vp, vp]
    L.b200s_set_material_history.argtypes = [vp, vp]
    L.b200s_get_value_rows.argtypes = [vp, i32, i64, i64, vp, vp, vp]
    L.b200s_get_f_rows.argtypes = [vp, i64, i64, vp]
    L.b200s_eas_old_step_length.argtypes = [vp]
    L.b200s_eas_old_step_length.restype = dbl
    L.b200s_set_eas_old_step_length.argtypes = [vp, dbl]
    L.b200s_attach_block.argtypes = [vp, vp]
    L.b200s_host_alloc_near.argtypes = [i32, i64, C.POINTER(i32)]
    L.b200s_host_alloc_near.restype = vp
    L.b200s_last_evaluate_info.argtypes = [vp, vp]
    L.b200s_set_time_step_size.argtypes = [vp, dbl]
    L.b200s_debug_element_order.argtypes = [vp, vp]
    L.b200s_last_d2h_value_bytes.argtypes = [vp]
    L.b200s_last_d2h_value_bytes.restype = i64
    L.b200s_symmetric_transfer_status.argtypes = [vp]
    L.b200s_symmetric_transfer_status.restype = C.c_char_p
    _lib = L
    return L


def drucker_prager_params(YOUNG, NUE, ISOHARD, TOL, C, ETA, XI, ETABAR, MAXITER=50, TANG="consistent", **_):
    """Parameter block of MAT_DRUCKER_PRAGER from the keys of the MAT_Struct_DruckerPrager input line."""
    return np.array([YOUNG, NUE, ISOHARD, TOL, C, ETA, XI, ETABAR, MAXITER, 0.0 if TANG == "elastic" else 1.0], float)


def _ptr(a):
    """Raw address of a numpy array, torch tensor, int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor (host or device)
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """One rank's device context (b200s_ctx)."""

    def __init__(self, device=0, rank=0, nranks=1):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.b200s_create(device, rank, nranks, C.byref(h))
        if rc != OK:
            raise B200SolidError(rc, "b200s_create failed (no CUDA device?)")
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.b200s_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise B200SolidError(rc, self.L.b200s_last_error(self.h).decode())
        return rc

    def set_option(self, key, value):
        self._ck(self.L.b200s_set_option(self.h, key.encode(), float(value)))

    def set_mesh(self, celltype, X, conn):
        X = np.ascontiguousarray(X, np.float64)
        conn = np.ascontiguousarray(conn, np.int32)
        self.celltype, self.nn = celltype, conn.shape[1]
        self.n_col_nodes, self.n_col_eles = X.shape[0], conn.shape[0]
        self._ck(self.L.b200s_set_mesh(self.h, celltype, X.shape[0], _ptr(X), conn.shape[0], _ptr(conn)))

    def set_element_properties(self, kinem=KINEM_NONLINEAR_TOTLAG, tech=TECH_NONE,
                               prestress=PRESTRESS_NONE):
        self._ck(self.L.b200s_set_element_properties(self.h, kinem, tech, prestress))

    def set_material(self, kind, params, per_ele_param=None):
        params = np.ascontiguousarray(params, np.float64)
        if per_ele_param is not None:
            per_ele_param = np.ascontiguousarray(per_ele_param, np.float64)
        self._ck(self.L.b200s_set_material(self.h, kind, _ptr(params), len(params), _ptr(per_ele_param)))

    def set_material_elasthyper(self, summands):
        """summands: list of (name or id, [input-file parameters]) or (name, params, mode)."""
        kinds = np.asarray([SUM[s[0]] if isinstance(s[0], str) else int(s[0]) for s in summands], np.int32)
        modes = np.asarray([s[2] if len(s) > 2 else 0 for s in summands], np.int32)
        par = np.zeros((len(summands), 4))
        for k, s in enumerate(summands):
            par[k, :len(s[1])] = s[1]
        self._ck(self.L.b200s_set_material_elasthyper(self.h, len(summands), _ptr(kinds), _ptr(modes), _ptr(par)))

    def set_element_fibers(self, fiber1):
        f = None if fiber1 is None else np.ascontiguousarray(fiber1, np.float64).reshape(self.n_col_eles, 3)
        self._ck(self.L.b200s_set_element_fibers(self.h, _ptr(f)))

    def set_time_step_size(self, dt):
        self._ck(self.L.b200s_set_time_step_size(self.h, float(dt)))

    def set_density(self, density):
        self._ck(self.L.b200s_set_density(self.h, float(density)))

    def set_row_elements(self, is_row):
        is_row = None if is_row is None else np.ascontiguousarray(is_row, np.uint8)
        self._ck(self.L.b200s_set_row_elements(self.h, _ptr(is_row)))

    def set_state(self, name, v_row):
        v_row = np.ascontiguousarray(v_row, np.float64)
        self._ck(self.L.b200s_set_state(self.h, name.encode(), _ptr(v_row)))

    def get_mass(self, want_matrix=True, want_inertia=True):
        M = np.empty(self.nnz) if want_matrix else None
        fi = np.empty(self.n_rows) if want_inertia else None
        self._ck(self.L.b200s_get_mass(self.h, _ptr(M), _ptr(fi)))
        return M, fi

    def update_prestress(self):
        self._ck(self.L.b200s_update_prestress(self.h))

    def recover(self, du_row, step_length=1.0):
        du = np.ascontiguousarray(du_row, np.float64)
        self._ck(self.L.b200s_recover(self.h, _ptr(du), float(step_length)))

    def recover_linesearch(self, new_step_length):
        self._ck(self.L.b200s_recover_linesearch(self.h, float(new_step_length)))

    def update_step(self):
        self._ck(self.L.b200s_update_step(self.h))

    def reset_step(self):
        self._ck(self.L.b200s_reset_step(self.h))

    def get_element_history(self):
        h = np.empty(self.L.b200s_element_history_size(self.h))
        self._ck(self.L.b200s_get_element_history(self.h, _ptr(h)))
        return h

    def set_element_history(self, hist):
        h = np.ascontiguousarray(hist, np.float64).ravel()
        assert h.size == self.L.b200s_element_history_size(self.h)
        self._ck(self.L.b200s_set_element_history(self.h, _ptr(h)))

    def get_material_history(self):
        """History material state, (n_col_eles, num_gp, 14) flattened: Mat::PlasticDruckerPrager::pack."""
        h = np.empty(self.L.b200s_material_history_size(self.h))
        self._ck(self.L.b200s_get_material_history(self.h, _ptr(h)))
        return h

    def set_material_history(self, hist):
        h = np.ascontiguousarray(hist, np.float64).ravel()
        assert h.size == self.L.b200s_material_history_size(self.h)
        self._ck(self.L.b200s_set_material_history(self.h, _ptr(h)))

    def get_gp_history(self):
        h = np.empty(self.L.b200s_gp_history_size(self.h))
        self._ck(self.L.b200s_get_gp_history(self.h, _ptr(h)))
        return h

    def set_gp_history(self, hist):
        h = np.ascontiguousarray(hist, np.float64).ravel()
        assert h.size == self.L.b200s_gp_history_size(self.h)
        self._ck(self.L.b200s_set_gp_history(self.h, _ptr(h)))

    def get_values(self, which=MATRIX_STIFFNESS):
        v = np.empty(self.nnz)
        self._ck(self.L.b200s_get_values(self.h, which, _ptr(v)))
        return v

    def get_value_rows(self, row0, row1, which=MATRIX_STIFFNESS, want_values=True):
        """(rowptr relative to the first entry, colind, values) of the local rows [row0,row1) of the resident CSR."""
        rp = np.zeros(row1 - row0 + 1, np.int64)
        self._ck(self.L.b200s_get_value_rows(self.h, which, row0, row1, None, None, _ptr(rp)))
        ci = np.zeros(int(rp[-1]), np.int32)
        v = np.zeros(int(rp[-1])) if want_values else None
        self._ck(self.L.b200s_get_value_rows(self.h, which, row0, row1, _ptr(v), _ptr(ci), None))
        return rp, ci, v

    def get_f_rows(self, row0, row1):
        f = np.zeros(row1 - row0)
        self._ck(self.L.b200s_get_f_rows(self.h, row0, row1, _ptr(f)))
        return f

    def attach_block(self, primary):
        """This context becomes a further element block of `primary`'s system matrix (b200s_attach_block)."""
        self._ck(self.L.b200s_attach_block(self.h, primary.h))
        self._primary = primary  # keep it alive

    def last_evaluate_info(self):
        o = np.zeros(5, np.int64)
        self._ck(self.L.b200s_last_evaluate_info(self.h, _ptr(o)))
        return dict(kernel_launches=int(o[0]), pipeline_chunks=int(o[1]), halo_overlapped=int(o[2]),
                    row_phases=int(o[3]), ghost_elements=int(o[4]))

    def element_order(self):
        o = np.zeros(self.n_col_eles, np.int32)
        self._ck(self.L.b200s_debug_element_order(self.h, _ptr(o)))
        return o

    def last_d2h_value_bytes(self):
        return self.L.b200s_last_d2h_value_bytes(self.h)

    def symmetric_transfer_status(self):
        return self.L.b200s_symmetric_transfer_status(self.h).decode()

    def eas_old_step_length(self):
        return self.L.b200s_eas_old_step_length(self.h)

    def set_eas_old_step_length(self, v):
        self._ck(self.L.b200s_set_eas_old_step_length(self.h, float(v)))

    def calc_stress(self, num_gp, stress_type=STRESS_2PK, strain_type=STRAIN_GL):
        st = np.empty((self.n_col_eles, num_gp, 6)) if stress_type != STRESS_NONE else None
        en = np.empty((self.n_col_eles, num_gp, 6)) if strain_type != STRAIN_NONE else None
        self._ck(self.L.b200s_calc_stress(self.h, stress_type, strain_type, _ptr(st), _ptr(en)))
        return st, en

    def calc_energy(self):
        e = C.c_double(0)
        self._ck(self.L.b200s_calc_energy(self.h, C.byref(e)))
        return e.value

    # ---- external loads (Discretization::evaluate_neumann) ----
    def add_surface_neumann(self, neumann_type, face_nodes, onoff, val):
        f = np.ascontiguousarray(face_nodes, np.int32)
        o = np.ascontiguousarray(onoff, np.int32)
        v = np.ascontiguousarray(val, np.float64)
        assert len(o) >= 3 and len(v) >= 3
        rc = self.L.b200s_add_surface_neumann(self.h, int(neumann_type), len(f), _ptr(f), _ptr(o), _ptr(v))
        if rc < 0:
            self._ck(rc)
        return rc

    def add_volume_neumann(self, onoff, val, ele_ids=None):
        o = np.ascontiguousarray(onoff, np.int32)
        v = np.ascontiguousarray(val, np.float64)
        e = None if ele_ids is None else np.ascontiguousarray(ele_ids, np.int32)
        rc = self.L.b200s_add_volume_neumann(self.h, 0 if e is None else len(e), _ptr(e), _ptr(o), _ptr(v))
        if rc < 0:
            self._ck(rc)
        return rc

    def clear_neumann(self):
        self._ck(self.L.b200s_clear_neumann(self.h))

    def neumann_gp_coords(self, cond, flags=0):
        n = self.L.b200s_neumann_gp_count(self.h, int(cond))
        if n < 0:
            raise B200SolidError(ERR_INVALID, f"no Neumann condition {cond}")
        out = np.zeros((n, 3))
        self._ck(self.L.b200s_neumann_gp_coords(self.h, int(cond), int(flags), _ptr(out)))
        return out

    def set_neumann_factors(self, cond, dof_fac=None, gp_fac=None):
        d = None if dof_fac is None else np.ascontiguousarray(dof_fac, np.float64)
        g = None if gp_fac is None else np.ascontiguousarray(gp_fac, np.float64)
        self._ck(self.L.b200s_set_neumann_factors(self.h, int(cond), _ptr(d), _ptr(g)))

    def evaluate_neumann(self, u_row=None, flags=0):
        u = None if u_row is None else np.ascontiguousarray(u_row, np.float64)
        f = np.zeros(self.n_rows)
        self._ck(self.L.b200s_evaluate_neumann(self.h, _ptr(u), int(flags), _ptr(f)))
        return f

    def apply_dirichlet(self, dbctoggle_row, diagonalblock=True, which=MATRIX_STIFFNESS):
        t = np.ascontiguousarray(dbctoggle_row, np.float64)
        self._ck(self.L.b200s_apply_dirichlet(self.h, which, _ptr(t), int(diagonalblock)))

    def apply_dirichlet_to_system(self, x_row, b_row, dbcval_row, dbctoggle_row):
        """In place on the numpy arrays x_row / b_row (either may be None)."""
        v = np.ascontiguousarray(dbcval_row, np.float64)
        t = np.ascontiguousarray(dbctoggle_row, np.float64)
        self._ck(self.L.b200s_apply_dirichlet_to_system(self.h, _ptr(x_row), _ptr(b_row), _ptr(v), _ptr(t)))

    def multiply(self, x_row, which=MATRIX_STIFFNESS):
        x = np.ascontiguousarray(x_row, np.float64)
        y = np.empty_like(x)
        self._ck(self.L.b200s_multiply(self.h, which, _ptr(x), _ptr(y)))
        return y

    def last_multiply_ms(self):
        return self.L.b200s_last_multiply_ms(self.h)

    def solve_cg(self, b_row, x0=None, rtol=1e-10, maxit=1000):
        b = np.ascontiguousarray(b_row, np.float64)
        x = np.zeros_like(b) if x0 is None else np.array(x0, np.float64)
        it, rel = C.c_int(0), C.c_double(0)
        self._ck(self.L.b200s_solve_cg(self.h, _ptr(b), _ptr(x), float(rtol), int(maxit), C.byref(it), C.byref(rel)))
        return x, it.value, rel.value

    def set_graph(self, rowptr, colind, dof_row_lid, dof_col_lid):
        rowptr = rowptr if not isinstance(rowptr, np.ndarray) else np.ascontiguousarray(rowptr, np.int64)
        colind = colind if not isinstance(colind, np.ndarray) else np.ascontiguousarray(colind, np.int32)
        dof_row_lid = np.ascontiguousarray(dof_row_lid, np.int32)
        dof_col_lid = np.ascontiguousarray(dof_col_lid, np.int32)
        n_rows = (rowptr.shape[0] if hasattr(rowptr, "shape") else len(rowptr)) - 1
        self._ck(self.L.b200s_set_graph(self.h, n_rows, _ptr(rowptr), _ptr(colind), _ptr(dof_row_lid),
                                        _ptr(dof_col_lid)))

    def build_graph(self, n_row_nodes, node_row_lid, node_col_lid, fetch=False):
        """Device-side graph construction + installation; returns nnz (and rowptr, colind if fetch)."""
        node_row_lid = np.ascontiguousarray(node_row_lid, np.int32)
        node_col_lid = np.ascontiguousarray(node_col_lid, np.int32)
        nnz = C.c_int64(0)
        self._ck(self.L.b200s_build_graph(self.h, n_row_nodes, _ptr(node_row_lid), _ptr(node_col_lid),
                                          None, None, C.byref(nnz)))
        if not fetch:
            return nnz.value
        rowptr = np.zeros(3 * n_row_nodes + 1, np.int64)
        colind = np.zeros(nnz.value, np.int32)
        self._ck(self.L.b200s_build_graph(self.h, n_row_nodes, _ptr(node_row_lid), _ptr(node_col_lid),
                                          _ptr(rowptr), _ptr(colind), C.byref(nnz)))
        return nnz.value, rowptr, colind

    def set_halo(self, peers, send_row_nodes, recv_col_nodes):
        """send_row_nodes: per peer row-NODE indices (row lid of dof 0 = 3*index)."""
        n = len(peers)
        peer = np.ascontiguousarray(peers, np.int32)
        so = np.zeros(n + 1, np.int64)
        ro = np.zeros(n + 1, np.int64)
        for i in range(n):
            so[i + 1] = so[i] + len(send_row_nodes[i])
            ro[i + 1] = ro[i] + len(recv_col_nodes[i])
        s = (3 * np.concatenate(send_row_nodes).astype(np.int64)).astype(np.int32) if n and so[-1] else np.zeros(1, np.int32)
        r = np.concatenate(recv_col_nodes).astype(np.int32) if n and ro[-1] else np.zeros(1, np.int32)
        self._ck(self.L.b200s_set_halo(self.h, n, _ptr(peer), _ptr(so), _ptr(np.ascontiguousarray(s)),
                                       _ptr(ro), _ptr(np.ascontiguousarray(r))))

    def set_row_export(self, plan):
        """plan: fe_export.FePlan (b200s_set_row_export)."""
        n = len(plan.peers)
        peer = np.ascontiguousarray(plan.peers if n else [0], np.int32)
        sv = np.ascontiguousarray(np.asarray(plan.send_val if n else [(0, 0)], np.int64).ravel())
        sf = np.ascontiguousarray(np.asarray(plan.send_f if n else [(0, 0)], np.int64).ravel())
        ro = np.zeros(n + 1, np.int64)
        fo = np.zeros(n + 1, np.int64)
        for i in range(n):
            ro[i + 1] = ro[i] + len(plan.recv_pos[i])
            fo[i + 1] = fo[i] + len(plan.recv_fpos[i])
        rp = np.ascontiguousarray(np.concatenate(plan.recv_pos) if n and ro[-1] else np.zeros(1), np.int64)
        rf = np.ascontiguousarray(np.concatenate(plan.recv_fpos) if n and fo[-1] else np.zeros(1), np.int32)
        self.L.b200s_set_row_export.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        self._ck(self.L.b200s_set_row_export(self.h, n, _ptr(peer), _ptr(sv), _ptr(sf), _ptr(ro), _ptr(rp), _ptr(fo), _ptr(rf)))

    def add_exported(self, peer_index, values=None, f=None):
        values = None if values is None else np.ascontiguousarray(values, np.float64)
        f = None if f is None else np.ascontiguousarray(f, np.float64)
        self.L.b200s_add_exported.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self._ck(self.L.b200s_add_exported(self.h, int(peer_index), _ptr(values), _ptr(f)))

    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        self._ck(self.L.b200s_nccl_unique_id(self.h, C.addressof(buf)))
        return bytes(buf)

    def comm_init_nccl(self, uid):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._ck(self.L.b200s_comm_init_nccl(self.h, C.addressof(buf)))

    def evaluate(self, action, u_row=None, K_values=None, f_row=None, check=True):
        """Synchronous evaluate.  Pointers may be numpy arrays, torch tensors or None."""
        err = ErrorInfo()
        rc = self.L.b200s_evaluate(self.h, action, _ptr(u_row), _ptr(K_values), _ptr(f_row), C.byref(err))
        if check:
            self._ck(rc)
        return rc, err

    def evaluate_async(self, action):
        self._ck(self.L.b200s_evaluate_async(self.h, action))

    def synchronize(self, check=True):
        err = ErrorInfo()
        rc = self.L.b200s_synchronize(self.h, C.byref(err))
        if check:
            self._ck(rc)
        return rc, err

    @property
    def nnz(self):
        return self.L.b200s_nnz(self.h)

    @property
    def n_rows(self):
        return self.L.b200s_n_rows(self.h)

    def device_ptrs(self):
        return dict(values=self.L.b200s_device_values(self.h), f=self.L.b200s_device_f(self.h),
                    u_row=self.L.b200s_device_u_row(self.h), u_col=self.L.b200s_device_u_col(self.h),
                    stream=self.L.b200s_stream(self.h))

    def timers(self):
        t = np.zeros(6)
        self.L.b200s_get_timers(self.h, _ptr(t))
        return dict(halo=t[0], zero=t[1], kernel=t[2], device=t[3], h2d=t[4], d2h=t[5])

    def launch_count(self):
        return self.L.b200s_launch_count(self.h)

    def kernel_time(self, reset=False):
        """(sum_ms, count) of the element kernel launches harvested so far."""
        s, n = C.c_double(0), C.c_int64(0)
        self._ck(self.L.b200s_kernel_time(self.h, C.byref(s), C.byref(n), int(reset)))
        return s.value, n.value


def microbench(which, size=0, device=0):
    L = load()
    r = C.c_double(0)
    rc = L.b200s_microbench(device, which, size, C.byref(r))
    if rc != OK:
        raise B200SolidError(rc, f"microbench {which} failed")
    return r.value


def measure_fp64_peak(device=0):
    L = load()
    tf, mhz = C.c_double(0), C.c_double(0)
    rc = L.b200s_measure_fp64_peak(device, C.byref(tf), C.byref(mhz))
    if rc != OK:
        raise B200SolidError(rc, "fp64 peak measurement failed")
    return tf.value, mhz.value

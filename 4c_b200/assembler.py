"""Host-side mirror of the reference call sequence around the hot path, for one rank.

    Solid::ModelEvaluator::Structure::apply_force_stiff_internal   (model_evaluator_structure.cpp:442-472)
      discret().set_state("displacement", u)                       -> DeviceAssembler.set_state
      discret().evaluate(p{action}, stiff, nullptr, fint, ...)     -> DeviceAssembler.evaluate
        AssembleStrategy::zero / assemble_matrix1 / assemble_vector1 (4C_fem_general_assemblestrategy.cpp)

Everything numeric happens in lib4c_b200_solid.so; this class only owns the context and keeps the
names and argument meaning of the reference interface so that tests read like the reference's.
"""
import numpy as np

from . import capi

ACTIONS = {"struct_calc_nlnstiff": capi.ACTION_NLNSTIFF,
           "struct_calc_internalforce": capi.ACTION_INTERNALFORCE,
           "struct_calc_nlnstiffmass": capi.ACTION_NLNSTIFFMASS,
           "struct_calc_nlnstifflmass": capi.ACTION_NLNSTIFFLMASS,
           "struct_calc_internalinertiaforce": capi.ACTION_INTERNALINERTIAFORCE}
STIFF_ACTIONS = (capi.ACTION_NLNSTIFF, capi.ACTION_NLNSTIFFMASS, capi.ACTION_NLNSTIFFLMASS)
NUM_GP = {capi.HEX8: 8, capi.HEX27: 27, capi.TET10: 4, capi.TET4: 1, capi.HEX20: 27, capi.WEDGE6: 6, capi.PYRAMID5: 8, capi.HEX18: 18}


class DeviceAssembler:
    def __init__(self, rank_mesh, matkind, matparams, device=0, per_ele_param=None, graph=None,
                 tech=capi.TECH_NONE, density=None, prestress=capi.PRESTRESS_NONE, fibers=None,
                 kinem=capi.KINEM_NONLINEAR_TOTLAG, options=None):
        """rank_mesh: partition.RankMesh.  graph: (rowptr, colind) from the reference's matrix, or
        None to build the FillComplete-equivalent graph on the device.  tech: element technology of
        the block (TECH key of the SOLID element line).  matkind MAT_ELASTHYPER: matparams is the
        summand list [(name, [params]), ...].  density: DENS of the material (mass actions)."""
        rm = self.rm = rank_mesh
        self.ctx = capi.Context(device, rm.rank, rm.nranks)
        for k, v in (options or {}).items():   # library options that act at set_mesh / set_graph (b200s_set_option)
            self.ctx.set_option(k, v)
        self.ctx.set_mesh(rm.celltype, rm.X, rm.conn)
        if tech != capi.TECH_NONE or prestress != capi.PRESTRESS_NONE or kinem != capi.KINEM_NONLINEAR_TOTLAG:
            self.ctx.set_element_properties(kinem=kinem, tech=tech, prestress=prestress)
        if matkind == capi.MAT_ELASTHYPER:
            self.ctx.set_material_elasthyper(matparams)
        else:
            self.ctx.set_material(matkind, matparams, per_ele_param)
        if fibers is not None:  # FIBER1 per column element
            self.ctx.set_element_fibers(fibers)
        if density is not None:
            self.ctx.set_density(density)
        if getattr(rm, "ele_is_row", None) is not None:
            self.ctx.set_row_elements(rm.ele_is_row)
        if graph is None:
            self.nnz = self.ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid)
        else:
            rowptr, colind = graph
            d = np.arange(3, dtype=np.int32)
            row = np.where(rm.node_row_lid[:, None] >= 0, 3 * rm.node_row_lid[:, None] + d, -1)
            col = 3 * rm.node_col_lid[:, None] + d
            self.ctx.set_graph(rowptr, colind, row.astype(np.int32).ravel(), col.astype(np.int32).ravel())
            self.nnz = self.ctx.nnz
        self.n_rows = 3 * rm.n_row_nodes
        if rm.nranks > 1:
            self.ctx.set_halo(rm.peers, rm.send_row_nodes, rm.recv_col_nodes)
        self._u = None

    def graph(self):
        """(rowptr, colind) of the installed device graph (host copies)."""
        rm = self.rm
        _, rowptr, colind = self.ctx.build_graph(rm.n_row_nodes, rm.node_row_lid, rm.node_col_lid,
                                                 fetch=True)
        return rowptr, colind

    def set_state(self, name, u_row):
        v = np.ascontiguousarray(u_row, np.float64).ravel()
        assert v.size == self.n_rows
        if name == "displacement":
            self._u = v
        self.ctx.set_state(name, v)

    def evaluate(self, action="struct_calc_nlnstiff", check=True):
        """Returns (K_values or None, f, error_info) as host arrays."""
        act = ACTIONS[action]
        K = np.empty(self.nnz) if act in STIFF_ACTIONS else None
        f = np.empty(self.n_rows)
        rc, err = self.ctx.evaluate(act, self._u, K, f, check=check)
        return K, f, err

    def evaluate_resident(self, action="struct_calc_nlnstiff", check=True):
        """Like evaluate, but K stays in the context's device CSR (for apply_dirichlet / multiply /
        solve_cg): returns (None, f, error_info); only vectors cross the ABI."""
        f = np.empty(self.n_rows)
        rc, err = self.ctx.evaluate(ACTIONS[action], self._u, None, f, check=check)
        return None, f, err

    def recover(self, du_row, step_length=1.0):
        """struct_calc_recover with the Newton direction as "residual displacement" (EAS alpha update)."""
        self.ctx.recover(np.ascontiguousarray(du_row, np.float64).ravel(), step_length)

    def update_prestress(self):
        """struct_update_prestress for the current displacement state (PRESTRESS_TECH mulf)."""
        self.ctx.update_prestress()

    def mass(self, want_matrix=True):
        """(M_values, f_inertia) of the last mass action."""
        return self.ctx.get_mass(want_matrix=want_matrix)

    def calc_stress(self, iostress="2pk", iostrain="gl"):
        """struct_calc_stress: (stress, strain) arrays [n_col_eles, num_gp, 6]."""
        st = {"none": capi.STRESS_NONE, "2pk": capi.STRESS_2PK, "cauchy": capi.STRESS_CAUCHY}[iostress]
        en = {"none": capi.STRAIN_NONE, "gl": capi.STRAIN_GL, "ea": capi.STRAIN_EA}[iostrain]
        return self.ctx.calc_stress(NUM_GP[self.rm.celltype], st, en)

    def calc_energy(self):
        """struct_calc_energy: internal energy of this rank's row elements."""
        return self.ctx.calc_energy()

    def add_surface_neumann(self, neumann_type, face_nodes, onoff, val):
        """One DESIGN SURF NEUMANN CONDITIONS entry; face_nodes in column-node lids of this rank
        (every face of the condition geometry all of whose nodes are column nodes here). Returns its id."""
        t = {"live": capi.NEUMANN_LIVE, "pseudo_orthopressure": capi.NEUMANN_PSEUDO_ORTHOPRESSURE,
             "orthopressure": capi.NEUMANN_ORTHOPRESSURE}[neumann_type.lower()]
        return self.ctx.add_surface_neumann(t, face_nodes, onoff, val)

    def add_volume_neumann(self, onoff, val, ele_ids=None):
        return self.ctx.add_volume_neumann(onoff, val, ele_ids)

    def evaluate_neumann(self, u_row=None, loadlin=False, mulf_active=False):
        """Discretization::evaluate_neumann on the block: external force on the row map; with loadlin the
        orthopressure load stiffness is added to the stiffness values resident in the context."""
        flags = (capi.NEUMANN_LOADLIN if loadlin else 0) | (capi.NEUMANN_MULF_ACTIVE if mulf_active else 0)
        return self.ctx.evaluate_neumann(u_row, flags)

    def close(self):
        self.ctx.close()


def parity_report(rowptr, values, ref_values, f, ref_f, rtol=1e-12):
    """Parity norm of SURVEY.md 8(d): per entry |a-b| <= rtol * max(|a|,|b|,s) with s the max-abs of
    the entry's node block row (3 dof rows of one node) for K, of the whole vector for f.
    Returns (ok, worst_ratio_K, worst_ratio_f) where ratio = |a-b| / (rtol * scale)."""
    rowptr = np.asarray(rowptr)
    n_rows = len(rowptr) - 1
    rowmax = np.maximum.reduceat(np.maximum(np.abs(values), np.abs(ref_values)),
                                 rowptr[:-1].clip(max=len(values) - 1))
    rowmax[np.diff(rowptr) == 0] = 0.0
    blk = np.repeat(rowmax.reshape(-1, 3).max(axis=1), 3)
    scale = np.repeat(blk, np.diff(rowptr))
    assert len(scale) == len(values) and n_rows % 3 == 0
    dK = np.abs(values - ref_values)
    rK = float((dK / (rtol * np.maximum(scale, 1e-300))).max()) if len(values) else 0.0
    sf = max(np.abs(f).max(), np.abs(ref_f).max(), 1e-300)
    rf = float((np.abs(f - ref_f) / (rtol * sf)).max())
    finite = np.isfinite(values).all() and np.isfinite(f).all()
    return bool(finite and rK <= 1.0 and rf <= 1.0), rK, rf


class RowExportAssembler:
    """Row-element mode (FE_MATRIX: SparseMatrix::fe_assemble + complete(), 4C_linalg_sparsematrix.cpp:598-630,705-710):
    the rank evaluates its ROW elements only, contributions to rows of ghost nodes go to export rows behind its own
    and travel to their owners after the element kernels (b200s_set_row_export).  view / rowptr / colind / plan come
    from fe_export.fe_plan.  evaluate() returns the rank's own rows; with a communicator the export has happened
    inside, without one (tests) exported() / add_exported() move the rows by hand."""

    def __init__(self, view, rowptr, colind, plan, matkind, matparams, device=0, options=None):
        rm = self.rm = view.rm
        self.view, self.plan = view, plan
        self.ctx = capi.Context(device, rm.rank, rm.nranks)
        opts = {"owned_rows": plan.n_rows_owned, "owned_nnz": plan.nnz_owned}
        opts.update(options or {})
        for k, v in opts.items():
            self.ctx.set_option(k, v)
        self.ctx.set_mesh(rm.celltype, rm.X, view.conn)
        self.ctx.set_material(matkind, matparams)
        d = np.arange(3, dtype=np.int32)
        ext = view.node_row_lid_ext
        row = np.where(ext[:, None] >= 0, 3 * ext[:, None] + d, -1)
        col = 3 * rm.node_col_lid[:, None] + d
        self.ctx.set_graph(rowptr, colind, row.astype(np.int32).ravel(), col.astype(np.int32).ravel())
        self.ctx.set_row_export(plan)
        self.n_rows, self.nnz = plan.n_rows_owned, plan.nnz_owned
        self.n_rows_ext, self.nnz_ext = 3 * view.n_row_nodes_ext, int(rowptr[-1])
        if rm.nranks > 1:
            self.ctx.set_halo(rm.peers, rm.send_row_nodes, rm.recv_col_nodes)
        self._u = None

    def set_state(self, name, u_row):
        v = np.ascontiguousarray(u_row, np.float64).ravel()
        assert v.size == self.n_rows
        self._u = v
        self.ctx.set_state(name, v)

    def evaluate(self, action="struct_calc_nlnstiff", check=True):
        act = ACTIONS[action]
        K = np.empty(self.nnz) if act in STIFF_ACTIONS else None
        f = np.empty(self.n_rows)
        rc, err = self.ctx.evaluate(act, self._u, K, f, check=check)
        return K, f, err

    def exported(self, peer_index):
        """(values, f) slices of the export rows for peer peer_index as they stand in the device buffers."""
        (v0, nv), (f0, nf) = self.plan.send_val[peer_index], self.plan.send_f[peer_index]
        vals = self.ctx.get_values()[v0:v0 + nv] if nv else np.zeros(0)
        f = self.ctx.get_f_rows(f0, f0 + nf) if nf else np.zeros(0)
        return vals, f

    def own_rows(self):
        """(values, f) of the rank's own rows as they stand in the device buffers."""
        return self.ctx.get_values()[:self.nnz], self.ctx.get_f_rows(0, self.n_rows)

    def close(self):
        self.ctx.close()

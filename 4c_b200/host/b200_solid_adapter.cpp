// b200_solid_adapter.cpp -- see b200_solid_adapter.hpp.  Host C++ only; every numeric step is a call
// into lib4c_b200_solid.so through the C ABI.
#include "b200_solid_adapter.hpp"

#include <cstring>

namespace B200Solid
{
  DeviceElementBlock::DeviceElementBlock(int device, int rank, int nranks)
  {
    const int rc = b200s_create(device, rank, nranks, &ctx_);
    if (rc != B200S_OK)
      throw Exception(rc, "b200s_create failed: no CUDA device (the device path has no CPU fallback)");
  }

  DeviceElementBlock::~DeviceElementBlock() { b200s_destroy(ctx_); }

  void DeviceElementBlock::check(int rc) const
  {
    if (rc == B200S_OK) return;
    const std::string msg = b200s_last_error(ctx_);
    if (rc == B200S_ERR_NEG_JACOBIAN)
      // message of FOUR_C_ASSERT_ALWAYS in 4C_solid_ele_calc_lib.hpp:558
      throw Exception(rc, "Determinant of jacobian is smaller or equal to zero at one node of the element: " + msg);
    throw Exception(rc, msg);
  }

  bool DeviceElementBlock::supported(int celltype, int material_kind)
  {
    const bool cell = celltype == B200S_HEX8 || celltype == B200S_HEX27 || celltype == B200S_TET10;
    const bool mat = material_kind == B200S_MAT_COUP_NEOHOOKE || material_kind == B200S_MAT_ISO_VOL ||
                     material_kind == B200S_MAT_SVK || material_kind == B200S_MAT_ELASTHYPER;
    return cell && mat;
  }

  bool DeviceElementBlock::supported(int celltype, int kinem, int tech, int prestress, int material_kind)
  {
    // general kernel: displacement based (+ MULF), pyramid5 also F-bar (+ MULF)
    const bool general_cell = celltype == B200S_TET4 || celltype == B200S_HEX20 || celltype == B200S_WEDGE6 || celltype == B200S_PYRAMID5 ||
                              celltype == B200S_HEX18;
    const bool cell = celltype == B200S_HEX8 || celltype == B200S_HEX27 || celltype == B200S_TET10 || general_cell;
    if (!cell) return false;
    if (general_cell && tech != B200S_TECH_NONE && !(celltype == B200S_PYRAMID5 && tech == B200S_TECH_FBAR)) return false;
    const bool hyper = material_kind == B200S_MAT_COUP_NEOHOOKE || material_kind == B200S_MAT_ISO_VOL ||
                       material_kind == B200S_MAT_SVK || material_kind == B200S_MAT_ELASTHYPER;
    if (kinem == B200S_KINEM_LINEAR)
    {
      // small strains: displacement based (the hyperelastic kinds and the history material) or hex8 EAS
      // (EASFormulation<hex8, h8_9 | h8_21, linear>)
      if (prestress != B200S_PRESTRESS_NONE) return false;
      if (tech == B200S_TECH_EAS_MILD || tech == B200S_TECH_EAS_FULL) return celltype == B200S_HEX8 && hyper;
      if (tech != B200S_TECH_NONE) return false;
      return hyper || material_kind == B200S_MAT_DRUCKER_PRAGER;
    }
    if (kinem != B200S_KINEM_NONLINEAR_TOTLAG) return false;
    // finite-strain history material: displacement based (all three cells) or EAS (hex8), no prestress
    if (material_kind == B200S_MAT_PLASTIC_NLNLOG)  // every cell (general kernel beyond hex8), EAS on hex8
      return prestress == B200S_PRESTRESS_NONE &&
             (tech == B200S_TECH_NONE || (celltype == B200S_HEX8 && (tech == B200S_TECH_EAS_MILD || tech == B200S_TECH_EAS_FULL)));
    if (material_kind == B200S_MAT_VISCONEOHOOKE)
      return prestress == B200S_PRESTRESS_NONE &&
             (tech == B200S_TECH_NONE || (celltype == B200S_HEX8 && (tech == B200S_TECH_EAS_MILD || tech == B200S_TECH_EAS_FULL)));
    if (!hyper) return false;
    // F-bar: hex8 and pyramid5, also with MULF (FbarEvaluators, FBarMulfEvaluators: 4C_solid_ele_factory.hpp:59-60,79-81)
    const bool fbar_cell = celltype == B200S_HEX8 || celltype == B200S_PYRAMID5;
    if (prestress == B200S_PRESTRESS_MULF) return tech == B200S_TECH_NONE || (tech == B200S_TECH_FBAR && fbar_cell);
    if (prestress != B200S_PRESTRESS_NONE) return false;
    if (tech == B200S_TECH_NONE) return true;
    if (tech == B200S_TECH_FBAR) return fbar_cell;
    if (tech == B200S_TECH_SHELL_ANS || tech == B200S_TECH_SHELL_EAS || tech == B200S_TECH_SHELL_EAS_ANS)
      return celltype == B200S_HEX8;  // the wedge6 solid shells: CPU loop
    // EAS is a hex8 technology (4C_solid_ele_factory_lib.hpp)
    return celltype == B200S_HEX8 && (tech == B200S_TECH_EAS_MILD || tech == B200S_TECH_EAS_FULL);
  }

  void DeviceElementBlock::attach_to(DeviceElementBlock& primary)
  {
    check(b200s_attach_block(ctx_, primary.ctx_));
    attached_ = true;
  }

  double* DeviceElementBlock::alloc_staging(int device, std::int64_t n_doubles)
  {
    return static_cast<double*>(b200s_host_alloc_near(device, 8 * n_doubles, nullptr));
  }
  void DeviceElementBlock::free_staging(double* p) { b200s_host_free(p); }

  void DeviceElementBlock::set_block(const ElementBlockView& b)
  {
    check(b200s_set_mesh(ctx_, b.celltype, b.n_col_nodes, b.X, b.n_col_eles, b.conn));
  }

  void DeviceElementBlock::set_material(int kind, const std::vector<double>& params, const double* per_element)
  {
    check(b200s_set_material(ctx_, kind, params.data(), (int)params.size(), per_element));
  }

  void DeviceElementBlock::set_graph(const ElementBlockView& b, const CrsView& crs)
  {
    // Epetra's local CRS uses int offsets; the ABI takes 64-bit offsets (configs beyond 2^31 nnz)
    std::vector<std::int64_t> rowptr(crs.n_rows + 1);
    for (std::int64_t r = 0; r <= crs.n_rows; ++r) rowptr[r] = crs.index_offset[r];
    static_assert(sizeof(int) == sizeof(std::int32_t), "Epetra local ids are 32 bit");
    check(b200s_set_graph(ctx_, crs.n_rows, rowptr.data(), reinterpret_cast<const std::int32_t*>(crs.indices),
        b.dof_row_lid, b.dof_col_lid));
  }

  void DeviceElementBlock::set_import(const ImportView& imp)
  {
    const int n = (int)imp.peers.size();
    std::vector<std::int32_t> peer(imp.peers.begin(), imp.peers.end()), s, r;
    std::vector<std::int64_t> so(n + 1, 0), ro(n + 1, 0);
    for (int i = 0; i < n; ++i)
    {
      s.insert(s.end(), imp.send_row_lids[i].begin(), imp.send_row_lids[i].end());
      r.insert(r.end(), imp.recv_col_nodes[i].begin(), imp.recv_col_nodes[i].end());
      so[i + 1] = (std::int64_t)s.size(), ro[i + 1] = (std::int64_t)r.size();
    }
    if (s.empty()) s.push_back(0);
    if (r.empty()) r.push_back(0);
    check(b200s_set_halo(ctx_, n, peer.data(), so.data(), s.data(), ro.data(), r.data()));
  }

  void DeviceElementBlock::set_row_export(const RowExportView& e)
  {
    const int n = (int)e.peers.size();
    std::vector<std::int32_t> peer(e.peers.begin(), e.peers.end()), fpos;
    std::vector<std::int64_t> sv(2 * (size_t)n + 2, 0), sf(2 * (size_t)n + 2, 0), ro(n + 1, 0), fo(n + 1, 0), pos;
    for (int i = 0; i < n; ++i)
    {
      sv[2 * i] = e.send_val_start[i], sv[2 * i + 1] = e.send_val_count[i];
      sf[2 * i] = e.send_f_start[i], sf[2 * i + 1] = e.send_f_count[i];
      pos.insert(pos.end(), e.recv_pos[i].begin(), e.recv_pos[i].end());
      fpos.insert(fpos.end(), e.recv_fpos[i].begin(), e.recv_fpos[i].end());
      ro[i + 1] = (std::int64_t)pos.size(), fo[i + 1] = (std::int64_t)fpos.size();
    }
    if (peer.empty()) peer.push_back(0);
    if (pos.empty()) pos.push_back(0);
    if (fpos.empty()) fpos.push_back(0);
    check(b200s_set_row_export(ctx_, n, peer.data(), sv.data(), sf.data(), ro.data(), pos.data(), fo.data(), fpos.data()));
  }

  std::string DeviceElementBlock::nccl_unique_id()
  {
    std::string id(128, '\0');
    check(b200s_nccl_unique_id(ctx_, id.data()));
    return id;
  }

  void DeviceElementBlock::init_communicator(const char* id) { check(b200s_comm_init_nccl(ctx_, id)); }

  void DeviceElementBlock::evaluate(Action action, double* crs_values, double* f_row)
  {
    // an attached block reads the state its primary imported in this iteration
    if (!u_row_ && !attached_) throw Exception(B200S_ERR_INVALID, "set_state(\"displacement\") must precede evaluate");
    b200s_error_info info;
    std::memset(&info, 0, sizeof info);
    check(b200s_evaluate(ctx_, (int)action, attached_ ? nullptr : u_row_, crs_values, f_row, &info));
  }

  void DeviceElementBlock::set_material_elasthyper(const std::vector<std::int32_t>& kinds,
      const std::vector<std::int32_t>& modes, const std::vector<double>& params4, double density)
  {
    check(b200s_set_material_elasthyper(ctx_, (int)kinds.size(), kinds.data(), modes.empty() ? nullptr : modes.data(),
        params4.data()));
    if (density > 0.0) check(b200s_set_density(ctx_, density));
  }

  void DeviceElementBlock::set_element_properties(int kinem, int tech, int prestress)
  {
    check(b200s_set_element_properties(ctx_, kinem, tech, prestress));
  }

  void DeviceElementBlock::set_row_elements(const std::uint8_t* is_row) { check(b200s_set_row_elements(ctx_, is_row)); }

  void DeviceElementBlock::set_state(const char* name, const double* v_row)
  {
    if (!std::strcmp(name, "displacement")) u_row_ = v_row;
    check(b200s_set_state(ctx_, name, v_row));
  }

  void DeviceElementBlock::get_mass(double* mass_values, double* f_inertia)
  {
    check(b200s_get_mass(ctx_, mass_values, f_inertia));
  }

  void DeviceElementBlock::calc_stress(int stress_type, int strain_type, double* stress, double* strain)
  {
    check(b200s_calc_stress(ctx_, stress_type, strain_type, stress, strain));
  }

  double DeviceElementBlock::calc_energy()
  {
    double e = 0.0;
    check(b200s_calc_energy(ctx_, &e));
    return e;
  }

  void DeviceElementBlock::update_prestress() { check(b200s_update_prestress(ctx_)); }

  void DeviceElementBlock::recover(const double* du_row, double step_length)
  {
    check(b200s_recover(ctx_, du_row, step_length));
  }

  int DeviceElementBlock::add_surface_neumann(b200s_neumann_type type, std::int64_t n_faces,
      const std::int32_t* face_nodes, const std::int32_t* onoff, const double* val)
  {
    const int id = b200s_add_surface_neumann(ctx_, type, n_faces, face_nodes, onoff, val);
    if (id < 0) check(id);
    return id;
  }

  int DeviceElementBlock::add_volume_neumann(std::int64_t n_eles, const std::int32_t* ele_ids,
      const std::int32_t* onoff, const double* val)
  {
    const int id = b200s_add_volume_neumann(ctx_, n_eles, ele_ids, onoff, val);
    if (id < 0) check(id);
    return id;
  }

  void DeviceElementBlock::set_neumann_factors(int cond, const double* dof_fac, const double* gp_fac)
  {
    check(b200s_set_neumann_factors(ctx_, cond, dof_fac, gp_fac));
  }

  void DeviceElementBlock::evaluate_neumann(double* fext_row, bool loadlin, bool mulf_active, const double* u_row)
  {
    const int flags = (loadlin ? B200S_NEUMANN_LOADLIN : 0) | (mulf_active ? B200S_NEUMANN_MULF_ACTIVE : 0);
    check(b200s_evaluate_neumann(ctx_, u_row, flags, fext_row));
  }

  void DeviceElementBlock::update_step() { check(b200s_update_step(ctx_)); }
  void DeviceElementBlock::reset_step() { check(b200s_reset_step(ctx_)); }
  std::int64_t DeviceElementBlock::material_history_size() const { return b200s_material_history_size(ctx_); }
  void DeviceElementBlock::get_material_history(double* hist) { check(b200s_get_material_history(ctx_, hist)); }
  void DeviceElementBlock::set_material_history(const double* hist) { check(b200s_set_material_history(ctx_, hist)); }

  void DeviceElementBlock::apply_dirichlet(const double* dbctoggle_row, bool diagonalblock)
  {
    check(b200s_apply_dirichlet(ctx_, B200S_MATRIX_STIFFNESS, dbctoggle_row, diagonalblock ? 1 : 0));
  }

  void DeviceElementBlock::multiply(const double* x_row, double* y_row)
  {
    check(b200s_multiply(ctx_, B200S_MATRIX_STIFFNESS, x_row, y_row));
  }

  std::int64_t DeviceElementBlock::nnz() const { return b200s_nnz(ctx_); }
  std::int64_t DeviceElementBlock::n_rows() const { return b200s_n_rows(ctx_); }
}  // namespace B200Solid

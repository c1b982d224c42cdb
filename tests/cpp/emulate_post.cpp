// TEST INFRASTRUCTURE: runs the PRODUCT's __host__ __device__ point functions of
// 4c_b200/csrc/post_kernels.cuh, summands.cuh and solid_math.cuh on the CPU (one call per thread
// the kernels would launch), so the CPU suite checks the stress / energy / mass arithmetic, the
// generic MAT_ElastHyper summands and the mass scatter addressing against the oracle without a
// GPU.  Not shipped, not a fallback: the library has no host execution path.
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../4c_b200/csrc/post_kernels.cuh"
#include "../../4c_b200/csrc/generic_kernel.cuh"
#include "../../4c_b200/csrc/shape_tables.hpp"

using namespace b200s;

namespace
{
  // matkind 0..2: params = raw input parameters (as b200s_set_material);
  // matkind 3   : params = {n, n x {kind, mode, p0..p3}} (the oracle's packed block)
  // the oracle's block may carry its fibre material as {100: CoupAnisoExpo}{101: fibre vector}: the product's
  // summand id is SUM_COUP_ANISOEXPO and the (normalised) fibre goes through MatParams::fiber
  double g_fiber[3];
  bool setup_material(int matkind, const double* params, MatParams& mp, HyperMat& hm)
  {
    if (matkind != MAT_HYPER) return make_mat_params(matkind, params, mp);
    for (double& v : mp.p) v = 0.0;
    mp.fiber = nullptr;
    const int n = (int)params[0];
    hm.n = 0, hm.uses_modified = 0;
    for (int k = 0; k < n; ++k)
    {
      const double* q = params + 1 + 6 * k;
      int kind = (int)q[0];
      if (kind == 101)
      {
        const double nrm = std::sqrt(q[2] * q[2] + q[3] * q[3] + q[4] * q[4]);
        for (int d = 0; d < 3; ++d) g_fiber[d] = q[2 + d] / nrm;
        mp.fiber = g_fiber;
        continue;
      }
      if (kind == 100) kind = SUM_COUP_ANISOEXPO;
      if (hm.n >= HYPER_MAX_SUMMANDS) return false;
      if (!make_summand(kind, q + 2, (int)q[1], hm.s[hm.n])) return false;
      if (summand_is_modified(kind)) hm.uses_modified = 1;
      hm.n += 1;
    }
    mp.hyper = &hm;
    return true;
  }

  template <int MAT>
  void run_post(const PostArgs& A, double* energy)
  {
    double sum = 0.0;
    for (int64_t e = 0; e < A.n_ele; ++e)
      for (int g = 0; g < A.ngp; ++g)
      {
        const int64_t eid = A.ele_id ? A.ele_id[e] : e;
        const bool row = A.ele_is_row == nullptr || A.ele_is_row[eid] != 0;
        double s6[6], e6[6];
        sum += post_point<MAT>(A, e, g, energy != nullptr && row, A.stress ? s6 : nullptr,
            A.strain ? e6 : nullptr);
        for (int v = 0; v < 6; ++v)
        {
          if (A.stress && A.stress_type != STRESS_NONE) A.stress[(eid * A.ngp + g) * 6 + v] = s6[v];
          if (A.strain && A.strain_type != STRAIN_NONE) A.strain[(eid * A.ngp + g) * 6 + v] = e6[v];
        }
      }
    if (energy) *energy = sum;
  }

  template <int NN>
  void run_mass(const MassArgs& A)
  {
    for (int64_t e = 0; e < A.n_ele; ++e)
      for (int i = 0; i < NN; ++i)
      {
        double m[NN];
        mass_row<NN>(A, e, i, m);
        mass_scatter<NN>(A, e, i, m);
      }
  }
}  // namespace

// Mat::So3Material::evaluate + strain_energy of the product at one point: C (Voigt) in,
// S[6], Cm[36] (full symmetric 6x6, row-major = column-major) and psi out
extern "C" int emu_material(int matkind, const double* params, const double* C, double* S,
    double* Cm36, double* psi)
{
  MatParams mp;
  HyperMat hm;
  if (!setup_material(matkind, params, mp, hm)) return -3;
  double Cm[21];
  switch (matkind)
  {
    case 0:
      material_evaluate<MAT_COUP_NEOHOOKE>(mp, C, S, Cm), *psi = material_strain_energy<MAT_COUP_NEOHOOKE>(mp, C);
      break;
    case 1:
      material_evaluate<MAT_ISO_VOL>(mp, C, S, Cm), *psi = material_strain_energy<MAT_ISO_VOL>(mp, C);
      break;
    case 2:
      material_evaluate<MAT_SVK>(mp, C, S, Cm), *psi = material_strain_energy<MAT_SVK>(mp, C);
      break;
    default:
      material_evaluate<MAT_HYPER>(mp, C, S, Cm), *psi = material_strain_energy<MAT_HYPER>(mp, C);
  }
  for (int I = 0; I < 6; ++I)
    for (int J = 0; J < 6; ++J) Cm36[I + 6 * J] = Cm[c21(I, J)];
  return 0;
}

// MAT_VISCONEOHOOKE of the product at one point: in6 = {YOUNGS_SLOW POISSON DENS YOUNGS_FAST RELAX THETA}, dt = time
// step size, hist = VISCO_HIST doubles (updated in place) or null; -3 for parameters setup() rejects
extern "C" int emu_visco_neohooke(const double* in6, double dt, double* hist, const double* C, double* S, double* Cm36)
{
  MatParams mp;
  for (double& v : mp.p) v = 0.0;
  if (!make_visco_neohooke(in6, dt, mp.p)) return -3;
  mp.hyper = nullptr, mp.fiber = nullptr, mp.ext = nullptr, mp.hist = hist;
  double Cm[21];
  material_evaluate<MAT_VISCO>(mp, C, S, Cm);
  for (int I = 0; I < 6; ++I)
    for (int J = 0; J < 6; ++J) Cm36[I + 6 * J] = Cm[c21(I, J)];
  return 0;
}

// MAT_Struct_PlasticNlnLogNeoHooke of the product at one point: in10 = {YOUNG NUE DENS YIELD ISOHARD SATHARDENING HARDEXPO
// VISC RATE_DEPENDENCY TOL}, dt = time step size, hist = PLASTIC_HIST doubles (updated in place) or null
extern "C" int emu_plastic_nlnlog(const double* in10, double dt, double* hist, const double* C, double* S, double* Cm36)
{
  double q[PLASTIC_NPAR];
  if (!make_plastic_nlnlog(in10, dt, q)) return -3;
  MatParams mp;
  for (double& v : mp.p) v = 0.0;
  mp.hyper = nullptr, mp.fiber = nullptr, mp.ext = q, mp.hist = hist;
  double Cm[21];
  material_evaluate<MAT_PLASTIC>(mp, C, S, Cm);
  for (int I = 0; I < 6; ++I)
    for (int J = 0; J < 6; ++J) Cm36[I + 6 * J] = Cm[c21(I, J)];
  return 0;
}

// MAT_Struct_DruckerPrager of the product at one point: in10 = input-file parameters {YOUNG NUE ISOHARD TOL C ETA
// XI ETABAR MAXITER TANG}, hist = DP_HIST doubles (updated in place) or null; returns -3 for parameters the device
// rejects (non-symmetric consistent tangent)
extern "C" int emu_drucker_prager(const double* in10, double* hist, const double* C, double* S, double* Cm36)
{
  double q[DP_NPAR];
  if (!make_drucker_prager(in10, q)) return -3;
  MatParams mp;
  for (double& v : mp.p) v = 0.0;
  mp.hyper = nullptr, mp.fiber = nullptr, mp.ext = q, mp.hist = hist;
  if (!drucker_prager_tangent_is_symmetric(in10))
  {
    // ETA != ETABAR with the consistent tangent: the general kernel's material call (full 6x6)
    material_evaluate_full<MAT_DP>(mp, C, S, Cm36);
    return 1;
  }
  double Cm[21];
  material_evaluate<MAT_DP>(mp, C, S, Cm);
  for (int I = 0; I < 6; ++I)
    for (int J = 0; J < 6; ++J) Cm36[I + 6 * J] = Cm[c21(I, J)];
  return 0;
}

// The general (non-symmetric tangent) kernel of generic_kernel.cuh stepped on the host: one call of generic_block_row
// per (element, row node) and the kernel's scatter.  celltype 0 hex8 / 1 hex27 / 2 tet10, KINEM linear,
// MAT_Struct_DruckerPrager with its state store (n_ele * ngp * DP_HIST, updated in place).
// host twin of hex8eas::centroid_transformation (a warp routine): T0^-T = T(J0^-1) in strain-like Voigt notation and det J0
static void emu_centroid_t0(const double* X, const int32_t* ce, double* t0)
{
  double J0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ0[9];
  // dN_k/dxi_a(0) = +-1/8: parametric sign of corner k in direction a (hex8 node order)
  auto sgn = [](int k, int a) { return a == 0 ? (((k ^ (k >> 1)) & 1) ? 1.0 : -1.0) : (a == 1 ? ((k & 2) ? 1.0 : -1.0) : ((k & 4) ? 1.0 : -1.0)); };
  for (int k = 0; k < 8; ++k)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) J0[a + 3 * b] = fma(0.125 * sgn(k, a), X[3 * (int64_t)ce[k] + b], J0[a + 3 * b]);
  t0[36] = inv3(iJ0, J0);
  voigt_strain_transformation(iJ0, t0);
}

template <int MAT, int NN>
static int run_generic(const GenericArgs& Ga)
{
  const EvalArgs& A = Ga.A;
  int bad = 0;
  auto scatter = [&](int64_t e, int i, const double (*Krow)[9], const double* fi)
  {
    const NodeRow nr = A.node_row[A.conn[e * Ga.nn + i]];
    if (nr.rowstart < 0) return;
    for (int b = 0; b < 3; ++b) A.f[nr.frow + b] += fi[b];
    if (!Ga.stiff) return;
    const int32_t* pos = A.pos + e * Ga.pos_stride + i * Ga.nn;
    for (int j = 0; j < Ga.nn; ++j)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c) A.vals[nr.rowstart + (int64_t)b * nr.rowlen + pos[j] + c] += Krow[j][b + 3 * c];
  };
  if constexpr (NN == 8)
    if (Ga.dN_ans != nullptr && Ga.eas_hist != nullptr)
    {
      // TECH shell_eas_ans: the kernel's two steps around its __syncwarp -- every row thread of the element integrates and
      // stores, then every row thread condenses
      for (int64_t e = 0; e < A.n_ele; ++e)
      {
        double Krow[8][8][9], fi[8][3], dj = 1e300, df = 1e300;
        AnsEas ea[8];
        double* h = Ga.eas_hist + e * ANS_EAS_HIST;
        for (int i = 0; i < 8; ++i)
        {
          if (generic_block_row_ans<MAT>(Ga, e, i, Krow[i], fi[i], dj, df, &ea[i])) ++bad;
          ans_eas_store(h, i, ea[i]);
        }
        for (int i = 0; i < 8; ++i)
        {
          ans_eas_condense(h, i, ea[i], Ga.stiff != 0, Krow[i], fi[i]);
          scatter(e, i, Krow[i], fi[i]);
        }
      }
      return bad;
    }
  for (int64_t e = 0; e < A.n_ele; ++e)
    for (int i = 0; i < Ga.nn; ++i)
    {
      double Krow[NN][9], fi[3], dj = 1e300, df = 1e300;
      bool neg;
      if constexpr (NN == 8)
        neg = Ga.dN_ans != nullptr ? generic_block_row_ans<MAT>(Ga, e, i, Krow, fi, dj, df)
                                   : generic_block_row<MAT, NN>(Ga, e, i, Krow, fi, dj, df);
      else
        neg = generic_block_row<MAT, NN>(Ga, e, i, Krow, fi, dj, df);
      if (neg) ++bad;
      const NodeRow nr = A.node_row[A.conn[e * Ga.nn + i]];
      if (nr.rowstart < 0) continue;
      for (int b = 0; b < 3; ++b) A.f[nr.frow + b] += fi[b];
      if (!Ga.stiff) continue;
      const int32_t* pos = A.pos + e * Ga.pos_stride + i * Ga.nn;
      for (int j = 0; j < Ga.nn; ++j)
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c) A.vals[nr.rowstart + (int64_t)b * nr.rowlen + pos[j] + c] += Krow[j][b + 3 * c];
    }
  return bad;
}
template <int MAT>
static int run_generic_nn(const GenericArgs& Ga)
{
  switch (Ga.nn)
  {
    case 4: return run_generic<MAT, 4>(Ga);
    case 5: return run_generic<MAT, 5>(Ga);
    case 6: return run_generic<MAT, 6>(Ga);
    case 8: return run_generic<MAT, 8>(Ga);
    case 10: return run_generic<MAT, 10>(Ga);
    case 18: return run_generic<MAT, 18>(Ga);
    case 20: return run_generic<MAT, 20>(Ga);
    default: return run_generic<MAT, 27>(Ga);
  }
}
// The general kernel for the hyperelastic kinds 0..2 (tet4 / hex20 blocks run through it): celltype as in
// shape_tables.hpp (0 hex8, 1 hex27, 2 tet10, 3 tet4, 4 hex20), linear != 0 = KINEM linear
extern "C" int emu_generic(int celltype, int matkind, const double* params, int linear, int stiff, int64_t n_ele,
    const int32_t* conn, const double* X, const double* u_col, const int32_t* pos, int pos_stride,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow, int64_t n_col_nodes, double* vals,
    double* f, const double* gp_hist, int fbar, int ans, double* eas_hist)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k], node_row[k].frow = node_frow[k];
  const CellTables T = make_cell_tables(celltype);
  GenericArgs Ga;
  EvalArgs& A = Ga.A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos, A.node_row = node_row.data();
  A.per_ele_param = nullptr, A.ele_id = nullptr, A.gp_hist = gp_hist, A.fibers = nullptr, A.linear = linear;
  A.mat_hist = nullptr, A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0;
  if (!make_mat_params(matkind, params, A.mp)) return -3;
  Ga.dN_gp = T.dN_gp.data(), Ga.dN_node = T.dN_node.data(), Ga.w = T.w.data();
  Ga.nn = T.nn, Ga.ngp = T.ngp, Ga.pos_stride = pos_stride, Ga.stiff = stiff;
  Ga.dN_cen = fbar ? T.dN_cen.data() : nullptr, Ga.hist_pts = T.ngp + (fbar ? 1 : 0);
  Ga.dN_ans = (ans && !T.dN_ans.empty()) ? T.dN_ans.data() : nullptr, Ga.xi_gp = T.xi_gp.data();
  if (ans && !Ga.dN_ans) return -4;
  // TECH shell_eas_ans: EAS history (n_ele x ANS_EAS_HIST, updated in place) and the centroid table of
  // hex8eas::centroid_transformation, here through the host twin the EAS emulation uses
  std::vector<double> t0;
  Ga.eas_hist = nullptr, Ga.eas_t0 = nullptr;
  if (eas_hist)
  {
    if (!ans || T.nn != 8) return -5;
    t0.resize((size_t)n_ele * ANS_EAS_T0_STRIDE);
    for (int64_t e = 0; e < n_ele; ++e) emu_centroid_t0(X, conn + 8 * e, t0.data() + e * ANS_EAS_T0_STRIDE);
    Ga.eas_hist = eas_hist, Ga.eas_t0 = t0.data();
  }
  if (matkind == MAT_COUP_NEOHOOKE) return run_generic_nn<MAT_COUP_NEOHOOKE>(Ga);
  if (matkind == MAT_ISO_VOL) return run_generic_nn<MAT_ISO_VOL>(Ga);
  return run_generic_nn<MAT_SVK>(Ga);
}

template <int NN>
static int run_generic_dp(const GenericArgs& Ga)
{
  const EvalArgs& A = Ga.A;
  int bad = 0;
  for (int64_t e = 0; e < A.n_ele; ++e)
    for (int i = 0; i < Ga.nn; ++i)
    {
      double Krow[NN][9], fi[3], dj = 1e300, df = 1e300;
      if (generic_block_row<MAT_DP, NN>(Ga, e, i, Krow, fi, dj, df)) ++bad;
      const NodeRow nr = A.node_row[A.conn[e * Ga.nn + i]];
      if (nr.rowstart < 0) continue;
      for (int b = 0; b < 3; ++b) A.f[nr.frow + b] += fi[b];
      if (!Ga.stiff) continue;
      const int32_t* pos = A.pos + e * Ga.pos_stride + i * Ga.nn;
      for (int j = 0; j < Ga.nn; ++j)
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c) A.vals[nr.rowstart + (int64_t)b * nr.rowlen + pos[j] + c] += Krow[j][b + 3 * c];
    }
  return bad;
}
extern "C" int emu_generic_drucker_prager(int celltype, const double* in10, int stiff, int64_t n_ele, const int32_t* conn,
    const double* X, const double* u_col, const int32_t* pos, int pos_stride, const int64_t* node_rowstart,
    const int32_t* node_rowlen, const int32_t* node_frow, int64_t n_col_nodes, double* vals, double* f, double* mat_hist)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k], node_row[k].frow = node_frow[k];
  double q[DP_NPAR];
  if (!make_drucker_prager(in10, q)) return -3;
  const CellTables T = make_cell_tables(celltype);
  GenericArgs Ga;
  EvalArgs& A = Ga.A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos, A.node_row = node_row.data();
  A.per_ele_param = nullptr, A.ele_id = nullptr, A.gp_hist = nullptr, A.fibers = nullptr, A.linear = 1;
  A.mat_hist = mat_hist, A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0;
  for (double& v : A.mp.p) v = 0.0;
  A.mp.hyper = nullptr, A.mp.fiber = nullptr, A.mp.ext = q, A.mp.hist = nullptr;
  Ga.dN_gp = T.dN_gp.data(), Ga.dN_node = T.dN_node.data(), Ga.w = T.w.data();
  Ga.nn = T.nn, Ga.ngp = T.ngp, Ga.pos_stride = pos_stride, Ga.stiff = stiff;
  Ga.dN_cen = nullptr, Ga.hist_pts = T.ngp, Ga.dN_ans = nullptr, Ga.xi_gp = T.xi_gp.data();
  return T.nn == 8 ? run_generic_dp<8>(Ga) : (T.nn == 10 ? run_generic_dp<10>(Ga) : run_generic_dp<27>(Ga));
}

extern "C" int emu_post(int celltype, int fbar, int matkind, const double* params, const double* per_ele,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const uint8_t* ele_is_row,
    int stress_type, int strain_type, double* stress, double* strain, double* energy, const double* gp_hist)
{
  const CellTables T = make_cell_tables(celltype);
  MatParams mp;
  HyperMat hm;
  if (!setup_material(matkind, params, mp, hm)) return -3;
  PostArgs A;
  A.X = X, A.u = u_col, A.conn = conn, A.per_ele_param = per_ele, A.ele_id = nullptr, A.fibers = nullptr;
  A.ele_is_row = ele_is_row, A.dN_gp = T.dN_gp.data(), A.dN_cen = T.dN_cen.data(), A.w = T.w.data();
  A.mp = mp, A.n_ele = n_ele, A.nn = T.nn, A.ngp = T.ngp, A.fbar = fbar == 1, A.linear = fbar == 2;
  A.gp_hist = gp_hist, A.hist_pts = T.ngp + ((gp_hist && fbar) ? 1 : 0);
  A.mat_hist = nullptr;
  A.eas_neas = 0, A.eas_hist_size = 0, A.eas_t0_stride = 0, A.eas_hist = nullptr, A.eas_t0 = nullptr;
  A.xi_gp = T.xi_gp.data();
  A.dN_ans = nullptr;
  if (fbar == 3)  // TECH shell_ans (hex8)
  {
    if (T.dN_ans.empty()) return -4;
    A.dN_ans = T.dN_ans.data();
  }
  A.stress_type = stress_type, A.strain_type = strain_type;
  A.stress = stress, A.strain = strain, A.energy = nullptr;
  switch (matkind)
  {
    case 0:
      run_post<MAT_COUP_NEOHOOKE>(A, energy);
      break;
    case 1:
      run_post<MAT_ISO_VOL>(A, energy);
      break;
    case 2:
      run_post<MAT_SVK>(A, energy);
      break;
    default:
      run_post<MAT_HYPER>(A, energy);
  }
  return 0;
}

extern "C" int emu_mass(int celltype, double density, int lumped, int64_t n_ele, const int32_t* conn,
    const double* X, const double* acc_col, const int32_t* pos, const int64_t* node_rowstart,
    const int32_t* node_rowlen, const int32_t* node_frow, int64_t n_col_nodes, double* mass_vals,
    double* f_inertia)
{
  const CellTables T = make_cell_tables(celltype);
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k],
    node_row[k].frow = node_frow[k];
  MassArgs M;
  M.X = X, M.conn = conn, M.pos = pos, M.node_row = node_row.data(), M.acc = acc_col;
  M.N_mass = T.N_mass.data(), M.dN_mass = T.dN_mass.data(), M.w_mass = T.w_mass.data();
  M.dN_gp = T.dN_gp.data(), M.w = T.w.data(), M.density = density, M.n_ele = n_ele;
  M.nn = T.nn, M.ngp = T.ngp, M.ngp_mass = T.ngp_mass, M.pos_stride = T.nn * T.nn;
  M.lumped = lumped, M.mass_vals = mass_vals, M.f_inertia = f_inertia;
  if (celltype == 0)
    run_mass<8>(M);
  else if (celltype == 1)
    run_mass<27>(M);
  else
    run_mass<10>(M);
  return 0;
}

// MULF history: fresh set-up (what k_mulf_init does) and struct_update_prestress (k_mulf_update)
extern "C" int emu_mulf_history(int celltype, int update, int64_t n_ele, const int32_t* conn, const double* X,
    const double* u_col, double* hist, int with_centroid)
{
  const CellTables T = make_cell_tables(celltype);
  const int npt = T.ngp + (with_centroid ? 1 : 0);  // MULF + F-bar: the centroid history follows the Gauss points'
  for (int64_t e = 0; e < n_ele; ++e)
    for (int g = 0; g < npt; ++g)
    {
      double* h = hist + (e * npt + g) * MULF_HIST;
      const double* dN = g < T.ngp ? T.dN_gp.data() + (size_t)g * T.nn * 3 : T.dN_cen.data();
      const int32_t* ce = conn + e * T.nn;
      if (update)
        mulf_update_point(T.nn, dN, X, u_col, ce, h);
      else
      {
        double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ[9];
        for (int k = 0; k < T.nn; ++k)
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * X[3 * (int64_t)ce[k] + b];
        inv3(iJ, J);
        for (int i = 0; i < 9; ++i) h[i] = iJ[i], h[9 + i] = (i % 4 == 0) ? 1.0 : 0.0;
      }
    }
  return 0;
}

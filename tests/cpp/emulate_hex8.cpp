// TEST INFRASTRUCTURE: runs the PRODUCT's hex8 device functions (4c_b200/csrc/hex8_kernel.cuh,
// solid_math.cuh, shape_tables.hpp) on the CPU by stepping the 32 lanes of a warp through the
// kernel's phases sequentially.  Lets the CPU test-suite check the kernel arithmetic, the shared
// memory indexing and the scatter addressing against the oracle without a GPU.  Not shipped, not
// a fallback: the library itself has no host execution path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../4c_b200/csrc/hex8_kernel.cuh"
#include "../../4c_b200/csrc/hex8_fbar_kernel.cuh"
#include "../../4c_b200/csrc/shape_tables.hpp"

using namespace b200s;

template <int MAT, bool STIFF, int TECH>
static void run(const EvalArgs& A, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  constexpr bool FBAR = TECH == 1, MULF = TECH == 2, LINEAR = TECH == 3;  // as in k_hex8_eval_assemble
  constexpr int MP_SIZE = FBAR ? hex8fbar::MP_SIZE : hex8::MP_SIZE;
  const CellTables T = make_cell_tables(0);
  std::vector<double> tab_gp(hex8::TAB_SIZE, 0.0), tab_node(hex8::TAB_SIZE, 0.0);
  for (int t = 0; t < 8 * 24; ++t)
  {
    tab_gp[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = T.dN_gp[t];
    tab_node[(t / 24) * hex8::TAB_STRIDE + (t % 24)] = T.dN_node[t];
  }
  std::vector<double> sm((FBAR ? hex8fbar::WARP_DOUBLES : hex8::WARP_DOUBLES) + 2, 0.0);
  std::vector<NodeRow> nrow(hex8::EPW * hex8::NN);
  std::vector<int32_t> posbuf(hex8::EPW * hex8::NN * hex8::NN);
  hex8::WarpSmem S;
  // 16-byte align the double2 region; same carve-up as the kernel
  S.mp = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm.data()) + 15) & ~uintptr_t(15));
  S.g = S.mp + MP_SIZE;
  S.xu = S.g + hex8::G_SIZE;
  if (FBAR) S.h0 = S.xu + hex8::XU_SIZE;
  S.nrow = nrow.data(), S.pos = posbuf.data();
  *min_detJ = 1e300, *min_detF = 1e300, *n_bad = 0;
  const int64_t n_tiles = (A.n_ele + hex8::EPW - 1) / hex8::EPW;
  static double out54[32][54];
  static double Kacc[32][FBAR ? 8 : 5][9], facc[32][3];
  for (int64_t tile = 0; tile < n_tiles; ++tile)
  {
    const int64_t e0 = tile * hex8::EPW;
    for (int lane = 0; lane < 32; ++lane) hex8::gather(A, S, lane, e0);
    bool negel[4] = {false, false, false, false};
    for (int lane = 0; lane < 32; ++lane)
    {
      double dJ, dF;
      bool neg;
      if (FBAR)
        hex8fbar::phase1<MAT, STIFF>(
            A, S, tab_gp.data(), T.w.data(), lane, e0, out54[lane], dJ, dF, neg);
      else
        hex8::phase1<MAT, STIFF, MULF, LINEAR>(A, S, tab_gp.data(), tab_node.data(), T.w.data(), lane, e0,
            out54[lane], dJ, dF, neg);
      if (e0 + (lane >> 3) < A.n_ele)
      {
        if (dJ < *min_detJ) *min_detJ = dJ;
        if (dF < *min_detF) *min_detF = dF;
      }
      if (neg) negel[lane >> 3] = true;
    }
    for (bool b : negel) *n_bad += b;
    for (int lane = 0; lane < 32; ++lane) hex8::phase1_store<STIFF>(S, lane, out54[lane]);
    if constexpr (FBAR)
    {
      for (int lane = 0; lane < 32; ++lane) hex8fbar::phase2<STIFF>(S, lane, Kacc[lane], facc[lane]);
      for (int lane = 0; lane < 32; ++lane) hex8fbar::stage_ke<STIFF>(S, lane, Kacc[lane], facc[lane]);
    }
    else
    {
      for (int lane = 0; lane < 32; ++lane) hex8::phase2<STIFF>(S, lane, Kacc[lane], facc[lane]);
      for (int lane = 0; lane < 32; ++lane) hex8::stage_ke<STIFF>(S, lane, Kacc[lane], facc[lane]);
    }
    for (int lane = 0; lane < 32; ++lane) hex8::scatter<STIFF>(A, S, lane, e0);
  }
}

static int emu_hex8_tech(int tech, int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad,
    const double* gp_hist = nullptr)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k],
    node_row[k].frow = node_frow[k];
  EvalArgs A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos, A.node_row = node_row.data();
  A.per_ele_param = per_ele, A.ele_id = nullptr, A.gp_hist = nullptr, A.fibers = nullptr, A.linear = 0, A.mat_hist = nullptr;
  A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0;
  if (!make_mat_params(matkind, params, A.mp)) return -3;
  A.gp_hist = gp_hist;
  if (tech == 3) A.linear = 1;
#define RUN(M)                                                  \
  if (tech == 3 && stiff)                                       \
    run<M, true, 3>(A, min_detJ, min_detF, n_bad);              \
  else if (tech == 3)                                           \
    run<M, false, 3>(A, min_detJ, min_detF, n_bad);             \
  else if (tech == 2 && stiff)                                  \
    run<M, true, 2>(A, min_detJ, min_detF, n_bad);              \
  else if (tech == 2)                                           \
    run<M, false, 2>(A, min_detJ, min_detF, n_bad);             \
  else if (tech && stiff)                                       \
    run<M, true, 1>(A, min_detJ, min_detF, n_bad);              \
  else if (tech)                                                \
    run<M, false, 1>(A, min_detJ, min_detF, n_bad);             \
  else if (stiff)                                               \
    run<M, true, 0>(A, min_detJ, min_detF, n_bad);              \
  else                                                          \
    run<M, false, 0>(A, min_detJ, min_detF, n_bad)
  if (matkind == 0)
  {
    RUN(MAT_COUP_NEOHOOKE);
  }
  else if (matkind == 1)
  {
    RUN(MAT_ISO_VOL);
  }
  else
  {
    RUN(MAT_SVK);
  }
  return 0;
}

extern "C" int emu_hex8(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  return emu_hex8_tech(0, matkind, params, per_ele, stiff, n_ele, conn, X, u_col, pos, node_rowstart,
      node_rowlen, node_frow, n_col_nodes, vals, f, min_detJ, min_detF, n_bad);
}
// F-bar variant (hex8_fbar_kernel.cuh)
extern "C" int emu_hex8fbar(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  return emu_hex8_tech(1, matkind, params, per_ele, stiff, n_ele, conn, X, u_col, pos, node_rowstart,
      node_rowlen, node_frow, n_col_nodes, vals, f, min_detJ, min_detF, n_bad);
}

// MULF variant (TECH 2 of k_hex8_eval_assemble): gp_hist = n_ele * 8 * 18 doubles
extern "C" int emu_hex8mulf(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad,
    const double* gp_hist)
{
  return emu_hex8_tech(2, matkind, params, per_ele, stiff, n_ele, conn, X, u_col, pos, node_rowstart,
      node_rowlen, node_frow, n_col_nodes, vals, f, min_detJ, min_detF, n_bad, gp_hist);
}

// KINEM linear (EvalArgs::linear)
extern "C" int emu_hex8linear(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  return emu_hex8_tech(3, matkind, params, per_ele, stiff, n_ele, conn, X, u_col, pos, node_rowstart,
      node_rowlen, node_frow, n_col_nodes, vals, f, min_detJ, min_detF, n_bad);
}

// KINEM linear hex8 block with the history material MAT_Struct_DruckerPrager: mat_hist = n_ele * 8 * DP_HIST doubles
// (updated in place like the kernel does); in10 = input-file parameters
extern "C" int emu_hex8_drucker_prager(const double* in10, int stiff, int64_t n_ele, const int32_t* conn,
    const double* X, const double* u_col, const int32_t* pos, const int64_t* node_rowstart, const int32_t* node_rowlen,
    const int32_t* node_frow, int64_t n_col_nodes, double* vals, double* f, double* mat_hist)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k], node_row[k].frow = node_frow[k];
  double q[DP_NPAR];
  if (!make_drucker_prager(in10, q)) return -3;
  EvalArgs A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos, A.node_row = node_row.data();
  A.per_ele_param = nullptr, A.ele_id = nullptr, A.gp_hist = nullptr, A.fibers = nullptr, A.linear = 1;
  A.mat_hist = mat_hist;
  A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0;
  for (double& v : A.mp.p) v = 0.0;
  A.mp.hyper = nullptr, A.mp.fiber = nullptr, A.mp.ext = q, A.mp.hist = nullptr;
  double dj = 1e300, df = 1e300;
  int64_t bad = 0;
  if (stiff)
    run<MAT_DP, true, 3>(A, &dj, &df, &bad);
  else
    run<MAT_DP, false, 3>(A, &dj, &df, &bad);
  return (int)bad;
}

// nonlinear hex8 block with the finite-strain history material MAT_VISCONEOHOOKE: mat_hist = n_ele * 8 * VISCO_HIST
// doubles (updated in place like the kernel does); in6 = input-file parameters, dt = time step size
extern "C" int emu_hex8_visco(const double* in6, double dt, int stiff, int64_t n_ele, const int32_t* conn,
    const double* X, const double* u_col, const int32_t* pos, const int64_t* node_rowstart, const int32_t* node_rowlen,
    const int32_t* node_frow, int64_t n_col_nodes, double* vals, double* f, double* mat_hist)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k], node_row[k].frow = node_frow[k];
  EvalArgs A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos, A.node_row = node_row.data();
  A.per_ele_param = nullptr, A.ele_id = nullptr, A.gp_hist = nullptr, A.fibers = nullptr, A.linear = 0;
  A.mat_hist = mat_hist;
  A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0;
  for (double& v : A.mp.p) v = 0.0;
  if (!make_visco_neohooke(in6, dt, A.mp.p)) return -3;
  A.mp.hyper = nullptr, A.mp.fiber = nullptr, A.mp.ext = nullptr, A.mp.hist = nullptr;
  double dj = 1e300, df = 1e300;
  int64_t bad = 0;
  if (stiff)
    run<MAT_VISCO, true, 0>(A, &dj, &df, &bad);
  else
    run<MAT_VISCO, false, 0>(A, &dj, &df, &bad);
  return (int)bad;
}

// TEST INFRASTRUCTURE: the PRODUCT's tet10 device functions (4c_b200/csrc/tet10_kernel.cuh) stepped
// lane by lane on the CPU, like emulate_hex8.cpp.  Not shipped, not a fallback.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../4c_b200/csrc/shape_tables.hpp"
#include "../../4c_b200/csrc/tet10_kernel.cuh"

using namespace b200s;

template <int MAT, bool STIFF>
static void run(const EvalArgs& A, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  const CellTables T = make_cell_tables(2);
  std::vector<double> tab_gp(tet10::TAB_GP_SIZE, 0.0), tab_node(tet10::TAB_NODE_SIZE, 0.0);
  for (int t = 0; t < 4 * 30; ++t) tab_gp[(t / 30) * tet10::TAB_STRIDE + (t % 30)] = T.dN_gp[t];
  for (int t = 0; t < 10 * 30; ++t) tab_node[(t / 30) * tet10::TAB_STRIDE + (t % 30)] = T.dN_node[t];
  std::vector<double> sm(tet10::WARP_DOUBLES + 2, 0.0);
  std::vector<NodeRow> nrow(32);
  std::vector<int32_t> posbuf(tet10::POS_INTS + 4);
  tet10::WarpSmem S;
  S.mp = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm.data()) + 15) & ~uintptr_t(15));
  S.g = S.mp + tet10::MP_SIZE;
  S.xu = S.mp + tet10::STAGE_SIZE;
  S.nrow = nrow.data(), S.pos = posbuf.data();
  *min_detJ = 1e300, *min_detF = 1e300, *n_bad = 0;
  const int64_t n_tiles = (A.n_ele + tet10::EPW - 1) / tet10::EPW;
  static double out54[32][54];
  static double Kacc[32][tet10::NBLK][9], facc[32][3];
  for (int64_t tile = 0; tile < n_tiles; ++tile)
  {
    const int64_t e0 = tile * tet10::EPW;
    for (int lane = 0; lane < 32; ++lane) tet10::gather(A, S, lane, e0);
    bool negel[3] = {false, false, false};
    for (int lane = 0; lane < 32; ++lane)
    {
      const int el = lane / tet10::NN;
      const bool live = el < tet10::EPW && e0 + el < A.n_ele;
      const double dJn = tet10::node_jacobian_det(S, tab_node.data(), lane);
      double dJ, dF;
      tet10::phase1<MAT, STIFF>(A, S, tab_gp.data(), T.w.data(), lane, e0, out54[lane], dJ, dF);
      if (live)
      {
        if (dJn < *min_detJ) *min_detJ = dJn;
        if (dJ < *min_detJ) *min_detJ = dJ;
        if (dF < *min_detF) *min_detF = dF;
        if (!(dJn > 0.0)) negel[el] = true;
      }
    }
    for (bool b : negel) *n_bad += b;
    for (int lane = 0; lane < 32; ++lane) tet10::phase1_store<STIFF>(S, lane, out54[lane]);
    for (int lane = 0; lane < 32; ++lane) tet10::phase2<STIFF>(S, lane, Kacc[lane], facc[lane]);
    for (int lane = 0; lane < 32; ++lane) tet10::stage_ke<STIFF>(S, lane, Kacc[lane], facc[lane]);
    for (int lane = 0; lane < 32; ++lane) tet10::scatter<STIFF>(A, S, lane, e0);
  }
}

extern "C" int emu_tet10(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad)
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
#define RUN(M)                                                  \
  if (stiff)                                                    \
    run<M, true>(A, min_detJ, min_detF, n_bad);                 \
  else                                                          \
    run<M, false>(A, min_detJ, min_detF, n_bad)
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

// TEST INFRASTRUCTURE: the PRODUCT's hex27 device functions (4c_b200/csrc/hex27_kernel.cuh) stepped
// lane by lane (both warps of an element pair) on the CPU.  Not shipped, not a fallback.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../4c_b200/csrc/hex27_kernel.cuh"
#include "../../4c_b200/csrc/shape_tables.hpp"

using namespace b200s;

template <int MAT, bool STIFF>
static void run(const EvalArgs& A, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  const CellTables T = make_cell_tables(1);
  hex27::IntTables it;
  std::memset(&it, 0, sizeof it);
  const double(*hc)[3] = detail::hex_nodes();
  for (int k = 0; k < 27; ++k)
  {
    const int a = (int)hc[k][0] + 1, b = (int)hc[k][1] + 1, c = (int)hc[k][2] + 1;
    it.abc[k] = a | (b << 2) | (c << 4);
    it.node_of[a + 3 * b + 9 * c] = k;
  }
  std::vector<unsigned char> raw(hex27::PAIR_BYTES + 32);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw.data()) + 15) & ~uintptr_t(15));
  const hex27::PairSmem S = hex27::carve(base);
  *min_detJ = 1e300, *min_detF = 1e300, *n_bad = 0;
  static double Kacc[2][32][hex27::NBLK][9], facc[2][32][3];
  for (int64_t e = 0; e < A.n_ele; ++e)
  {
    for (int q = 0; q < 27; ++q) hex27::gather(A, S, it, q, e);
    bool neg = false;
    for (int g = 0; g < 27; ++g)
    {
      double out54[54], dJ, dF;
      hex27::phase1<MAT, STIFF>(A, S, T.dN_gp.data(), T.w.data(), g, e, out54, dJ, dF);
      if (dJ < *min_detJ) *min_detJ = dJ;
      if (dF < *min_detF) *min_detF = dF;
    }
    for (int n = 0; n < 27; ++n)
    {
      const double dJn = hex27::node_jacobian_det(S, it, n);
      if (dJn < *min_detJ) *min_detJ = dJn;
      if (!(dJn > 0.0)) neg = true;
    }
    *n_bad += neg;
    for (int w = 0; w < 2; ++w)
      for (int lane = 0; lane < 27; ++lane) hex27::phase2<STIFF>(S, w, lane, Kacc[w][lane], facc[w][lane]);
    for (int w = 0; w < 2; ++w)
    {
      double* stage = S.stage + w * hex27::STAGE_SIZE;
      if (STIFF)
        for (int s = 0; s < 2; ++s)
        {
          for (int lane = 0; lane < 27; ++lane) hex27::stage_blocks(stage, lane, s, Kacc[w][lane]);
          for (int lane = 0; lane < 32; ++lane) hex27::scatter_blocks(A, S, stage, w, s, lane);
        }
      if (w == 0)
        for (int lane = 0; lane < 27; ++lane) hex27::scatter_force(A, S, lane, facc[0][lane]);
    }
  }
}

extern "C" int emu_hex27(int matkind, const double* params, const double* per_ele, int stiff,
    int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, const int32_t* pos,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow,
    int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF, int64_t* n_bad)
{
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k],
    node_row[k].frow = node_frow[k];
  // the python helper passes the map with stride nn*nn; the kernel expects POS_STRIDE
  std::vector<int32_t> pos2((size_t)n_ele * hex27::POS_STRIDE, 0);
  for (int64_t e = 0; e < n_ele; ++e)
    std::memcpy(&pos2[e * hex27::POS_STRIDE], pos + e * 729, 729 * sizeof(int32_t));
  EvalArgs A;
  A.X = X, A.u = u_col, A.conn = conn, A.pos = pos2.data(), A.node_row = node_row.data();
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

// hex20 / hex18 blocks through the hex27 kernel (template flag PAD of k_hex27_eval_assemble): the kernel's steps with the
// product's padded tables (pad_tables_to_hex27), the padded connectivity / scatter map the caller built, the padding rows
// marked not owned as issue_rowpos<true> does, no nodal check (k_nodal_jacobian_check runs ahead of the kernel)
template <int MAT, bool STIFF>
static void run_padded(const EvalArgs& A, const double* dN, const double* w, double* min_detJ, double* min_detF)
{
  hex27::IntTables it;
  std::memset(&it, 0, sizeof it);
  std::vector<unsigned char> raw(hex27::PAIR_BYTES + 32);
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw.data()) + 15) & ~uintptr_t(15));
  const hex27::PairSmem S = hex27::carve(base);
  *min_detJ = 1e300, *min_detF = 1e300;
  static double Kacc[2][32][hex27::NBLK][9], facc[2][32][3];
  for (int64_t e = 0; e < A.n_ele; ++e)
  {
    for (int q = 0; q < 27; ++q)
    {
      hex27::gather(A, S, it, q, e);
      if (q >= A.nn_real) S.nrow[q] = NodeRow{-1, 0, -1};
    }
    for (int g = 0; g < 27; ++g)
    {
      double out54[54], dJ, dF;
      hex27::phase1<MAT, STIFF>(A, S, dN, w, g, e, out54, dJ, dF);
      if (dJ < *min_detJ) *min_detJ = dJ;
      if (dF < *min_detF) *min_detF = dF;
    }
    for (int wp = 0; wp < 2; ++wp)
      for (int lane = 0; lane < 27; ++lane) hex27::phase2<STIFF>(S, wp, lane, Kacc[wp][lane], facc[wp][lane]);
    for (int wp = 0; wp < 2; ++wp)
    {
      double* stage = S.stage + wp * hex27::STAGE_SIZE;
      if (STIFF)
        for (int s = 0; s < 2; ++s)
        {
          for (int lane = 0; lane < 27; ++lane) hex27::stage_blocks(stage, lane, s, Kacc[wp][lane]);
          for (int lane = 0; lane < 32; ++lane) hex27::scatter_blocks<true>(A, S, stage, wp, s, lane);
        }
      if (wp == 0)
        for (int lane = 0; lane < 27; ++lane) hex27::scatter_force(A, S, lane, facc[0][lane]);
    }
  }
}

// celltype as in shape_tables.hpp (4 hex20, 7 hex18); conn27 / pos27: [n_ele][27], [n_ele][729] with the padding of
// k_pad_conn / k_build_scatter_map(nn_real)
extern "C" int emu_hex27_padded(int celltype, int matkind, const double* params, int stiff, int64_t n_ele, const int32_t* conn27,
    const double* X, const double* u_col, const int32_t* pos27, const int64_t* node_rowstart, const int32_t* node_rowlen,
    const int32_t* node_frow, int64_t n_col_nodes, double* vals, double* f, double* min_detJ, double* min_detF)
{
  const CellTables T = make_cell_tables(celltype);
  if (T.nn > 27 || T.ngp > 27) return -4;
  std::vector<double> dN(27 * 81), w(28, 0.0);
  pad_tables_to_hex27(T, dN.data(), w.data());
  std::vector<NodeRow> node_row(n_col_nodes);
  for (int64_t k = 0; k < n_col_nodes; ++k)
    node_row[k].rowstart = node_rowstart[k], node_row[k].rowlen = node_rowlen[k], node_row[k].frow = node_frow[k];
  std::vector<int32_t> pos2((size_t)n_ele * hex27::POS_STRIDE, 0);
  for (int64_t e = 0; e < n_ele; ++e) std::memcpy(&pos2[e * hex27::POS_STRIDE], pos27 + e * 729, 729 * sizeof(int32_t));
  EvalArgs A;
  A.X = X, A.u = u_col, A.conn = conn27, A.pos = pos2.data(), A.node_row = node_row.data();
  A.per_ele_param = nullptr, A.ele_id = nullptr, A.gp_hist = nullptr, A.fibers = nullptr, A.linear = 0, A.mat_hist = nullptr;
  A.vals = vals, A.f = f, A.err = nullptr, A.n_ele = n_ele, A.scatter_mode = 0, A.nn_real = T.nn;
  if (!make_mat_params(matkind, params, A.mp)) return -3;
#define RUNP(M)                                                      \
  if (stiff)                                                         \
    run_padded<M, true>(A, dN.data(), w.data(), min_detJ, min_detF); \
  else                                                               \
    run_padded<M, false>(A, dN.data(), w.data(), min_detJ, min_detF)
  if (matkind == 0)
  {
    RUNP(MAT_COUP_NEOHOOKE);
  }
  else if (matkind == 1)
  {
    RUNP(MAT_ISO_VOL);
  }
  else
  {
    RUNP(MAT_SVK);
  }
  return 0;
}

// TEST INFRASTRUCTURE: runs the PRODUCT's __host__ __device__ load functions of
// 4c_b200/csrc/neumann_kernels.cuh on the CPU, one call per thread the kernels would launch
// (k_neumann_surface_force, k_neumann_surface_lin, k_neumann_volume_force, the *_gp_coords kernels),
// with the same owner-computes addressing.  Not shipped, not a fallback.
#include <cstdint>
#include <cmath>
#include <vector>

#include "../../4c_b200/csrc/neumann_kernels.cuh"
#include "../../4c_b200/csrc/shape_tables.hpp"

using namespace b200s;

namespace
{
  std::vector<NodeRow> node_rows(int64_t n, const int64_t* rowstart, const int32_t* rowlen, const int32_t* frow)
  {
    std::vector<NodeRow> nr(n);
    for (int64_t k = 0; k < n; ++k) nr[k].rowstart = rowstart[k], nr[k].rowlen = rowlen[k], nr[k].frow = frow[k];
    return nr;
  }
}  // namespace

// returns 0, or 1 if a face couples nodes without a graph entry (the kernel's flag)
extern "C" int emu_neumann_surface(int celltype, int kind, int64_t n_faces, const int32_t* faces, const double* X,
    const double* u_col, const int32_t* onoff, const double* val, const double* gp_fac, int64_t n_col_nodes,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow, const int32_t* node_colbase,
    const int32_t* colind, double* fext, double* K, double* gp_xyz)
{
  const FaceTables FT = make_face_tables(celltype);
  const std::vector<NodeRow> nr = node_rows(n_col_nodes, node_rowstart, node_rowlen, node_frow);
  NeumannSurfArgs A;
  A.T.nn = FT.nn, A.T.ngp = FT.ngp, A.T.N = FT.N.data(), A.T.d0 = FT.d0.data(), A.T.d1 = FT.d1.data();
  A.T.w = FT.w.data();
  A.X = X, A.u = u_col, A.faces = faces, A.n_faces = n_faces, A.kind = kind;
  for (int d = 0; d < 3; ++d) A.onoff[d] = onoff[d], A.val[d] = val[d];
  A.gp_fac = gp_fac, A.node_row = nr.data(), A.node_colbase = node_colbase, A.colind = colind;
  A.fext = fext, A.K = K;
  int flag = 0;
  double x[3 * FACE_MAX_NODES];
  for (int64_t f = 0; f < n_faces; ++f)
  {
    face_coordinates(A, f, x);
    for (int a = 0; a < FT.nn; ++a)
    {
      const NodeRow& r = nr[faces[f * FT.nn + a]];
      if (r.frow < 0) continue;
      double out[3];
      neumann_face_force(A, f, x, a, out);
      for (int d = 0; d < 3; ++d) fext[r.frow + d] += out[d];
      if (!K) continue;
      for (int b = 0; b < FT.nn; ++b)
      {
        const int lo = neumann_find_slot(r, colind, node_colbase[faces[f * FT.nn + b]]);
        if (lo < 0)
        {
          flag = 1;
          continue;
        }
        double blk[9];
        neumann_face_lin_block(A, f, x, a, b, blk);
        for (int dim = 0; dim < 3; ++dim)
          for (int k = 0; k < 3; ++k) K[r.rowstart + (int64_t)dim * r.rowlen + lo + k] += blk[dim + 3 * k];
      }
    }
    if (gp_xyz)
      for (int g = 0; g < FT.ngp; ++g) neumann_face_gp_position(A, x, g, gp_xyz + (f * FT.ngp + g) * 3);
  }
  return flag;
}

extern "C" int emu_neumann_volume(int celltype, int64_t n_ele, const int32_t* conn, const double* X,
    const int32_t* ele_slot, const int32_t* onoff, const double* val, const double* gp_fac, int64_t n_col_nodes,
    const int64_t* node_rowstart, const int32_t* node_rowlen, const int32_t* node_frow, double* fext)
{
  const CellTables T = make_cell_tables(celltype);
  const std::vector<NodeRow> nr = node_rows(n_col_nodes, node_rowstart, node_rowlen, node_frow);
  NeumannVolArgs A;
  A.X = X, A.conn = conn, A.ele_id = nullptr, A.ele_slot = ele_slot;
  A.N_gp = T.N_gp.data(), A.dN_gp = T.dN_gp.data(), A.w = T.w.data();
  A.n_ele = n_ele, A.nn = T.nn, A.ngp = T.ngp;
  for (int d = 0; d < 3; ++d) A.onoff[d] = onoff[d], A.val[d] = val[d];
  A.gp_fac = gp_fac, A.node_row = nr.data(), A.fext = fext;
  for (int64_t e = 0; e < n_ele; ++e)
    for (int a = 0; a < T.nn; ++a)
    {
      const NodeRow& r = nr[conn[e * T.nn + a]];
      if (r.frow < 0) continue;
      double out[3];
      if (!neumann_volume_force(A, e, a, out)) continue;
      for (int d = 0; d < 3; ++d) fext[r.frow + d] += out[d];
    }
  return 0;
}

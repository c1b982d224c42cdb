// linalg_kernels.cuh -- the step after the hot path, on the device, so that the
// assembled stiffness never has to return to the host:
//   SparseMatrix::apply_dirichlet(dbctoggle, diagonalblock), in-place branch
//                                              (4C_linalg_sparsematrix.cpp:816-925, :893-921)
//   apply_dirichlet_to_system(x, b, dbcval, dbctoggle)
//                                              (4C_linalg_utils_sparse_algebra_assemble.cpp:116-130)
//   SparseMatrix::multiply (Epetra_CrsMatrix::Multiply: import of x to the column map + local CSR
//   product)                                   (4C_linalg_sparsematrix.cpp, "multiply")
//   and a block-Jacobi preconditioned CG on top of them as the hand-off demonstrator.
// All HBM-bound streaming kernels: coalesced row reads, x gathered through L2.
#pragma once

#include <stdint.h>

#include "hex8_kernel.cuh"

namespace b200s
{
#if defined(__CUDACC__)
  // matrix-local column id of every owned row's own dof (the diagonal entry): the column map need not list
  // the owned dofs first / in row order, so "column id == row lid" is not assumed
  __global__ void k_row_diag_col(int64_t n_col_nodes, const NodeRow* node_row, const int32_t* node_colbase,
      int32_t* row_diagcol)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_col_nodes) return;
    const int64_t k = t / 3;
    const int d = (int)(t - 3 * k);
    const int32_t fr = node_row[k].frow;
    if (fr >= 0) row_diagcol[fr + d] = node_colbase[k] + d;
  }

  // one warp per CSR row: rows flagged 1.0 are zeroed; diagonalblock puts 1.0 on the diagonal.  missing counts
  // toggled rows without a diagonal entry in the graph (the caller reports them instead of leaving a zero row).
  __global__ void k_apply_dirichlet(int64_t n_rows, const int64_t* rowptr, const int32_t* colind,
      const int32_t* row_diagcol, const double* toggle, int diagonalblock, double* vals, int* missing)
  {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows || toggle[row] != 1.0) return;
    const int64_t s = rowptr[row], t = rowptr[row + 1];
    const int32_t dc = row_diagcol[row];
    bool found = false;
    for (int64_t q = s + lane; q < t; q += 32)
    {
      const bool diag = colind[q] == dc;
      found |= diag;
      vals[q] = (diagonalblock && diag) ? 1.0 : 0.0;
    }
    if (diagonalblock && !__any_sync(0xffffffffu, found) && lane == 0) atomicAdd(missing, 1);
  }

  __global__ void k_dirichlet_to_system(int64_t n, const double* toggle, const double* dbcval,
      double* x, double* b)
  {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || toggle[i] != 1.0) return;
    if (x) x[i] = dbcval[i];
    if (b) b[i] = dbcval[i];
  }

  // owned part of a vector on the matrix column map: x_mat[colbase(node) + d] = x_row[frow(node) + d]
  __global__ void k_row_to_matcol(int64_t n_col_nodes, const NodeRow* node_row,
      const int32_t* node_colbase, const double* x_row, double* x_mat)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n_col_nodes) return;
    const int64_t k = t / 3;
    const int d = (int)(t - 3 * k);
    const int32_t fr = node_row[k].frow;
    if (fr >= 0) x_mat[node_colbase[k] + d] = x_row[fr + d];
  }
  // ghost part: received node values (3 per node, in recv-list order) to the matrix column ids
  __global__ void k_unpack_halo_matcol(int64_t n, const int32_t* recv_col_lid,
      const int32_t* node_colbase, const double* buf, double* x_mat)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * n) return;
    const int64_t k = t / 3;
    x_mat[node_colbase[recv_col_lid[k]] + (int)(t - 3 * k)] = buf[t];
  }

  // General CSR product, one warp per row (any graph).  12 bytes per non-zero.
  // dot != nullptr: atomically accumulates sum_r w[r] * y[r] (CG's p.Ap with w = p).
  __global__ void __launch_bounds__(256) k_spmv_csr(int64_t n_rows, const int64_t* rowptr,
      const int32_t* colind, const double* vals, const double* x_mat, double* y, const double* w,
      double* dot)
  {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    if (row < n_rows)
    {
      const int64_t s = rowptr[row], t = rowptr[row + 1];
      for (int64_t q = s + lane; q < t; q += 32) acc = fma(vals[q], __ldg(x_mat + colind[q]), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && row < n_rows)
    {
      y[row] = acc;
      if (dot) atomicAdd(dot, acc * w[row]);
    }
  }

  // Node-blocked product for the layout the assembly fast path requires (three rows of a node share
  // one pattern made of column triplets): a warp handles one row node, a lane one column node per
  // step -- one column index per 9 values, 8.4 bytes per non-zero instead of 12.
  // Row nodes are the column nodes with frow >= 0 (owned), visited in column-node order.
  __global__ void __launch_bounds__(256) k_spmv_node3(int64_t n_col_nodes, const NodeRow* node_row,
      const int32_t* colind /* block column ids, k_build_blkcol */, const double* vals, const double* x_mat,
      double* y, const double* w, double* dot)
  {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_col_nodes) return;
    const NodeRow nr = node_row[k];
    if (nr.frow < 0) return;  // warp-uniform
    const int nblk = nr.rowlen / 3;
    const double* v0 = vals + nr.rowstart;
    const double* v1 = v0 + nr.rowlen;
    const double* v2 = v1 + nr.rowlen;
    // one column id per 3x3 block: rows of a node are consecutive, so rowstart / 9 is the block offset
    const int32_t* ci = colind + nr.rowstart / 9;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int q = lane; q < nblk; q += 32)
    {
      const int32_t c = ci[q];
      const double x0 = __ldg(x_mat + c), x1 = __ldg(x_mat + c + 1), x2 = __ldg(x_mat + c + 2);
      a0 = fma(v0[3 * q + 2], x2, fma(v0[3 * q + 1], x1, fma(v0[3 * q], x0, a0)));
      a1 = fma(v1[3 * q + 2], x2, fma(v1[3 * q + 1], x1, fma(v1[3 * q], x0, a1)));
      a2 = fma(v2[3 * q + 2], x2, fma(v2[3 * q + 1], x1, fma(v2[3 * q], x0, a2)));
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0)
    {
      y[nr.frow] = a0, y[nr.frow + 1] = a1, y[nr.frow + 2] = a2;
      if (dot) atomicAdd(dot, a0 * w[nr.frow] + a1 * w[nr.frow + 1] + a2 * w[nr.frow + 2]);
    }
  }

  // compressed column ids of the node-blocked layout: first column of every 3x3 block, nnz / 9 entries
  __global__ void k_build_blkcol(int64_t n_col_nodes, const NodeRow* node_row, const int32_t* colind,
      int32_t* blkcol)
  {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_col_nodes) return;
    const NodeRow nr = node_row[k];
    if (nr.frow < 0) return;
    for (int q = lane; q < nr.rowlen / 3; q += 32) blkcol[nr.rowstart / 9 + q] = colind[nr.rowstart + 3 * q];
  }

  // checks the node-blocked pattern k_spmv_node3 relies on; flag |= 1 otherwise
  __global__ void k_check_node3(int64_t n_col_nodes, const NodeRow* node_row, const int32_t* colind,
      int* flag)
  {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= n_col_nodes) return;
    const NodeRow nr = node_row[k];
    if (nr.frow < 0) return;
    bool bad = (nr.rowlen % 3) != 0 || (nr.rowstart % 9) != 0;
    const int32_t* ci = colind + nr.rowstart;
    for (int q = lane; q < nr.rowlen && !bad; q += 32)
    {
      const int32_t c = ci[q];
      if (c != ci[q - q % 3] + q % 3) bad = true;
      if (ci[nr.rowlen + q] != c || ci[2 * (int64_t)nr.rowlen + q] != c) bad = true;
    }
    if (bad) atomicOr(flag, 1);
  }

  // inverse of the 3x3 diagonal block of every owned node (block-Jacobi preconditioner);
  // a singular block falls back to the identity
  __global__ void k_block_diag_inverse(int64_t n_col_nodes, const NodeRow* node_row, const int32_t* node_colbase,
      const int32_t* colind, const double* vals, double* dinv /*9 per row node, at 3*frow*/)
  {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_col_nodes) return;
    const NodeRow nr = node_row[k];
    if (nr.frow < 0) return;
    const int32_t* ci = colind + nr.rowstart;
    const int32_t dc = node_colbase[k];  // matrix-local column id of the node's first dof
    int lo = 0, hi = nr.rowlen;
    while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (ci[mid] < dc)
        lo = mid + 1;
      else
        hi = mid;
    }
    double m[9], o[9];
    bool ok = lo + 3 <= nr.rowlen && ci[lo] == dc;
    if (ok)
    {
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c) m[b + 3 * c] = vals[nr.rowstart + (int64_t)b * nr.rowlen + lo + c];
      const double det = inv3(o, m);
      ok = det != 0.0 && isfinite(det);
    }
    double* out = dinv + 3 * (int64_t)nr.frow;
    for (int i = 0; i < 9; ++i) out[i] = ok ? o[i] : ((i % 4 == 0) ? 1.0 : 0.0);
  }

  // s layout: [0] p.Ap  [1] r.z (old)  [2] r.z (new)  [3] r.r
  // x += alpha p; r -= alpha Ap; z = Dinv r; s[2] += r.z; s[3] += r.r      (alpha = s[1]/s[0])
  __global__ void __launch_bounds__(256) k_cg_update(int64_t n_row_nodes, const double* dinv,
      const double* p, const double* Ap, double* x, double* r, double* z, double* s, int first)
  {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double rz = 0.0, rr = 0.0;
    if (k < n_row_nodes)
    {
      const double alpha = (first || s[0] == 0.0) ? 0.0 : s[1] / s[0];  // p.Ap == 0: converged exactly
      double rv[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        const int64_t i = 3 * k + d;
        if (!first)
        {
          x[i] = fma(alpha, p[i], x[i]);
          rv[d] = fma(-alpha, Ap[i], r[i]);
          r[i] = rv[d];
        }
        else
          rv[d] = r[i];
      }
      const double* D = dinv + 9 * k;
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        const double zv = D[d] * rv[0] + D[d + 3] * rv[1] + D[d + 6] * rv[2];
        z[3 * k + d] = zv;
        rz += zv * rv[d];
        rr += rv[d] * rv[d];
      }
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      rz += __shfl_xor_sync(0xffffffffu, rz, o);
      rr += __shfl_xor_sync(0xffffffffu, rr, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
      atomicAdd(s + 2, rz);
      atomicAdd(s + 3, rr);
    }
  }
  // p = z + beta p (beta = s[2]/s[1], 0 on the first iteration)
  __global__ void k_cg_direction(int64_t n, const double* z, double* p, const double* s, int first)
  {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (first)
    {
      p[i] = z[i];  // p is uninitialised memory here: 0 * NaN would poison it
      return;
    }
    const double beta = (s[1] == 0.0) ? 0.0 : s[2] / s[1];
    p[i] = fma(beta, p[i], z[i]);
  }
  // after the direction update: s[1] = s[2]; s[0] = s[2] = s[3] = 0; history[it] = r.r
  __global__ void k_cg_roll(double* s, double* history, int it)
  {
    history[it] = s[3];
    s[1] = s[2];
    s[0] = 0.0, s[2] = 0.0, s[3] = 0.0;
  }
  // r = b - y
  __global__ void k_residual(int64_t n, const double* b, const double* y, double* r)
  {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = b[i] - y[i];
  }
#endif
}  // namespace b200s

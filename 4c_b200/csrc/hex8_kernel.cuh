// hex8_kernel.cuh -- evaluate + assemble for hex8 displacement-based nonlinear solid elements.
//
// Replaces, for a block of hex8 elements, the body of the element loop of
// Core::FE::Discretization::evaluate (4C_fem_discretization_evaluate.cpp:85-117):
//   Solid<3>::evaluate(struct_calc_nlnstiff | struct_calc_internalforce)
//     -> SolidEleCalc<hex8,DisplacementBased>::evaluate_nonlinear_force_stiffness_mass
//        (src/solid_ele/4C_solid_ele_calc.cpp:137-281)
//   AssembleStrategy::assemble_matrix1/vector1 -> SparseMatrix::assemble, LinAlg::assemble
//        (4C_linalg_sparsematrix.cpp:339-467, 4C_linalg_utils_sparse_algebra_assemble.cpp:60-82)
//
// Mapping: one warp = 4 elements x 8 threads.
//   phase 0  thread q of an element stages node q (coordinates, displacement, CSR row info).
//   phase 1  thread g = Gauss point g: Jacobian, dN/dX (G), F, C, material (S, Cm), then the
//            45 unique entries of the symmetric operator M (material + geometric tangent) and
//            Pw = F S w go to shared memory.  The same thread checks det J at element node g.
//   phase 2  thread i = row node i: loops the 8 Gauss points, accumulates f_i and the 3x3 blocks
//            K_ij for j = i, i+1, .., i+4 (mod 8; the +4 pair only for i < 4), i.e. the 36 blocks
//            of the upper triangle, in registers.
//   phase 3  each block and its transpose are added to the CSR values of the owned rows
//            (red.global.add.f64), f_i to the residual.
// All cross-thread traffic stays inside one warp, so phases are separated by __syncwarp() only.
#pragma once

#include <stdint.h>

#include "solid_math.cuh"

namespace b200s
{
  struct ErrWord
  {
    unsigned long long n_neg_jac;
    long long first_bad;     // min element index with det J <= 0 (LLONG_MAX if none)
    double min_detJ;
    double min_detF;
    unsigned long long n_nonfinite;
  };

  struct EvalArgs
  {
    const double* X;              // 3*n_col_nodes
    const double* u;              // 3*n_col_nodes, column map
    const int32_t* conn;          // n_ele*nn
    const int32_t* pos;           // n_ele*nn*nn: offset of column block j inside the rows of node i
    const int64_t* node_rowstart; // n_col_nodes: rowptr[row of dof 0] or -1 (not owned)
    const int32_t* node_rowlen;   // n_col_nodes: entries per row of that node
    const int32_t* node_frow;     // n_col_nodes: local row of dof 0 or -1
    const double* per_ele_param;  // nullptr or n_ele
    double* vals;                 // nnz (may be nullptr for force only)
    double* f;                    // n_rows
    ErrWord* err;
    MatParams mp;
    int64_t n_ele;
    int scatter_mode;  // 0: red.global.add (product); 1: skip scatter, 2: plain stores (both only
                       // for profiling the cost split, results are then meaningless)
  };

  namespace hex8
  {
    constexpr int NN = 8, NGP = 8, EPW = 4;  // nodes, Gauss points, elements per warp
    // shared-memory geometry (in doubles); strides chosen so that 64/128-bit accesses of a warp
    // fall on distinct banks (see DESIGN.md "shared memory layout")
    constexpr int TAB_STRIDE = 25;                 // per Gauss point / node row of a dN table
    constexpr int TAB_SIZE = 8 * TAB_STRIDE;       // one table
    constexpr int MP_PAIRS = 27;                   // 45 M + 9 Pw = 54 doubles = 27 x double2
    constexpr int MP_SLOTS = EPW * 9;              // slot = el*9 + g (9: pad)
    constexpr int MP_SIZE = MP_PAIRS * MP_SLOTS * 2;
    constexpr int G_GSTRIDE = 25, G_ESTRIDE = 200; // G[el][g][k*3+a]
    constexpr int G_SIZE = EPW * G_ESTRIDE;
    constexpr int XU_ESTRIDE = 50;                 // staging X[k*3+d], u at +24 (aliases MP)
    constexpr int WARP_DOUBLES = MP_SIZE + G_SIZE;
    constexpr int WARP_INTS = EPW * NN * 4;        // rowstart (2 ints), rowlen, frow per node

    struct WarpSmem
    {
      double* mp;    // MP_SIZE (XU staging aliases the front)
      double* g;     // G_SIZE
      int64_t* rowstart;  // EPW*NN
      int32_t* rowlen;    // EPW*NN
      int32_t* frow;      // EPW*NN
    };

    B2_HD int mp_index(int pair, int el, int g) { return (pair * MP_SLOTS + el * 9 + g) * 2; }

    // ---- phase 0: stage the element's nodal data --------------------------------------------
    B2_HD void phase0(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0)
    {
      const int el = lane >> 3, q = lane & 7;
      const int64_t e = e0 + el;
      double* xu = S.mp + el * XU_ESTRIDE;
      if (e < A.n_ele)
      {
        const int64_t node = A.conn[e * NN + q];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          xu[3 * q + d] = A.X[3 * node + d];
          xu[24 + 3 * q + d] = A.u[3 * node + d];  // extract_values through lm (calc_lib.hpp:232-246)
        }
        S.rowstart[lane] = A.node_rowstart[node];
        S.rowlen[lane] = A.node_rowlen[node];
        S.frow[lane] = A.node_frow[node];
      }
      else
      {
        // padding element: unit cube, zero displacement, nothing owned
        const double c[3] = {(q == 1 || q == 2 || q == 5 || q == 6) ? 1.0 : 0.0,
            (q == 2 || q == 3 || q == 6 || q == 7) ? 1.0 : 0.0, (q >= 4) ? 1.0 : 0.0};
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          xu[3 * q + d] = c[d];
          xu[24 + 3 * q + d] = 0.0;
        }
        S.rowstart[lane] = -1;
        S.rowlen[lane] = 0;
        S.frow[lane] = -1;
      }
    }

    // J(a,b) = sum_k dN(a,k) X(b,k)  (multiply_nt, calc_lib.hpp:506-507)
    B2_HD void jacobian(const double* dN /*[k*3+a]*/, const double* X /*[k*3+b]*/, double* J)
    {
#pragma unroll
      for (int i = 0; i < 9; ++i) J[i] = 0.0;
#pragma unroll
      for (int k = 0; k < NN; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * X[3 * k + b];
    }

    // ---- phase 1: one Gauss point per thread ---------------------------------------------------
    // tab_gp / tab_node: dN tables with row stride TAB_STRIDE; wgt: Gauss weights.
    // Results: G to S.g, (M, Pw) kept in registers `out54` for the caller to store after a
    // __syncwarp() (the staging area they overwrite is still being read by other lanes).
    template <int MAT, bool STIFF>
    B2_HD void phase1(const EvalArgs& A, const WarpSmem& S, const double* tab_gp,
        const double* tab_node, const double* wgt, int lane, int64_t e0, double* out54,
        double& detJ_min, double& detF_min, bool& neg_jac)
    {
      const int el = lane >> 3, g = lane & 7;
      const int64_t e = e0 + el;
      const double* xu = S.mp + el * XU_ESTRIDE;
      double J[9], iJ[9];

      // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562): node g
      jacobian(tab_node + g * TAB_STRIDE, xu, J);
      const double detJ_node = det3(J);

      // evaluate_jacobian_mapping at Gauss point g (calc_lib.hpp:499-516)
      const double* dN = tab_gp + g * TAB_STRIDE;
      jacobian(dN, xu, J);
      inv3(iJ, J);
      const double detJ = det3(J);
      neg_jac = !(detJ_node > 0.0) && (e < A.n_ele);
      detJ_min = fmin(detJ_node, detJ);

      // N_XYZ = J^-1 dN (:511-513) and F = I + u N_XYZ^T (:646-660)
      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double* Gout = S.g + el * G_ESTRIDE + g * G_GSTRIDE;
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        double Gk[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          Gk[a] = iJ[a] * dN[3 * k] + iJ[a + 3] * dN[3 * k + 1] + iJ[a + 6] * dN[3 * k + 2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          Gout[3 * k + a] = Gk[a];
          const double ua = xu[24 + 3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F[a + 3 * b] += ua * Gk[b];
        }
      }
      detF_min = det3(F);

      // material: So3Material::evaluate(F, E=(C-I)/2) -> S, Cm (calc_lib.hpp:976-989)
      double C[6], Sv[6], Cm[21];
      right_cauchy_green(F, C);
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr && e < A.n_ele) mp.p[0] = A.per_ele_param[e];
      material_evaluate<MAT>(mp, C, Sv, Cm);

      const double w = detJ * wgt[g];  // integration factor (calc_lib.hpp:1270-1280)
      if (STIFF) material_geometric_operator(F, Sv, Cm, w, out54);
      first_piola_w(F, Sv, w, out54 + 45);
    }

    template <bool STIFF>
    B2_HD void phase1_store(const WarpSmem& S, int lane, const double* out54)
    {
      const int el = lane >> 3, g = lane & 7;
#pragma unroll
      for (int p = STIFF ? 0 : 22; p < MP_PAIRS; ++p)
      {
        double* dst = S.mp + mp_index(p, el, g);
        dst[0] = out54[2 * p];
        dst[1] = out54[2 * p + 1];
      }
    }

    // value idx (0..44 M, 45..53 Pw) of Gauss point g; the two halves of a pair share one
    // 128-bit load after common-subexpression elimination
    B2_HD double mp_value(const WarpSmem& S, int el, int g, int idx)
    {
      const d2 t = *reinterpret_cast<const d2*>(S.mp + mp_index(idx >> 1, el, g));
      return (idx & 1) ? t.y : t.x;
    }

    // ---- phase 2: one row node per thread, loop over Gauss points --------------------------
    // Kacc[k][b + 3c] = K_{i,(i+k)%8}(b,c), facc = f_i
    template <bool STIFF>
    B2_HD void phase2(const WarpSmem& S, int lane, double (*Kacc)[9], double* facc)
    {
      const int el = lane >> 3, i = lane & 7;
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int q = 0; q < 9; ++q) Kacc[k][q] = 0.0;
      facc[0] = facc[1] = facc[2] = 0.0;

#pragma unroll 1
      for (int g = 0; g < NGP; ++g)
      {
        const double* Gg = S.g + el * G_ESTRIDE + g * G_GSTRIDE;
        const double Gi[3] = {Gg[3 * i], Gg[3 * i + 1], Gg[3 * i + 2]};
        // f_i += Pw G_i  (calc_lib.hpp:1116-1121)
        double Pw[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) Pw[r] = mp_value(S, el, g, 45 + r);
#pragma unroll
        for (int a = 0; a < 3; ++a)
          facc[a] += Pw[a] * Gi[0] + Pw[a + 3] * Gi[1] + Pw[a + 6] * Gi[2];
        if (STIFF)
        {
          // out[b][q] = sum_a Gi(a) M[(a,b)][q]: stream the packed upper triangle once, every
          // off-diagonal value feeds two accumulators
          double out[3][9];
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int q = 0; q < 9; ++q) out[b][q] = 0.0;
#pragma unroll
          for (int p = 0; p < 9; ++p)
#pragma unroll
            for (int q = p; q < 9; ++q)
            {
              const double m = mp_value(S, el, g, m45(p, q));
              out[p % 3][q] += Gi[p / 3] * m;
              if (p != q) out[q % 3][p] += Gi[q / 3] * m;
            }
          // K_ij(b,c) += sum_d out[b][3d+c] G_j(d)   (calc_lib.hpp:1135-1149)
#pragma unroll
          for (int k = 0; k < 5; ++k)
          {
            const int j = (i + k) & 7;
            const double Gj[3] = {Gg[3 * j], Gg[3 * j + 1], Gg[3 * j + 2]};
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                Kacc[k][b + 3 * c] +=
                    out[b][c] * Gj[0] + out[b][3 + c] * Gj[1] + out[b][6 + c] * Gj[2];
          }
        }
      }
    }

#if defined(__CUDA_ARCH__)
#define B2_RED_ADD(ptr, v)      \
  do                            \
  {                             \
    if (A.scatter_mode == 0)    \
      atomicAdd((ptr), (v));    \
    else                        \
      *(ptr) = (v);             \
  } while (0)
#else
#define B2_RED_ADD(ptr, v) (*(ptr) += (v))
#endif

    // ---- phase 3: owner-computes scatter-add --------------------------------------------------
    // Rows of node i: rowstart_i + b*rowlen_i; the three columns of node j start at pos(i,j).
    template <bool STIFF>
    B2_HD void phase3(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0,
        const double (*Kacc)[9], const double* facc)
    {
      const int el = lane >> 3, i = lane & 7;
      const int64_t e = e0 + el;
      if (e >= A.n_ele) return;
      if (A.scatter_mode == 1)
      {
        // profiling only: keep the accumulators alive without touching memory
        double s = facc[0] + facc[1] + facc[2];
        if (STIFF)
          for (int k = 0; k < 5; ++k)
            for (int q = 0; q < 9; ++q) s += Kacc[k][q];
        if (s == 1.2345e300) A.f[0] = s;
        return;
      }
      const int64_t rs_i = S.rowstart[lane];
      const int32_t len_i = S.rowlen[lane];
      const int32_t fr = S.frow[lane];
      if (fr >= 0)
      {
#pragma unroll
        for (int a = 0; a < 3; ++a) B2_RED_ADD(A.f + fr + a, facc[a]);
      }
      if (!STIFF) return;
      const int32_t* pos = A.pos + e * (NN * NN);
#pragma unroll
      for (int k = 0; k < 5; ++k)
      {
        if (k == 4 && i >= 4) break;
        const int j = (i + k) & 7;
        if (rs_i >= 0)
        {
          double* base = A.vals + rs_i + pos[i * NN + j];
#pragma unroll
          for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int c = 0; c < 3; ++c) B2_RED_ADD(base + (int64_t)b * len_i + c, Kacc[k][b + 3 * c]);
        }
        if (k > 0)
        {
          const int lj = (el << 3) + j;
          const int64_t rs_j = S.rowstart[lj];
          if (rs_j >= 0)
          {
            const int32_t len_j = S.rowlen[lj];
            double* base = A.vals + rs_j + pos[j * NN + i];
            // K_ji = K_ij^T
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
              for (int b = 0; b < 3; ++b)
                B2_RED_ADD(base + (int64_t)c * len_j + b, Kacc[k][b + 3 * c]);
          }
        }
      }
    }
  }  // namespace hex8
}  // namespace b200s

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
// Mapping: one warp = 4 elements x 8 threads, software pipelined over the warp's tiles.
//   gather   thread q of an element stages node q (coordinates, displacement, CSR row info) and
//            the warp stages the tile's scatter map; on the device these are cp.async copies
//            issued one tile ahead, so the dependent loads conn -> node -> X,u never stall math.
//   phase 1  thread g = Gauss point g: Jacobian, dN/dX (G), F, C, material (S, Cm), then the
//            45 unique entries of the symmetric operator M (material + geometric tangent) and
//            Pw = F S w go to shared memory.  The same thread checks det J at element node g.
//   phase 2  thread i = row node i: loops the 8 Gauss points, accumulates f_i and the 3x3 blocks
//            K_ij for j = i, i+1, .., i+4 (mod 8; the +4 pair only for i < 4), i.e. the 36 blocks
//            of the upper triangle, in registers.
//   stage    the 36 blocks and their transposes are written to shared memory as the full
//            row-major 24x24 K_e (aliasing the dead M/G buffers), f_e next to it.
//   scatter  the warp walks K_e row by row: 32 consecutive lanes = 32 consecutive entries of
//            K_e = runs of 24 consecutive columns of one CSR row (3-6 contiguous doubles per
//            column node), added with red.global.add.f64 to the owned rows; f_e to the residual.
//            Coalesced runs reach 2.5x the RED rate of per-lane scattered rows (profiles/).
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

  // CSR rows of one column node: rowstart = rowptr[row of dof 0] or -1 (not owned by this rank),
  // rowlen = entries per row (equal for the three dofs), frow = local row of dof 0 or -1
  struct alignas(16) NodeRow
  {
    int64_t rowstart;
    int32_t rowlen;
    int32_t frow;
  };

  struct EvalArgs
  {
    const double* X;              // 3*n_col_nodes
    const double* u;              // 3*n_col_nodes, column map
    const int32_t* conn;          // n_ele*nn
    const int32_t* pos;           // n_ele*nn*nn: offset of column block j inside the rows of node i
    const NodeRow* node_row;      // n_col_nodes: CSR rows of the node's three dofs
    const double* per_ele_param;  // nullptr or n_ele (indexed like conn, i.e. in device order)
    const int32_t* ele_id;        // nullptr or n_ele: caller's column-element index of device element e
    const double* fibers;         // nullptr or 3*n_ele unit fibre vectors (FIBER1), device element order
    const double* gp_hist;        // MULF only: n_ele * ngp * MULF_HIST doubles in device element order
    double* mat_hist;             // history material (MAT_DP): n_ele * ngp * DP_HIST doubles, device element order
    double* vals;                 // nnz (may be nullptr for force only)
    double* f;                    // n_rows
    ErrWord* err;
    MatParams mp;
    int64_t n_ele;
    int linear;        // KINEM linear (solid_math.cuh linear_kinematics): a template parameter of the element
                       // kernels (the headline path pays nothing); this flag only labels the launch
    int scatter_mode;  // 0: red.global.add (product); 1: skip scatter, 2: plain stores (both only
                       // for profiling the cost split, results are then meaningless)
    int64_t e_begin = 0;  // first device element of this launch's segment: every per-element pointer above is
                          // already offset by it (the kernels index from 0); launch wrappers with further
                          // per-element arrays (EAS history) use it
    // hex20 blocks through the hex27 kernel (b200_solid.cu, launch_hex20_as_hex27): shape tables of the serendipity cell
    // padded to 27 nodes (device pointer to a Hex27Const, nullptr = the hex27 tables) and the number of real nodes --
    // element nodes >= nn_real are padding: no rows, no columns, no nodal Jacobian check inside the kernel
    const void* tab_override = nullptr;
    int nn_real = 0;  // 0: every node is real
  };

  namespace hex8
  {
    constexpr int NN = 8, NGP = 8, EPW = 4;  // nodes, Gauss points, elements per warp
    // shared-memory geometry (in doubles); strides chosen so that 64/128-bit accesses of a warp
    // fall on distinct banks (see DESIGN.md "shared memory layout")
    constexpr int TAB_STRIDE = 26;                 // per Gauss point / node row of a dN table (even: rows 16-byte aligned)
    constexpr int TAB_SIZE = 8 * TAB_STRIDE;       // one table
    constexpr int MP_PAIRS = 27;                   // 45 M + 9 Pw = 54 doubles = 27 x double2
    constexpr int MP_SLOTS = EPW * 9;              // slot = el*9 + g (9: pad)
    constexpr int MP_SIZE = MP_PAIRS * MP_SLOTS * 2;
    constexpr int G_GSTRIDE = 25, G_ESTRIDE = 200; // G[el][g][k*3+a]
    constexpr int G_SIZE = EPW * G_ESTRIDE;
    constexpr int XU_ESTRIDE = 50;                 // gathered X[k*3+d], u at +24
    constexpr int XU_SIZE = EPW * XU_ESTRIDE;
    constexpr int KE_ESTRIDE = 24 * 24 + 8;        // staged K_e per element (+8: bank offset of 16)
    constexpr int FE_OFFSET = EPW * KE_ESTRIDE;    // staged f_e (EPW*24) behind the K_e's
    constexpr int WARP_DOUBLES = MP_SIZE + G_SIZE + XU_SIZE;
    static_assert(FE_OFFSET + EPW * 24 <= MP_SIZE + G_SIZE, "K_e/f_e staging aliases M/Pw/G");
    constexpr int WARP_BYTES = WARP_DOUBLES * 8 + EPW * NN * 16 + EPW * NN * NN * 4 + 16;  // + 16: the warp's mbarrier

    struct WarpSmem
    {
      double* mp;    // MP_SIZE; K_e / f_e staging aliases mp..g after phase 2
      double* g;     // G_SIZE (directly behind mp)
      double* xu;    // XU_SIZE
      NodeRow* nrow; // EPW*NN
      int32_t* pos;  // EPW*NN*NN
      double* h0 = nullptr;  // F-bar only (hex8_fbar_kernel.cuh): EPW*24
    };

    B2_HD int mp_index(int pair, int el, int g) { return (pair * MP_SLOTS + el * 9 + g) * 2; }

    // ---- gather: stage the tile's nodal data and scatter map --------------------------------
    // padding element (beyond n_ele): unit cube, zero displacement, nothing owned
    B2_HD void gather_pad(const WarpSmem& S, int lane)
    {
      const int el = lane >> 3, q = lane & 7;
      double* xu = S.xu + el * XU_ESTRIDE;
      const double c[3] = {(q == 1 || q == 2 || q == 5 || q == 6) ? 1.0 : 0.0,
          (q == 2 || q == 3 || q == 6 || q == 7) ? 1.0 : 0.0, (q >= 4) ? 1.0 : 0.0};
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        xu[3 * q + d] = c[d];
        xu[24 + 3 * q + d] = 0.0;
      }
    }
    // synchronous flavour (host emulation of the device's cp.async pipeline)
    B2_HD void gather(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0)
    {
      const int el = lane >> 3, q = lane & 7;
      const int64_t e = e0 + el;
      double* xu = S.xu + el * XU_ESTRIDE;
      if (e < A.n_ele)
      {
        const int64_t node = A.conn[e * NN + q];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          xu[3 * q + d] = A.X[3 * node + d];
          xu[24 + 3 * q + d] = A.u[3 * node + d];  // extract_values through lm (calc_lib.hpp:232-246)
        }
        S.nrow[lane] = A.node_row[node];
#pragma unroll
        for (int j = 0; j < NN; ++j) S.pos[lane * NN + j] = A.pos[e * (NN * NN) + q * NN + j];
      }
      else
      {
        gather_pad(S, lane);
        S.nrow[lane].rowstart = -1, S.nrow[lane].rowlen = 0, S.nrow[lane].frow = -1;
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
    template <int MAT, bool STIFF, bool MULF = false, bool LINEAR = false>
    B2_HD void phase1(const EvalArgs& A, const WarpSmem& S, const double* tab_gp,
        const double* tab_node, const double* wgt, int lane, int64_t e0, double* out54,
        double& detJ_min, double& detF_min, bool& neg_jac)
    {
      const int el = lane >> 3, g = lane & 7;
      const int64_t e = e0 + el;
      const double* xu = S.xu + el * XU_ESTRIDE;
      double J[9], iJ[9];

      // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562): node g.
      // At a hex8 corner only the node itself and its three edge neighbours have non-zero
      // parametric derivatives (+-1/2), so the reference's sum over 8 nodes reduces exactly to
      // J(a,:) = xi_a/2 (X_g - X_nbr_a); tab_node (the full table) documents and tests that.
      (void)tab_node;
      {
        const double sr = ((g ^ (g >> 1)) & 1) ? 0.5 : -0.5, ss = (g & 2) ? 0.5 : -0.5,
                     st = (g & 4) ? 0.5 : -0.5;
        const double *xq = xu + 3 * g, *xr = xu + 3 * (g ^ 1), *xs = xu + 3 * (g ^ 3),
                     *xt = xu + 3 * (g ^ 4);
#pragma unroll
        for (int b = 0; b < 3; ++b)
        {
          J[0 + 3 * b] = sr * (xq[b] - xr[b]);
          J[1 + 3 * b] = ss * (xq[b] - xs[b]);
          J[2 + 3 * b] = st * (xq[b] - xt[b]);
        }
      }
      const double detJ_node = det3(J);

      // evaluate_jacobian_mapping at Gauss point g (calc_lib.hpp:499-516)
      const double* dNs = tab_gp + g * TAB_STRIDE;
      double dN[24];
#if defined(__CUDA_ARCH__)
#pragma unroll
      for (int p = 0; p < 12; ++p)
      {
        const d2 t = *reinterpret_cast<const d2*>(dNs + 2 * p);
        dN[2 * p] = t.x, dN[2 * p + 1] = t.y;
      }
#else
      for (int p = 0; p < 24; ++p) dN[p] = dNs[p];
#endif
      // the eight lanes of an element read the same X and u: 128-bit loads halve the shared-memory wavefronts
      double xr[24], ur[24];
#if defined(__CUDA_ARCH__)
#pragma unroll
      for (int p = 0; p < 12; ++p)
      {
        const d2 tx = *reinterpret_cast<const d2*>(xu + 2 * p);  // xu is 16-byte aligned (XU_ESTRIDE even)
        const d2 tu = *reinterpret_cast<const d2*>(xu + 24 + 2 * p);
        xr[2 * p] = tx.x, xr[2 * p + 1] = tx.y, ur[2 * p] = tu.x, ur[2 * p + 1] = tu.y;
      }
#else
      for (int p = 0; p < 24; ++p) xr[p] = xu[p], ur[p] = xu[24 + p];
#endif
      jacobian(dN, xr, J);
      const double detJ = inv3(iJ, J);  // same expression as det3(J)
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
          const double ua = ur[3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F[a + 3 * b] += ua * Gk[b];
        }
      }
      if (MULF && e < A.n_ele)  // MulfFormulation::evaluate: only F changes, the gradients stay
        mulf_deformation_gradient<NN>(A.gp_hist + (e * NGP + g) * MULF_HIST, dN, xu + 24, F);
      detF_min = det3(F);

      // material: So3Material::evaluate(F, E=(C-I)/2) -> S, Cm (calc_lib.hpp:976-989)
      double C[6], Sv[6], Cm[21];
      right_cauchy_green(F, C);
      if (LINEAR) linear_kinematics(F, C);  // compile-time in this kernel (TECH 3): the headline path pays nothing
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr && e < A.n_ele) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr && e < A.n_ele) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr && e < A.n_ele)
        mp.hist = A.mat_hist + (e * NGP + g) * material_hist_doubles(MAT);
      material_evaluate<MAT>(mp, C, Sv, Cm);

      // integration factor (calc_lib.hpp:1270-1280), folded into S and Cm once
      const double w = detJ * wgt[g];
#pragma unroll
      for (int v = 0; v < 6; ++v) Sv[v] *= w;
      if (STIFF)
      {
#pragma unroll
        for (int v = 0; v < 21; ++v) Cm[v] *= w;
        if (LINEAR)
        {
          const double zero[6] = {0, 0, 0, 0, 0, 0};  // no geometric stiffness
          material_geometric_operator(F, zero, Cm, out54);
        }
        else
          material_geometric_operator(F, Sv, Cm, out54);
      }
      first_piola(F, Sv, out54 + 45);
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
          facc[a] = fma(Pw[a + 6], Gi[2], fma(Pw[a + 3], Gi[1], fma(Pw[a], Gi[0], facc[a])));
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
              out[p % 3][q] = fma(Gi[p / 3], m, out[p % 3][q]);
              if (p != q) out[q % 3][p] = fma(Gi[q / 3], m, out[q % 3][p]);
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
                Kacc[k][b + 3 * c] = fma(out[b][6 + c], Gj[2],
                    fma(out[b][3 + c], Gj[1], fma(out[b][c], Gj[0], Kacc[k][b + 3 * c])));
          }
        }
      }
    }

#if defined(__CUDA_ARCH__) && defined(B200S_L2_HINTS) && (B200S_L2_HINTS & 2)
    // experiment (profiles/r02w_summary.md): REDs into the CSR values with an L2 evict-last policy
    __device__ __forceinline__ void red_add_keep(double* p, double v)
    {
      unsigned long long pol;
      asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol));
    }
#define B2_RED_ADD(ptr, v) ::b200s::hex8::red_add_keep((ptr), (v))
#elif defined(__CUDA_ARCH__)
#define B2_RED_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define B2_RED_ADD(ptr, v) (*(ptr) += (v))
#endif

    // ---- stage: accumulators -> full row-major K_e (and f_e) in shared memory -------------------
    // Caller guarantees (__syncwarp) that no lane still reads M/Pw/G, which this overwrites.
    // Bank check: lanes of one element differ by 75 doubles (= 22 banks of 4 bytes), the next
    // element starts 16 banks further: the 16 lanes of a 64-bit store phase hit 16 distinct pairs.
    template <bool STIFF>
    B2_HD void stage_ke(const WarpSmem& S, int lane, const double (*Kacc)[9], const double* facc)
    {
      const int el = lane >> 3, i = lane & 7;
      double* fe = S.mp + FE_OFFSET + el * 24;
      fe[3 * i] = facc[0], fe[3 * i + 1] = facc[1], fe[3 * i + 2] = facc[2];
      if (!STIFF) return;
      double* Ke = S.mp + el * KE_ESTRIDE;
#pragma unroll
      for (int k = 0; k < 5; ++k)
      {
        if (k == 4 && i >= 4) break;
        const int j = (i + k) & 7;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c)
          {
            Ke[(3 * i + b) * 24 + 3 * j + c] = Kacc[k][b + 3 * c];
            if (k > 0) Ke[(3 * j + c) * 24 + 3 * i + b] = Kacc[k][b + 3 * c];  // K_ji = K_ij^T
          }
      }
    }

    // ---- scatter: owner-computes scatter-add of the staged K_e / f_e ------------------------------
    // Rows of node i: rowstart_i + b*rowlen_i; the three columns of node j start at pos(i,j)
    // (SparseMatrix::assemble fast path, 4C_linalg_sparsematrix.cpp:404-441).  Lane l < 24 owns
    // column l of K_e (column node l/3, dof l%3) and walks the 24 rows: every RED instruction adds
    // 24 consecutive entries of one K_e row into one CSR row.
    template <bool STIFF>
    B2_HD void scatter(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0)
    {
      if (lane >= 24) return;
      const int jl = lane / 3, cl = lane - 3 * jl;
#pragma unroll 1
      for (int el = 0; el < EPW; ++el)
      {
        if (e0 + el >= A.n_ele) break;
        const NodeRow* nr = S.nrow + el * NN;
        {
          const int32_t fr = nr[jl].frow;
          if (fr >= 0) B2_RED_ADD(A.f + fr + cl, S.mp[FE_OFFSET + el * 24 + lane]);
        }
        if (!STIFF) continue;
        const double* Ke = S.mp + el * KE_ESTRIDE + lane;
        const int32_t* pos = S.pos + el * (NN * NN) + jl;
        // all shared-memory reads first (row info, map entry, the lane's 24 K_e entries), then the
        // REDs: no branch between the loads of different row nodes, so they overlap
        int64_t off[NN];
        int32_t len[NN];
#pragma unroll
        for (int i = 0; i < NN; ++i)
        {
          const NodeRow r = nr[i];
          off[i] = r.rowstart >= 0 ? r.rowstart + pos[i * NN] + cl : -1;  // -1: row node not owned
          len[i] = r.rowlen;
        }
#pragma unroll
        for (int i = 0; i < NN; ++i)
        {
          const double v0 = Ke[(3 * i) * 24], v1 = Ke[(3 * i + 1) * 24], v2 = Ke[(3 * i + 2) * 24];
          if (off[i] >= 0)
          {
            double* dst = A.vals + off[i];
            B2_RED_ADD(dst, v0);
            B2_RED_ADD(dst + len[i], v1);
            B2_RED_ADD(dst + 2 * (int64_t)len[i], v2);
          }
        }
      }
    }
  }  // namespace hex8
}  // namespace b200s

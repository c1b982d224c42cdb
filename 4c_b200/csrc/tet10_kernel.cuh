// tet10_kernel.cuh -- evaluate + assemble for tet10 displacement-based nonlinear solid elements
// (4 Gauss points for force/stiffness: tet_4point, solid_ele/4C_solid_ele_calc_lib_integration.hpp:55-58).
//
// Same phase structure as hex8_kernel.cuh (read that header first); the differences:
//   * one warp = 3 elements x 10 lanes (lanes 30,31 idle), lane = (element, node) in the gather,
//     phase 2, stage; lanes with node index < 4 additionally run one Gauss point each in phase 1;
//   * the nodal Jacobian check (calc_lib.hpp:543-562) is a full sum over the 10 shape-function
//     derivatives at node q, done by lane (element, q) before phase 1;
//   * phase 2: row node i accumulates the blocks K_ij, j = i, i+1, .., i+5 (mod 10; the +5 pair only
//     for i < 5) = the 55 blocks of the upper triangle; K_e is 30 x 30;
//   * the scatter walks 30 rows with lanes l < 30 = column l of K_e.
// tet10 is the scatter-heavy cell of the three: 930 RED lanes per 13 k FMA (hex8: 600 per 20 k).
#pragma once

#include <stdint.h>

#include "hex8_kernel.cuh"

namespace b200s
{
  namespace tet10
  {
    constexpr int NN = 10, NGP = 4, EPW = 3, NDOF = 30, NBLK = 6;
    constexpr int TAB_STRIDE = 31;                  // per Gauss point / node row of a dN table (30 used)
    constexpr int TAB_GP_SIZE = NGP * TAB_STRIDE, TAB_NODE_SIZE = NN * TAB_STRIDE;
    constexpr int MP_PAIRS = 27;                    // 45 M + 9 Pw
    constexpr int MP_SLOTS = EPW * NGP + 1;         // slot = el*4 + g (+1: pad)
    constexpr int MP_SIZE = MP_PAIRS * MP_SLOTS * 2;
    constexpr int G_GSTRIDE = 31, G_ESTRIDE = 129;  // G[el][g][k*3+a]; 129 = 1 mod 16: the 64-bit
                                                     // loads of 3 elements x 10 lanes (stride 3) miss each other
    constexpr int G_SIZE = EPW * G_ESTRIDE;
    constexpr int XU_ESTRIDE = 62;                  // gathered X[k*3+d], u at +30
    constexpr int XU_SIZE = EPW * XU_ESTRIDE;
    // staged K_e: row stride 32.  With j = (i + k) mod 10 the direct store of lane i lands on bank pair
    // (3*32 + 3) i + 3 k = 3 i + const (mod 16), the wrapped column blocks (j - 10: -30 doubles = +2) can only meet a
    // lane i + 6, which is wrapped as well, and the transposed store (row 3 j + c) wraps by a multiple of 16:
    // both are conflict free inside an element (stride 30 gave 13 i and 2-way conflicts on up to four lanes).
    constexpr int KE_RS = 32;
    constexpr int KE_ESTRIDE = NDOF * KE_RS + 2;    // per element (962 = 2 mod 16)
    constexpr int FE_OFFSET = EPW * KE_ESTRIDE;
    constexpr int STAGE_SIZE = FE_OFFSET + EPW * NDOF + 2;  // M/Pw/G alias the front of the stage
    static_assert(MP_SIZE + G_SIZE <= STAGE_SIZE, "M/Pw/G live inside the K_e staging area");
    constexpr int WARP_DOUBLES = STAGE_SIZE + XU_SIZE;
    constexpr int POS_INTS = EPW * NN * NN;         // 300 ints = 75 chunks of 16 bytes
    constexpr int WARP_BYTES = WARP_DOUBLES * 8 + 32 * 16 + ((POS_INTS * 4 + 15) / 16) * 16;

    struct WarpSmem
    {
      double* mp;    // MP_SIZE; the K_e / f_e staging (STAGE_SIZE) starts here too
      double* g;     // G_SIZE (directly behind mp)
      double* xu;    // XU_SIZE
      NodeRow* nrow; // 32 (30 used)
      int32_t* pos;  // POS_INTS
    };

    B2_HD int mp_index(int pair, int el, int g) { return (pair * MP_SLOTS + el * NGP + g) * 2; }

    // unit tet with mid-edge nodes: padding element beyond n_ele
    B2_HD void gather_pad(const WarpSmem& S, int lane)
    {
      const int el = lane / NN, q = lane - el * NN;
      if (el >= EPW) return;
      // node coordinates of the unit tet in half units, 2 bits per entry, entry = 3*q + d:
      // corners (0,0,0),(2,0,0),(0,2,0),(0,0,2); mid-edge nodes (1,0,0),(1,1,0),(0,1,0),(0,0,1),
      // (1,0,1),(0,1,1)   (4C_fem_general_utils_local_connectivity_matrices.hpp:314)
      unsigned long long code = 0;
      const int half[30] = {0, 0, 0, 2, 0, 0, 0, 2, 0, 0, 0, 2, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0,
          1, 0, 1, 1};
#pragma unroll
      for (int t = 0; t < 30; ++t) code |= (unsigned long long)half[t] << (2 * t);
      double* xu = S.xu + el * XU_ESTRIDE;
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        xu[3 * q + d] = 0.5 * (double)((code >> (2 * (3 * q + d))) & 3ull);
        xu[NDOF + 3 * q + d] = 0.0;
      }
    }

    // synchronous flavour (host emulation of the device's cp.async pipeline)
    B2_HD void gather(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0)
    {
      const int el = lane / NN, q = lane - el * NN;
      if (el >= EPW) return;
      const int64_t e = e0 + el;
      double* xu = S.xu + el * XU_ESTRIDE;
      if (e < A.n_ele)
      {
        const int64_t node = A.conn[e * NN + q];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          xu[3 * q + d] = A.X[3 * node + d];
          xu[NDOF + 3 * q + d] = A.u[3 * node + d];
        }
        S.nrow[lane] = A.node_row[node];
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
          for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(dN[3 * k + a], X[3 * k + b], J[a + 3 * b]);
    }

    // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562): lane = node q
    B2_HD double node_jacobian_det(const WarpSmem& S, const double* tab_node, int lane)
    {
      const int el = lane / NN, q = lane - el * NN;
      if (el >= EPW) return 1e300;
      double J[9];
      jacobian(tab_node + q * TAB_STRIDE, S.xu + el * XU_ESTRIDE, J);
      return det3(J);
    }

    // ---- phase 1: lanes (el, g < 4) run one Gauss point each --------------------------------------
    template <int MAT, bool STIFF, bool MULF = false, bool LINEAR = false>
    B2_HD void phase1(const EvalArgs& A, const WarpSmem& S, const double* tab_gp, const double* wgt,
        int lane, int64_t e0, double* out54, double& detJ_gp, double& detF_gp)
    {
      const int el = lane / NN, g = lane - el * NN;
      detJ_gp = detF_gp = 1e300;
      if (el >= EPW || g >= NGP) return;
      const int64_t e = e0 + el;
      const double* xu = S.xu + el * XU_ESTRIDE;
      double J[9], iJ[9];
      const double* dN = tab_gp + g * TAB_STRIDE;
      jacobian(dN, xu, J);
      const double detJ = inv3(iJ, J);
      detJ_gp = detJ;

      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double* Gout = S.g + el * G_ESTRIDE + g * G_GSTRIDE;
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        double Gk[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          Gk[a] = fma(iJ[a + 6], dN[3 * k + 2], fma(iJ[a + 3], dN[3 * k + 1], iJ[a] * dN[3 * k]));
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          Gout[3 * k + a] = Gk[a];
          const double ua = xu[NDOF + 3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F[a + 3 * b] = fma(ua, Gk[b], F[a + 3 * b]);
        }
      }
      if (MULF && e < A.n_ele)
        mulf_deformation_gradient<NN>(A.gp_hist + (e * NGP + g) * MULF_HIST, dN, xu + NDOF, F);
      detF_gp = det3(F);

      double C[6], Sv[6], Cm[21];
      right_cauchy_green(F, C);
      if (LINEAR) linear_kinematics(F, C);
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr && e < A.n_ele) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr && e < A.n_ele) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr && e < A.n_ele)
        mp.hist = A.mat_hist + (e * NGP + g) * material_hist_doubles(MAT);
      material_evaluate<MAT>(mp, C, Sv, Cm);

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

    // ---- phase 1 split in two (device kernel): the first half ends with the material response kept in registers
    // (F, w S, w C = 36 doubles), the second half -- after the X/u refill has been issued -- forms the operator M
    // and P w and stores them; the 54 operator entries never have to stay live across the refill.
    template <int MAT, bool STIFF, bool MULF = false, bool LINEAR = false>
    B2_HD void phase1a(const EvalArgs& A, const WarpSmem& S, const double* tab_gp, const double* wgt, int lane,
        int64_t e0, double* F, double* Sv, double* Cm, double& detJ_gp, double& detF_gp)
    {
      const int el = lane / NN, g = lane - el * NN;
      detJ_gp = detF_gp = 1e300;
      if (el >= EPW || g >= NGP) return;
      const int64_t e = e0 + el;
      const double* xu = S.xu + el * XU_ESTRIDE;
      double J[9], iJ[9];
      const double* dN = tab_gp + g * TAB_STRIDE;
      jacobian(dN, xu, J);
      const double detJ = inv3(iJ, J);
      detJ_gp = detJ;
#pragma unroll
      for (int q = 0; q < 9; ++q) F[q] = (q % 4 == 0) ? 1.0 : 0.0;
      double* Gout = S.g + el * G_ESTRIDE + g * G_GSTRIDE;
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        double Gk[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          Gk[a] = fma(iJ[a + 6], dN[3 * k + 2], fma(iJ[a + 3], dN[3 * k + 1], iJ[a] * dN[3 * k]));
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          Gout[3 * k + a] = Gk[a];
          const double ua = xu[NDOF + 3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F[a + 3 * b] = fma(ua, Gk[b], F[a + 3 * b]);
        }
      }
      if (MULF && e < A.n_ele)
        mulf_deformation_gradient<NN>(A.gp_hist + (e * NGP + g) * MULF_HIST, dN, xu + NDOF, F);
      detF_gp = det3(F);
      double C[6];
      right_cauchy_green(F, C);
      if (LINEAR) linear_kinematics(F, C);
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr && e < A.n_ele) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr && e < A.n_ele) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr && e < A.n_ele)
        mp.hist = A.mat_hist + (e * NGP + g) * material_hist_doubles(MAT);
      material_evaluate<MAT>(mp, C, Sv, Cm);
      const double w = detJ * wgt[g];
#pragma unroll
      for (int v = 0; v < 6; ++v) Sv[v] *= w;
      if (STIFF)
      {
#pragma unroll
        for (int v = 0; v < 21; ++v) Cm[v] *= w;
      }
    }
    template <bool STIFF, bool LINEAR = false>
    B2_HD void phase1b(const WarpSmem& S, int lane, const double* F, const double* Sv, const double* Cm)
    {
      const int el = lane / NN, g = lane - el * NN;
      if (el >= EPW || g >= NGP) return;
      double out54[54];
      if (STIFF)
      {
        if (LINEAR)
        {
          const double zero[6] = {0, 0, 0, 0, 0, 0};  // no geometric stiffness
          material_geometric_operator(F, zero, Cm, out54);
        }
        else
          material_geometric_operator(F, Sv, Cm, out54);
      }
      first_piola(F, Sv, out54 + 45);
#pragma unroll
      for (int p = STIFF ? 0 : 22; p < MP_PAIRS; ++p)
      {
        double* dst = S.mp + mp_index(p, el, g);
        dst[0] = out54[2 * p];
        dst[1] = out54[2 * p + 1];
      }
    }

    template <bool STIFF>
    B2_HD void phase1_store(const WarpSmem& S, int lane, const double* out54)
    {
      const int el = lane / NN, g = lane - el * NN;
      if (el >= EPW || g >= NGP) return;
#pragma unroll
      for (int p = STIFF ? 0 : 22; p < MP_PAIRS; ++p)
      {
        double* dst = S.mp + mp_index(p, el, g);
        dst[0] = out54[2 * p];
        dst[1] = out54[2 * p + 1];
      }
    }

    B2_HD double mp_value(const WarpSmem& S, int el, int g, int idx)
    {
      const d2 t = *reinterpret_cast<const d2*>(S.mp + mp_index(idx >> 1, el, g));
      return (idx & 1) ? t.y : t.x;
    }

    // ---- phase 2: one row node per lane, loop over the Gauss points -------------------------------
    // Kacc[k][b + 3c] = K_{i,(i+k)%10}(b,c), facc = f_i
    template <bool STIFF>
    B2_HD void phase2(const WarpSmem& S, int lane, double (*Kacc)[9], double* facc)
    {
      const int el = lane / NN, i = lane - el * NN;
#pragma unroll
      for (int k = 0; k < NBLK; ++k)
#pragma unroll
        for (int q = 0; q < 9; ++q) Kacc[k][q] = 0.0;
      facc[0] = facc[1] = facc[2] = 0.0;
      if (el >= EPW) return;

#if defined(B200S_TET10_UNROLL_GP) && B200S_TET10_UNROLL_GP  // experiment: profiles/r02w_summary.md
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int g = 0; g < NGP; ++g)
      {
        const double* Gg = S.g + el * G_ESTRIDE + g * G_GSTRIDE;
        const double Gi[3] = {Gg[3 * i], Gg[3 * i + 1], Gg[3 * i + 2]};
        double Pw[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) Pw[r] = mp_value(S, el, g, 45 + r);
#pragma unroll
        for (int a = 0; a < 3; ++a)
          facc[a] = fma(Pw[a + 6], Gi[2], fma(Pw[a + 3], Gi[1], fma(Pw[a], Gi[0], facc[a])));
        if (STIFF)
        {
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
#pragma unroll
          for (int k = 0; k < NBLK; ++k)
          {
            int j = i + k;
            j = j >= NN ? j - NN : j;
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

    // ---- stage: accumulators -> full row-major 30x30 K_e (and f_e) in shared memory -------------
    template <bool STIFF>
    B2_HD void stage_ke(const WarpSmem& S, int lane, const double (*Kacc)[9], const double* facc)
    {
      const int el = lane / NN, i = lane - el * NN;
      if (el >= EPW) return;
      double* fe = S.mp + FE_OFFSET + el * NDOF;
      fe[3 * i] = facc[0], fe[3 * i + 1] = facc[1], fe[3 * i + 2] = facc[2];
      if (!STIFF) return;
      double* Ke = S.mp + el * KE_ESTRIDE;
#pragma unroll
      for (int k = 0; k < NBLK; ++k)
      {
        if (k == NBLK - 1 && i >= NN / 2) break;
        int j = i + k;
        j = j >= NN ? j - NN : j;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c)
          {
            Ke[(3 * i + b) * KE_RS + 3 * j + c] = Kacc[k][b + 3 * c];
            if (k > 0) Ke[(3 * j + c) * KE_RS + 3 * i + b] = Kacc[k][b + 3 * c];  // K_ji = K_ij^T
          }
      }
    }

    // ---- scatter: lane l < 30 owns column l of K_e and walks the 30 rows -------------------------
    template <bool STIFF>
    B2_HD void scatter(const EvalArgs& A, const WarpSmem& S, int lane, int64_t e0)
    {
      if (lane >= NDOF) return;
      const int jl = lane / 3, cl = lane - 3 * jl;
#pragma unroll 1
      for (int el = 0; el < EPW; ++el)
      {
        if (e0 + el >= A.n_ele) break;
        const NodeRow* nr = S.nrow + el * NN;
        {
          const int32_t fr = nr[jl].frow;
          if (fr >= 0) B2_RED_ADD(A.f + fr + cl, S.mp[FE_OFFSET + el * NDOF + lane]);
        }
        if (!STIFF) continue;
        const double* Ke = S.mp + el * KE_ESTRIDE + lane;
        const int32_t* pos = S.pos + el * (NN * NN) + jl;
        // one row node per step, loads and REDs interleaved.  (Measured: hoisting all shared-memory reads ahead of the
        // REDs, as hex8::scatter does, removes the latency stalls ncu shows here and makes the kernel SLOWER, 14.40 ->
        // 14.86 ms on config 4 -- the REDs themselves are the limit, 3.56e9 lane-adds in runs of three doubles against
        // 412 G/s of the run-3 microbenchmark = 8.6 ms, and issuing them in bursts only backs up the LSU queue;
        // profiles/r02w_summary.md)
#pragma unroll
        for (int i = 0; i < NN; ++i)
        {
          const NodeRow r = nr[i];
          if (r.rowstart < 0) continue;
          double* dst = A.vals + (r.rowstart + pos[i * NN] + cl);
          B2_RED_ADD(dst, Ke[(3 * i) * KE_RS]);
          B2_RED_ADD(dst + r.rowlen, Ke[(3 * i + 1) * KE_RS]);
          B2_RED_ADD(dst + 2 * (int64_t)r.rowlen, Ke[(3 * i + 2) * KE_RS]);
        }
      }
    }
  }  // namespace tet10
}  // namespace b200s

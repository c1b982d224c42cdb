// hex27_kernel.cuh -- evaluate + assemble for hex27 displacement-based nonlinear solid elements
// (27 nodes, 81 dofs, 27 Gauss points hex_27point with the reference's truncated constants).
//
// K_e is 81 x 81 (52 KB): it neither fits the registers of one warp nor is it staged as a whole.
// One element is processed by a PAIR of warps w = 0,1 that share the element's shared-memory data:
//   gather   lanes (w, l) stage X, u, CSR row info of the 27 nodes and the 729 (+3 pad) scatter-map
//            entries with cp.async, one element ahead (same pipeline as hex8_kernel.cuh).
//   phase 1  one Gauss point per lane: warp 0 takes points 0..13, warp 1 points 14..26; the idle lanes
//            of both warps run the nodal Jacobian check (calc_lib.hpp:543-562) of the 27 nodes, using
//            that at a node only the three nodes of the same parameter line have non-zero derivatives.
//            Outputs per Gauss point: dN/dX (81 doubles), the 45 unique entries of the symmetric 9x9
//            operator M and P w = F S w (see solid_math.cuh) -> shared memory.
//   phase 2  lane i < 27 = row node i.  The 378 upper-triangle node pairs are {i, i+k mod 27},
//            k = 0..13; warp w accumulates the 7 blocks k = 7w .. 7w+6 of its row node over the 27
//            Gauss points in registers (63 doubles); `out = dN_i . M` (81 FMA) is formed once per
//            warp and Gauss point.  Warp 0 also accumulates f_i.
//   scatter  in two sub-passes (4 + 3 blocks) the blocks are staged per lane and added to the CSR rows:
//            for row node i a warp contributes the column nodes i+k directly and i-k transposed, every
//            RED instruction covers one CSR row with up to 12 column entries in 3-double runs.
// Pair-level synchronisation uses named barriers (bar.sync id, 64); everything else is warp-local.
#pragma once

#include <stdint.h>

#include "hex8_kernel.cuh"

namespace b200s
{
  namespace hex27
  {
    constexpr int NN = 27, NGP = 27, NDOF = 81, WPE = 2, NBLK = 7, SUB = 4;
    constexpr int GP_W0 = 14;                      // Gauss points of warp 0 (warp 1: 13)
    constexpr int TAB_STRIDE = 81;                 // dN table row (162 words = 2 mod 32: conflict free)
    constexpr int TAB_SIZE = NGP * TAB_STRIDE;
    constexpr int G_GSTRIDE = 81, G_SIZE = NGP * G_GSTRIDE;  // G[g][k*3+a]
    constexpr int MP_PAIRS = 27, MP_SLOTS = 28;    // 45 M + 9 Pw as 27 double2, slot = g
    constexpr int MP_SIZE = MP_PAIRS * MP_SLOTS * 2;
    constexpr int XU_SIZE = 164;                   // X[k*3+d] (81), u at +81, +2 pad
    constexpr int STAGE_LSTRIDE = SUB * 9 + 1;     // per lane (37 doubles = 10 banks: conflict free)
    constexpr int STAGE_SIZE = 1000;               // 27 * 37 = 999 per warp
    constexpr int POS_STRIDE = 732;                // 729 map entries per element padded to 16 bytes
    constexpr int PAIR_DOUBLES = G_SIZE + 1 + MP_SIZE + XU_SIZE + WPE * STAGE_SIZE;  // +1: 16 B align
    constexpr int PAIR_BYTES = PAIR_DOUBLES * 8 + 32 * 16 + POS_STRIDE * 4;
    static_assert((G_SIZE + 1) % 2 == 0 && PAIR_BYTES % 16 == 0, "16-byte aligned carve-up");

    // small integer tables (shared memory): node -> (a,b,c) parameter-line indices 0..2 packed
    // a | b<<2 | c<<4, and the inverse a+3b+9c -> node
    // (node order: 4C_fem_general_utils_local_connectivity_matrices.hpp:291)
    struct IntTables
    {
      int32_t abc[32];
      int32_t node_of[32];
    };

    struct PairSmem
    {
      double* g;      // G_SIZE
      double* mp;     // MP_SIZE (16-byte aligned)
      double* xu;     // XU_SIZE
      double* stage;  // WPE * STAGE_SIZE; warp w uses stage + w*STAGE_SIZE
      NodeRow* nrow;  // 32 (27 used)
      int32_t* pos;   // POS_STRIDE
    };

    B2_HD PairSmem carve(unsigned char* base)
    {
      PairSmem S;
      S.g = reinterpret_cast<double*>(base);
      S.mp = S.g + G_SIZE + 1;
      S.xu = S.mp + MP_SIZE;
      S.stage = S.xu + XU_SIZE;
      S.nrow = reinterpret_cast<NodeRow*>(S.stage + WPE * STAGE_SIZE);
      S.pos = reinterpret_cast<int32_t*>(S.nrow + 32);
      return S;
    }

    B2_HD int mp_index(int pair, int g) { return (pair * MP_SLOTS + g) * 2; }
    B2_HD double mp_value(const PairSmem& S, int g, int idx)
    {
      const d2 t = *reinterpret_cast<const d2*>(S.mp + mp_index(idx >> 1, g));
      return (idx & 1) ? t.y : t.x;
    }

    // reference coordinates of node q in {-1,0,1}^3 from the packed table
    B2_HD void gather_pad(const PairSmem& S, const IntTables& T, int q)
    {
      const int abc = T.abc[q];
#pragma unroll
      for (int d = 0; d < 3; ++d)
      {
        S.xu[3 * q + d] = (double)(((abc >> (2 * d)) & 3) - 1);
        S.xu[NDOF + 3 * q + d] = 0.0;
      }
    }

    // synchronous gather of node q (host emulation of the cp.async pipeline)
    B2_HD void gather(const EvalArgs& A, const PairSmem& S, const IntTables& T, int q, int64_t e)
    {
      if (e < A.n_ele)
      {
        const int64_t node = A.conn[e * NN + q];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          S.xu[3 * q + d] = A.X[3 * node + d];
          S.xu[NDOF + 3 * q + d] = A.u[3 * node + d];
        }
        S.nrow[q] = A.node_row[node];
        for (int j = 0; j < NN; ++j) S.pos[q * NN + j] = A.pos[e * POS_STRIDE + q * NN + j];
      }
      else
      {
        gather_pad(S, T, q);
        S.nrow[q].rowstart = -1, S.nrow[q].rowlen = 0, S.nrow[q].frow = -1;
      }
    }

    // ---- nodal Jacobian check at node n (calc_lib.hpp:543-562) ------------------------------------
    // dN_k/dxi_a at a node is dL_{a_k}(x_a) * delta * delta: only the three nodes on the parameter line
    // through n in direction a contribute, with dL_0(x) = x - 1/2, dL_1(x) = -2x, dL_2(x) = x + 1/2
    // (shape functions ...shapefunctions.hpp:706-830); the reference's sum over 27 nodes has 24 exact
    // zeros per entry.
    B2_HD double node_jacobian_det(const PairSmem& S, const IntTables& T, int n)
    {
      const int abc = T.abc[n];
      const int idx[3] = {abc & 3, (abc >> 2) & 3, (abc >> 4) & 3};
      const int strd[3] = {1, 3, 9};
      const int lin = idx[0] + 3 * idx[1] + 9 * idx[2];
      double J[9];
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        const double x = (double)(idx[a] - 1);
        const double dl[3] = {x - 0.5, -2.0 * x, x + 0.5};
        const int base = lin - idx[a] * strd[a];
        double r[3] = {0, 0, 0};
#pragma unroll
        for (int t = 0; t < 3; ++t)
        {
          const double* X = S.xu + 3 * T.node_of[base + t * strd[a]];
#pragma unroll
          for (int b = 0; b < 3; ++b) r[b] = fma(dl[t], X[b], r[b]);
        }
        J[a] = r[0], J[a + 3] = r[1], J[a + 6] = r[2];
      }
      return det3(J);
    }

    // ---- phase 1: Gauss point g -----------------------------------------------------------------
    template <int MAT, bool STIFF, bool MULF = false, bool LINEAR = false>
    B2_HD void phase1(const EvalArgs& A, const PairSmem& S, const double* tab_gp, const double* wgt,
        int g, int64_t e, double* out54, double& detJ_gp, double& detF_gp)
    {
      const double* xu = S.xu;
      const double* dN = tab_gp + g * TAB_STRIDE;
      double J[9], iJ[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) J[i] = 0.0;
#pragma unroll 3
      for (int k = 0; k < NN; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(dN[3 * k + a], xu[3 * k + b], J[a + 3 * b]);
      const double detJ = inv3(iJ, J);
      detJ_gp = detJ;

      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double* Gout = S.g + g * G_GSTRIDE;
#pragma unroll 3
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
#pragma unroll
      for (int p = STIFF ? 0 : 22; p < MP_PAIRS; ++p)
      {
        double* dst = S.mp + mp_index(p, g);
        dst[0] = out54[2 * p];
        dst[1] = out54[2 * p + 1];
      }
    }

    // ---- phase 2: warp w, row node i; Kacc[kk][b+3c] = K_{i,(i+7w+kk)%27}(b,c) --------------------
    template <bool STIFF>
    B2_HD void phase2(const PairSmem& S, int w, int i, double (*Kacc)[9], double* facc)
    {
#pragma unroll
      for (int k = 0; k < NBLK; ++k)
#pragma unroll
        for (int q = 0; q < 9; ++q) Kacc[k][q] = 0.0;
      facc[0] = facc[1] = facc[2] = 0.0;
      if (!STIFF && w != 0) return;

#pragma unroll 1
      for (int g = 0; g < NGP; ++g)
      {
        const double* Gg = S.g + g * G_GSTRIDE;
        const double Gi[3] = {Gg[3 * i], Gg[3 * i + 1], Gg[3 * i + 2]};
        if (w == 0)
        {
          double Pw[9];
#pragma unroll
          for (int r = 0; r < 9; ++r) Pw[r] = mp_value(S, g, 45 + r);
#pragma unroll
          for (int a = 0; a < 3; ++a)
            facc[a] = fma(Pw[a + 6], Gi[2], fma(Pw[a + 3], Gi[1], fma(Pw[a], Gi[0], facc[a])));
        }
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
              const double m = mp_value(S, g, m45(p, q));
              out[p % 3][q] = fma(Gi[p / 3], m, out[p % 3][q]);
              if (p != q) out[q % 3][p] = fma(Gi[q / 3], m, out[q % 3][p]);
            }
#pragma unroll
          for (int kk = 0; kk < NBLK; ++kk)
          {
            int j = i + NBLK * w + kk;
            j = j >= NN ? j - NN : j;
            const double Gj[3] = {Gg[3 * j], Gg[3 * j + 1], Gg[3 * j + 2]};
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                Kacc[kk][b + 3 * c] = fma(out[b][6 + c], Gj[2],
                    fma(out[b][3 + c], Gj[1], fma(out[b][c], Gj[0], Kacc[kk][b + 3 * c])));
          }
        }
      }
    }

    // ---- stage sub-pass s (blocks kk = 4s .. min(4s+3, 6)) of lane i into the warp's buffer ------
    B2_HD void stage_blocks(double* stage, int i, int s, const double (*Kacc)[9])
    {
      double* dst = stage + i * STAGE_LSTRIDE;
#pragma unroll
      for (int t = 0; t < SUB; ++t)
      {
        const int kk = SUB * s + t;
        if (kk < NBLK)
        {
#pragma unroll
          for (int q = 0; q < 9; ++q) dst[9 * t + q] = Kacc[kk < NBLK ? kk : 0][q];
        }
      }
    }

    // ---- scatter of sub-pass s: lane l < 12 (9 in the second sub-pass) = (block t, column dof c) --
    // direct:      K_{i, i+k}(b,c)   = stage[i][t][b+3c]      -> row (i,b), columns of node i+k
    // transposed:  K_{i, i-k}(b,c)   = stage[i-k][t][c+3b]    -> row (i,b), columns of node i-k  (k > 0)
    // lanes 0..11 direct, lanes 16..27 transposed: one RED instruction = one CSR row, 24 entries.
    // PAD: hex20 blocks through this kernel (element nodes >= 20 are padding, map entries -1); a compile-time flag, because
    // the run-time test of the map entry alone costs the hex27 kernel 1.8 % (profiles/r02h_hex27_ab.log)
    template <bool PAD = false>
    B2_HD void scatter_blocks(const EvalArgs& A, const PairSmem& S, const double* stage, int w, int s,
        int lane)
    {
      const bool tr = lane >= 16;
      const int l = lane & 15;
      const int t = l / 3, c = l - 3 * t;
      const int kk = SUB * s + t;
      const int k = NBLK * w + kk;
      const bool on = t < SUB && kk < NBLK && !(tr && k == 0);
      // three row nodes per step: their shared-memory reads (row info, map entry, staged block entries) first, then
      // the REDs -- no branch between the loads, so their latencies overlap
#pragma unroll 1
      for (int i0 = 0; i0 < NN; i0 += 3)
      {
        int64_t off[3];
        int32_t len[3];
        double v[3][3];
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
          const int i = i0 + q;
          const NodeRow r = S.nrow[i];
          int j = tr ? i - k : i + k;
          j = j < 0 ? j + NN : (j >= NN ? j - NN : j);
          const int32_t pj = S.pos[i * NN + (on ? j : 0)];  // < 0: padding column (hex20 blocks through this kernel)
          const bool act = on && r.rowstart >= 0 && (!PAD || pj >= 0);  // the row-ownership part is warp uniform
          off[q] = act ? r.rowstart + pj + c : -1;
          len[q] = r.rowlen;
          const double* src = tr ? stage + j * STAGE_LSTRIDE + 9 * t + c : stage + i * STAGE_LSTRIDE + 9 * t + 3 * c;
          const int sb = tr ? 3 : 1;  // stride over the row dof b inside the staged block
          v[q][0] = on ? src[0] : 0.0, v[q][1] = on ? src[sb] : 0.0, v[q][2] = on ? src[2 * sb] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q)
          if (off[q] >= 0)
          {
            double* dst = A.vals + off[q];
            B2_RED_ADD(dst, v[q][0]);
            B2_RED_ADD(dst + len[q], v[q][1]);
            B2_RED_ADD(dst + 2 * (int64_t)len[q], v[q][2]);
          }
      }
    }

    // f_i of row node i (warp 0): three REDs per lane
    B2_HD void scatter_force(const EvalArgs& A, const PairSmem& S, int i, const double* facc)
    {
      const int32_t fr = S.nrow[i].frow;
      if (fr < 0) return;
      B2_RED_ADD(A.f + fr, facc[0]);
      B2_RED_ADD(A.f + fr + 1, facc[1]);
      B2_RED_ADD(A.f + fr + 2, facc[2]);
    }
  }  // namespace hex27
}  // namespace b200s

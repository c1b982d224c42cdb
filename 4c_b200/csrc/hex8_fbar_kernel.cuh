// hex8_fbar_kernel.cuh -- evaluate + assemble for hex8 solid elements with the F-bar element
// technology (TECH fbar): SolidEleCalc<hex8, FBarFormulation<hex8>>
//   src/solid_ele/4C_solid_ele_calc_fbar.hpp:58-180      (prepare / evaluate / add_* of the formulation)
//   src/solid_ele/4C_solid_ele_calc_lib_fbar.hpp:47-160  (fbar factor, H-operator, extra stiffness)
//
// Same warp mapping and gather / stage / scatter as hex8_kernel.cuh (one warp = 4 elements x 8
// lanes); what changes:
//   phase 1  every lane also evaluates the centroid mapping (J0, dN0/dX, F0, det F0, F0^-1; exact
//            corner-sign form of the xi = 0 derivatives) -- 8-fold redundant, ~300 flop, cheaper than
//            a broadcast -- and node q = lane%8 stores h0_q(c) = sum_d F0^-1(d,c) dN0_q/dX_d.
//            At the Gauss point: fbar = (det F0 / det F)^(1/3), material at Cbar = fbar^2 C,
//            S scaled by w/fbar (force, geometric stiffness), Cm by w fbar (elastic stiffness),
//            and the two extra operators of the F-bar tangent
//               Q    = F ((Cm:C) - S)/3       (3x3; bopccg + bops of the reference = Q dN_i/dX)
//               Finv = F^-1
//            so that the reference's  K(idof,jdof) += Hop(jdof) (bops + bopccg)(idof)  reads
//               K_ij(b,c) += v_i(b) (h0_j(c) - sum_d Finv(d,c) dN_j/dX_d),   v_i = Q dN_i/dX.
//   phase 2  the F-bar tangent is NOT symmetric: lane i accumulates the whole block row i (8 blocks,
//            72 doubles).  The -v_i Finv part is folded into out_i = dN_i/dX . M before the block
//            products, the h0 part is separable: V_i = sum_gp v_i is accumulated like f_i and applied
//            once after the Gauss loop.
//   stage    plain row-major copy of the 8 blocks (no transposes).
// Shared memory: 9 more operator pairs per Gauss point (Q, Finv) and h0 (24 doubles per element):
// 31,040 B per warp; the kernel runs 7 warps in one CTA per SM (216.6 KB).
#pragma once

#include "hex8_kernel.cuh"

namespace b200s
{
  namespace hex8fbar
  {
    using hex8::EPW;
    using hex8::NGP;
    using hex8::NN;
    constexpr int MP_PAIRS = 36;  // 45 M + 9 Pw + 9 Q + 9 Finv = 72 doubles = 36 x double2
    constexpr int MP_SIZE = MP_PAIRS * hex8::MP_SLOTS * 2;
    constexpr int H0_SIZE = EPW * 24;
    constexpr int WARP_DOUBLES = MP_SIZE + hex8::G_SIZE + hex8::XU_SIZE + H0_SIZE;
    static_assert(hex8::FE_OFFSET + EPW * 24 <= MP_SIZE + hex8::G_SIZE, "K_e/f_e staging aliases M/G");
    constexpr int WARP_BYTES = WARP_DOUBLES * 8 + EPW * NN * 16 + EPW * NN * NN * 4;
    constexpr int IDX_Q = 54, IDX_FINV = 63;  // value indices behind M (0..44) and Pw (45..53)

    // parametric sign of corner k in direction a (hex8 node order)
    B2_HD double sgn(int k, int a)
    {
      return a == 0 ? (((k ^ (k >> 1)) & 1) ? 1.0 : -1.0)
                    : (a == 1 ? ((k & 2) ? 1.0 : -1.0) : ((k & 4) ? 1.0 : -1.0));
    }

    // ---- phase 1 ----------------------------------------------------------------------------------
    // out54: M (45) and Pw (9) for hex8::phase1_store; Q, Finv and h0 are written to shared memory
    // here (the operator area is free: the previous tile's scatter ended with a __syncwarp).
    template <int MAT, bool STIFF>
    B2_HD void phase1(const EvalArgs& A, const hex8::WarpSmem& S, const double* tab_gp,
        const double* wgt, int lane, int64_t e0, double* out54, double& detJ_min, double& detF_min,
        bool& neg_jac)
    {
      const int el = lane >> 3, g = lane & 7;
      const int64_t e = e0 + el;
      const double* xu = S.xu + el * hex8::XU_ESTRIDE;
      double J[9], iJ[9];

      // nodal Jacobian check at node g (calc_lib.hpp:543-562), see hex8::phase1
      {
        const double sr = 0.5 * sgn(g, 0), ss = 0.5 * sgn(g, 1), st = 0.5 * sgn(g, 2);
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

      // FBarFormulation::prepare (fbar.hpp:58-66): centroid mapping; dN_k/dxi_a(0) = sgn(k,a)/8
#pragma unroll
      for (int i = 0; i < 9; ++i) J[i] = 0.0;
#pragma unroll
      for (int k = 0; k < NN; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) J[a + 3 * b] += (0.125 * sgn(k, a)) * xu[3 * k + b];
      inv3(iJ, J);
      double F0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double G0g[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        double Gk[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
          Gk[a] = iJ[a] * (0.125 * sgn(k, 0)) + iJ[a + 3] * (0.125 * sgn(k, 1)) +
                  iJ[a + 6] * (0.125 * sgn(k, 2));
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          const double ua = xu[24 + 3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F0[a + 3 * b] += ua * Gk[b];
          if (k == g) G0g[a] = Gk[a];
        }
      }
      double iF0[9];
      const double detF0 = inv3(iF0, F0);
      if (STIFF)
      {
        // centroid half of the H-operator for node g (lib_fbar.hpp:94-97)
        double* h0 = S.h0 + el * 24 + 3 * g;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          h0[c] = iF0[3 * c] * G0g[0] + iF0[1 + 3 * c] * G0g[1] + iF0[2 + 3 * c] * G0g[2];
      }

      // Jacobian mapping and deformation gradient at Gauss point g
      const double* dN = tab_gp + g * hex8::TAB_STRIDE;
      hex8::jacobian(dN, xu, J);
      const double detJ = inv3(iJ, J);
      neg_jac = !(detJ_node > 0.0) && (e < A.n_ele);
      detJ_min = fmin(detJ_node, detJ);
      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double* Gout = S.g + el * hex8::G_ESTRIDE + g * hex8::G_GSTRIDE;
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
      double iF[9];
      const double detF = inv3(iF, F);
      detF_min = detF;

      // evaluate_fbar_factor (lib_fbar.hpp:47-51) and the strains of F_bar = fbar F
      const double fbar = pow(detF0 / detF, 1.0 / 3.0);
      double C[6], Cbar[6], Sv[6], Cm[21];
      right_cauchy_green(F, C);
      {
        double Fb[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Fb[i] = fbar * F[i];
        right_cauchy_green(Fb, Cbar);
      }
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr && e < A.n_ele) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr && e < A.n_ele) mp.fiber = A.fibers + 3 * e;
      material_evaluate<MAT>(mp, Cbar, Sv, Cm);

      // force / geometric stiffness carry w/fbar, the elastic stiffness w*fbar (fbar.hpp:152-176)
      const double w = detJ * wgt[g];
      const double ws = w / fbar, wc = w * fbar;
#pragma unroll
      for (int v = 0; v < 6; ++v) Sv[v] *= ws;
      if (STIFF)
      {
#pragma unroll
        for (int v = 0; v < 21; ++v) Cm[v] *= wc;
        material_geometric_operator(F, Sv, Cm, out54);
        // add_fbar_stiffness_matrix (lib_fbar.hpp:121-160): A = ((Cm : C) - S)/3 with the factors
        // w fbar/3 and -w/fbar/3 already inside Cm and Sv;  Q = F A
        double Av[6];
#pragma unroll
        for (int I = 0; I < 6; ++I)
        {
          double v = 0.0;
#pragma unroll
          for (int Jv = 0; Jv < 6; ++Jv) v += Cm[c21(I, Jv)] * C[Jv] * (Jv < 3 ? 1.0 : 2.0);
          Av[I] = (v - Sv[I]) * (1.0 / 3.0);
        }
        double Q[9];
        first_piola(F, Av, Q);
        const int slot = el * 9 + g;
#pragma unroll
        for (int r = 0; r < 9; ++r)
        {
          S.mp[((IDX_Q + r) / 2 * hex8::MP_SLOTS + slot) * 2 + ((IDX_Q + r) & 1)] = Q[r];
          S.mp[((IDX_FINV + r) / 2 * hex8::MP_SLOTS + slot) * 2 + ((IDX_FINV + r) & 1)] = iF[r];
        }
      }
      first_piola(F, Sv, out54 + 45);
    }

    // ---- phase 2: lane = row node i, full block row -------------------------------------------
    // Kacc[k][b + 3c] = K_ij(b,c) with j = (i + k) mod 8 (rotated like the displacement-based kernel: the
    // staging stores of the eight lanes of an element then fall on distinct banks), facc = f_i
    template <bool STIFF>
    B2_HD void phase2(const hex8::WarpSmem& S, int lane, double (*Kacc)[9], double* facc)
    {
      const int el = lane >> 3, i = lane & 7;
#pragma unroll
      for (int j = 0; j < NN; ++j)
#pragma unroll
        for (int q = 0; q < 9; ++q) Kacc[j][q] = 0.0;
      facc[0] = facc[1] = facc[2] = 0.0;
      double Vacc[3] = {0.0, 0.0, 0.0};

#pragma unroll 1
      for (int g = 0; g < NGP; ++g)
      {
        const double* Gg = S.g + el * hex8::G_ESTRIDE + g * hex8::G_GSTRIDE;
        const double Gi[3] = {Gg[3 * i], Gg[3 * i + 1], Gg[3 * i + 2]};
        double Pw[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) Pw[r] = hex8::mp_value(S, el, g, 45 + r);
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
              const double m = hex8::mp_value(S, el, g, m45(p, q));
              out[p % 3][q] = fma(Gi[p / 3], m, out[p % 3][q]);
              if (p != q) out[q % 3][p] = fma(Gi[q / 3], m, out[q % 3][p]);
            }
          // v_i = Q dN_i/dX  (= bops + bopccg of node i), V_i += v_i
          double v[3];
#pragma unroll
          for (int b = 0; b < 3; ++b)
          {
            v[b] = fma(hex8::mp_value(S, el, g, IDX_Q + b + 6), Gi[2],
                fma(hex8::mp_value(S, el, g, IDX_Q + b + 3), Gi[1],
                    hex8::mp_value(S, el, g, IDX_Q + b) * Gi[0]));
            Vacc[b] += v[b];
          }
          // out[b][3d+c] -= v_i(b) Finv(d,c): the Gauss-point half of the H-operator
#pragma unroll
          for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int c = 0; c < 3; ++c)
            {
              const double fi = hex8::mp_value(S, el, g, IDX_FINV + d + 3 * c);
#pragma unroll
              for (int b = 0; b < 3; ++b) out[b][3 * d + c] = fma(-v[b], fi, out[b][3 * d + c]);
            }
#pragma unroll
          for (int k = 0; k < NN; ++k)
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
      if (STIFF)
      {
        // centroid half: K_ij(b,c) += V_i(b) h0_j(c)
        const double* h0 = S.h0 + el * 24;
#pragma unroll
        for (int k = 0; k < NN; ++k)
#pragma unroll
          for (int c = 0; c < 3; ++c)
          {
            const double h = h0[3 * ((i + k) & 7) + c];
#pragma unroll
            for (int b = 0; b < 3; ++b) Kacc[k][b + 3 * c] = fma(Vacc[b], h, Kacc[k][b + 3 * c]);
          }
      }
    }

    // ---- stage: block row i -> rows 3i..3i+2 of the row-major K_e ---------------------------------
    template <bool STIFF>
    B2_HD void stage_ke(const hex8::WarpSmem& S, int lane, const double (*Kacc)[9], const double* facc)
    {
      const int el = lane >> 3, i = lane & 7;
      double* fe = S.mp + hex8::FE_OFFSET + el * 24;
      fe[3 * i] = facc[0], fe[3 * i + 1] = facc[1], fe[3 * i + 2] = facc[2];
      if (!STIFF) return;
      double* Ke = S.mp + el * hex8::KE_ESTRIDE;
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        const int j = (i + k) & 7;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) Ke[(3 * i + b) * 24 + 3 * j + c] = Kacc[k][b + 3 * c];
      }
    }
  }  // namespace hex8fbar
}  // namespace b200s

// hex8_eas_mma_kernel.cuh -- hex8 EAS evaluate + assemble with the element matrices formed on the FP64 tensor
// cores (mma.sync m8n8k4, "DMMA"); same reference functions and the same element history as hex8_eas_kernel.cuh
// (read that header first; src/solid_ele/4C_solid_ele_calc_eas.hpp:44-290, 4C_solid_ele_calc_eas_helpers.hpp).
//
// Why: everything an EAS element integrates is one symmetric product.  With, per Gauss point g,
//     P_g = [ B_g | Mtilde_g ]            6 x (24 + NEAS)          (strain-like rows)
//     Q_g = [ w C_g P_g | w S_g ]         6 x (24 + NEAS + 1)
// stacked over the Gauss points (k = 6 g + I, 48 rows), the matrix  S = P^T Q  holds
//     S[0:24, 0:24]   = sum_g B^T  w C B   = material part of K_dd          (integrate_eas / add_elastic_stiffness)
//     S[0:24, 24: ]   = sum_g B^T  w C Mt  = K_da
//     S[24: , 24: ]   = sum_g Mt^T w C Mt  = K_aa
//     S[0:24, last]   = sum_g B^T  w S     = f_int,      S[24:, last] = sum_g Mt^T w S = s
// i.e. a (24+NEAS+1)^2 x 48 GEMM of which only the upper 8 x 8 tiles are needed (15 tiles for eas_mild, 21 for
// eas_full).  The warp-per-element FMA kernel spends its time on broadcast shared-memory loads (one 128-bit load per
// two FMAs) with 24 or 21 of 32 lanes busy; on the tensor path the operands are 8 x 4 / 4 x 8 register fragments
// (one 64-bit shared load per lane and tile feeds 256 FMAs), all 32 lanes work, and the accumulators (2 doubles per
// tile and lane) never leave the register file.  The condensation  mK = -K_da K_aa^-1  and  K_dd += mK K_da^T  are
// two more small GEMMs on the same path, the geometric stiffness  kg(i,j) = sum_g G_i . (w S_g) . G_j  one 8x8x24
// product.  K_aa^-1 stays on the FMA pipe: Gauss-Jordan without pivoting, lane = row, rows in registers (see
// warp_spd_inverse_rot).
//
// Work per element (one warp):
//   gather: cp.async into a double-buffered stage, one element ahead (issued by the kernel around element_mma)
//   phase 1 (lane g < 8 = Gauss point g)                      as hex8_eas_kernel.cuh
//   4 x { build P, Q of two Gauss points (12 rows) in shared memory: per point 30 lanes push the 24 columns of B and
//         the 6 columns of T0invT through w C (hoisted into registers), then the Mtilde columns are scaled copies;
//         3 k-steps: 2 NT fragment loads, NT (NT + 1) / 2 DMMA }
//   geometric 8 x 8 (6 DMMA), K_aa / K_da / s / f from the accumulators to shared memory (K_da also to the history),
//   K_aa^-1, history, mK (DMMA), f, K_dd += mK K_da^T (DMMA), stage K_e, scatter (as hex8::scatter).
//
// Fragment layout of mma.sync.aligned.m8n8k4.row.col.f64 (PTX ISA, "Matrix Fragments for mma.m8n8k4 with .f64"):
//   A (8 x 4): lane l holds A[l >> 2][l & 3];  B (4 x 8): lane l holds B[l & 3][l >> 2];
//   C, D (8 x 8): lane l holds C[l >> 2][2 (l & 3)] and C[l >> 2][2 (l & 3) + 1].
// Both operands of S = P^T Q are read as  M[k0 + (l & 3)][x0 + (l >> 2)]  from row-major [k][LD] arrays; with
// LD = 4 or 12 (mod 16) the 16 lanes of a half warp hit 16 different 8-byte banks.
#pragma once

#include "hex8_eas_kernel.cuh"

namespace b200s
{
  namespace hex8eas
  {
#if defined(__CUDACC__)
    __device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
    {
      asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
          : "+d"(c[0]), "+d"(c[1])
          : "d"(a), "d"(b));
    }

    // 1 / x without the slow path of the IEEE division: hardware estimate (relative error < 2^-20) and two Newton steps,
    // a result within an ulp or two; used for the pivots of K_aa^-1, whose operator is defined up to round-off anyway
    __device__ __forceinline__ double fast_rcp(double x)
    {
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
      r = fma(fma(-x, r, 1.0), r, r);
      return fma(fma(-x, r, 1.0), r, r);
    }

    // shared memory of one warp, in doubles
    template <int NEAS>
    struct MmaLayout
    {
      static constexpr int NA = (NEAS + 3) & ~3;
      static constexpr int NCU = ND + NEAS;               // columns of P: B | Mtilde
      static constexpr int CF = NCU;                      // column of Q that carries w S
      static constexpr int NT = (NCU + 1 + 7) / 8;        // 8-column tiles of S: 5 (mild), 6 (full)
      static constexpr int NCP = 8 * NT;
      static constexpr int LD = NEAS > 12 ? 52 : 44;      // P / Q row stride (4 / 12 mod 16)
      static constexpr int NK = (NEAS + 3) / 4;           // k-steps over the EAS parameters: 3 / 6
      static constexpr int NAT = (NEAS + 1 + 7) / 8;      // 8-column tiles of mK incl. the column -K_da K_aa^-1 s
      static constexpr int LDK = NEAS > 12 ? 28 : 20;     // K_da, mK, K_aa^-1 row stride (12 / 4 mod 16)
      static_assert(LD >= NCP && (LD % 16 == 12 || LD % 16 == 4), "P / Q row stride");
      static_assert(LDK >= 8 * NAT && LDK >= 4 * NK && (LDK % 16 == 12 || LDK % 16 == 4), "K_da / mK row stride");
      // gather stage (double buffered by the kernel, filled with cp.async one element ahead): X[24], u[24], T0invT 6x6
      // column-major + detJ0 (40), alpha[NA], NodeRow[8], pos[64], conn[8]
      static constexpr int ST_XU = 0, ST_T0 = 48, ST_ALPHA = 88, ST_NROW = ST_ALPHA + NA, ST_POS = ST_NROW + 2 * NN;
      static constexpr int ST_CONN = ST_POS + NN * NN / 2;  // node ids of the element after the next one (kernel)
      static constexpr int STAGE = ST_CONN + NN / 2;       // doubles per stage
      static constexpr int SG = 2 * STAGE;                // s[NA]
      static constexpr int FE = SG + NA;                  // f_e[24]
      static constexpr int BUF = FE + ND;                 // pivot row exchange of the inverse, 2 x NA
      static constexpr int CT = BUF + 2 * NA;             // w C T0invT of the two Gauss points of a pair: [gg][q][I]
      // Gauss-loop buffers: per Gauss point w C[21] (16-byte aligned), F[9], w S[6], G[24], scale; then P, Q of a pair
      static constexpr int GP = CT + 72;
      static constexpr int GP_STRIDE = 62;
      static constexpr int GP_C = 0, GP_F = 22, GP_S = 31, GP_G = 37, GP_SCALE = 61;
      static constexpr int PM = GP + NGP * GP_STRIDE;
      static constexpr int QM = PM + 12 * LD;
      static constexpr int LOOP_END = QM + 12 * LD;
      static_assert(GP % 2 == 0 && GP_STRIDE % 2 == 0, "w C is read with 128-bit loads");
      // after the loop (GP / P / Q dead): K_aa, then [K_aa^-1 | K_aa^-1 s] (4 NK rows), then mK (24 rows); K_da behind;
      // K_e (24 x 26) is staged over the front once both are consumed
      static constexpr int AUG = GP;
      static constexpr int MK = AUG;
      static constexpr int KDA = AUG + ND * LDK;
      static constexpr int KE = GP, KE_RS = 26;
      static_assert(4 * NK <= ND && KDA + ND * LDK <= LOOP_END && KE + ND * KE_RS <= LOOP_END,
          "post-loop work areas fit into the dead Gauss-loop buffers");
      static constexpr int DOUBLES = LOOP_END;
      static constexpr int BYTES = DOUBLES * 8;
    };

    // In-place inverse of a symmetric positive definite N x N matrix, lane r holds row r in a[] (lanes >= N pass
    // zeros).  Gauss-Jordan without pivoting like warp_spd_inverse_t, but the rows live in registers although the step
    // loop stays rolled: every step rotates the row by one place (the update of entry k is written to place k - 1, the
    // new pivot-column entry goes to the last place), so that the pivot column always sits in a[0] and no register
    // is indexed dynamically; after N steps the rotation is the identity.  The pivot lane publishes its row through the
    // double-buffered scratch `buf` (2 N doubles): one __syncwarp per step.
    template <int N>
    __device__ __forceinline__ void warp_spd_inverse_rot(double (&a)[N], double* buf, int lane)
    {
      double scale = 1.0;  // the row of a lane that has been the pivot is kept unscaled: true row = scale * a[]
#pragma unroll 1
      for (int c = 0; c < N; ++c)
      {
        double* b = buf + (c & 1) * N;
        if (lane == c)
        {
#pragma unroll
          for (int k = 0; k < N; ++k) b[k] = a[k];
        }
        __syncwarp();
        // a(r, k) -= (a(r, c) / a(c, c)) a(c, k) is linear in row r, so it acts on the unscaled rows as it stands; the
        // pivot lane passes f = 0, rotates, and from now on carries the factor 1 / a(c, c)
        const bool piv = lane == c;
        const double d = fast_rcp(b[0]);
        const double f = piv ? 0.0 : a[0] * d;
#pragma unroll
        for (int k = 1; k < N; ++k) a[k - 1] = fma(-f, b[k], a[k]);
        a[N - 1] = piv ? 1.0 : -f;
        scale = piv ? d : scale;
      }
#pragma unroll
      for (int k = 0; k < N; ++k) a[k] *= scale;
      __syncwarp();
    }

    // phase 1 of both EAS kernels: lane g < 8 = Gauss point g (and element node g for the Jacobian check); leaves
    // F, w S, w C, dN/dX and detJ0 / detJ in the Gauss point's block of shared memory (layout LT)
    template <int MAT, int NEAS, class LT>
    __device__ __forceinline__ void gauss_points(const EvalArgs& A, double* sm, const double* xu, const double* T0,
        const double* alpha, const double* tab_gp, const double* wgt, const double* gxi, int lane, int64_t e, double detJ0,
        double& detJ_min, double& detF_min, bool& neg_jac)
    {
      if (lane >= NGP) return;
      const int g = lane;
      double J[9], iJ[9];
      {
        // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562), corner form
        const double sr = 0.5 * hex8fbar::sgn(g, 0), ss = 0.5 * hex8fbar::sgn(g, 1), st = 0.5 * hex8fbar::sgn(g, 2);
        const double *xq = xu + 3 * g, *xr = xu + 3 * (g ^ 1), *xs = xu + 3 * (g ^ 3), *xt = xu + 3 * (g ^ 4);
#pragma unroll
        for (int b = 0; b < 3; ++b)
        {
          J[0 + 3 * b] = sr * (xq[b] - xr[b]);
          J[1 + 3 * b] = ss * (xq[b] - xs[b]);
          J[2 + 3 * b] = st * (xq[b] - xt[b]);
        }
        const double dn = det3(J);
        neg_jac = !(dn > 0.0);
        detJ_min = fmin(detJ_min, dn);
      }
      const double* dN = tab_gp + g * hex8::TAB_STRIDE;
      hex8::jacobian(dN, xu, J);
      const double detJ = inv3(iJ, J);
      detJ_min = fmin(detJ_min, detJ);
      double* gp = sm + LT::GP + g * LT::GP_STRIDE;
      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      double G[24];  // kept in registers until X, u and the shape function table are consumed: no load of this loop has to
                     // be ordered behind a store it might alias with
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
#pragma unroll
        for (int a = 0; a < 3; ++a)
          G[3 * k + a] = iJ[a] * dN[3 * k] + iJ[a + 3] * dN[3 * k + 1] + iJ[a + 6] * dN[3 * k + 2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          const double ua = xu[24 + 3 * k + a];
#pragma unroll
          for (int b = 0; b < 3; ++b) F[a + 3 * b] += ua * G[3 * k + b];
        }
      }
#pragma unroll
      for (int i = 0; i < 24; ++i) gp[LT::GP_G + i] = G[i];
      detF_min = fmin(detF_min, det3(F));
      // E_enh = E + stress-like(Mtilde alpha), Mtilde = detJ0/detJ T0invT M(xi)
      // (evaluate_eas_kinematics, eas_helpers.hpp:405-448); C_enh = C + 2 (...), shear rows halved
      const double xr = gxi[3 * g], xs = gxi[3 * g + 1], xt = gxi[3 * g + 2];
      const double scale = detJ0 / detJ;
      double C[6];
      right_cauchy_green(F, C);
      // KINEM linear (EASFormulation<..., linear>, eas_helpers.hpp:648-666): linear strain, B-operator of F = I; the
      // geometric term stays, as in the reference's add_stiffness_matrix (eas.hpp:142-157)
      if (A.linear) linear_kinematics(F, C);
      {
        double v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < NEAS; ++c)
        {
          constexpr int rows[21] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 3, 3, 4, 4, 5, 5, 0, 0, 1, 1, 2, 2};
          constexpr int mono[21] = {0, 1, 2, 0, 1, 1, 2, 0, 2, 4, 5, 3, 4, 3, 5, 3, 4, 3, 5, 4, 5};
          v[rows[c]] = fma(monomial(mono[c], xr, xs, xt), alpha[c], v[rows[c]]);
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
        {
          double t = 0.0;
#pragma unroll
          for (int q = 0; q < 6; ++q) t = fma(T0[k + 6 * q], v[q], t);
          C[k] += (k < 3 ? 2.0 : 1.0) * scale * t;
        }
      }
      double Sv[6], Cm[21];
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr)
        mp.hist = A.mat_hist + (e * 8 + g) * material_hist_doubles(MAT);
      material_evaluate<MAT>(mp, C, Sv, Cm);
      const double w = detJ * wgt[g];
#pragma unroll
      for (int i = 0; i < 9; ++i) gp[LT::GP_F + i] = F[i];
#pragma unroll
      for (int i = 0; i < 6; ++i) gp[LT::GP_S + i] = w * Sv[i];
#pragma unroll
      for (int i = 0; i < 21; ++i) gp[LT::GP_C + i] = w * Cm[i];
      gp[LT::GP_SCALE] = scale;
    }

    // The warp's work on one element whose gather stage (X, u, T0invT, detJ0, alpha, row info, scatter map) is in shared
    // memory; sm: the warp's work area (MmaLayout<NEAS>), hist: the element's history block in HBM.
    template <int MAT, bool STIFF, int NEAS>
    __device__ __forceinline__ void element_mma(const EvalArgs& A, double* sm, const double* stage,
        const double* tab_gp /*[g][k*3+a], stride TAB_STRIDE*/, const double* wgt, const double* gxi /*[g][3]*/,
        const int (&m_row)[(NEAS + 1 + 15) / 16], const int (&m_mono)[(NEAS + 1 + 15) / 16], int lane, int64_t e,
        double* hist, double& detJ_min, double& detF_min, bool& neg_jac)
    {
      // m_row[r], m_mono[r]: Mtilde column 16 r + (lane & 15) of this lane -- Voigt row (-1: the w S column, -2: idle) and
      // monomial of its non-zero; gxi[24 + 8 m + g]: monomial m at Gauss point g
      using L = MmaLayout<NEAS>;
      constexpr int NT = L::NT, LD = L::LD, LDK = L::LDK, NK = L::NK, NAT = L::NAT;
      const double* xu = stage + L::ST_XU;
      const double* T0 = stage + L::ST_T0;
      const double* alpha = stage + L::ST_ALPHA;
      const NodeRow* nrow = reinterpret_cast<const NodeRow*>(stage + L::ST_NROW);
      const int32_t* pos = reinterpret_cast<const int32_t*>(stage + L::ST_POS);
      const int lq = lane >> 2, lr = lane & 3;  // fragment coordinates: "group" and "thread in group"

      gauss_points<MAT, NEAS, L>(A, sm, xu, T0, alpha, tab_gp, wgt, gxi, lane, e, T0[36], detJ_min, detF_min, neg_jac);
      __syncwarp();

      // ---- S = P^T Q over the Gauss points, two at a time ------------------------------------------------------
      double acc[NT][NT][2];
#pragma unroll
      for (int tm = 0; tm < NT; ++tm)
#pragma unroll
        for (int tn = 0; tn < NT; ++tn) acc[tm][tn][0] = acc[tm][tn][1] = 0.0;
      double* Pm = sm + L::PM;
      double* Qm = sm + L::QM;
      // column CF of P and the padding columns behind it in P and Q are zero for the whole element
      for (int t = lane; t < 12 * (L::NCP - L::CF); t += 32)
      {
        const int kk = t / (L::NCP - L::CF), c = L::CF + (t - kk * (L::NCP - L::CF));
        Pm[kk * LD + c] = 0.0;
        if (c > L::CF) Qm[kk * LD + c] = 0.0;
      }
      constexpr int MR = (NEAS + 1 + 15) / 16;
#pragma unroll 1
      for (int pr = 0; pr < NGP / 2; ++pr)
      {
        // per Gauss point of the pair one round of 30 lanes: the 24 columns of the B-operator and the 6 columns of
        // T0invT go through  q = (w C) p  (w C hoisted into registers: 11 loads, then six independent FMA chains)
        double* CT = sm + L::CT;
#pragma unroll 1
        for (int gg = 0; gg < 2; ++gg)
        {
          const int g = 2 * pr + gg;
          const double* gp = sm + L::GP + g * L::GP_STRIDE;
          if (lane < ND + 6)
          {
            double cw[22];
            {
              const d2* c2 = reinterpret_cast<const d2*>(gp + L::GP_C);
#pragma unroll
              for (int i = 0; i < 11; ++i)
              {
                const d2 t = c2[i];
                cw[2 * i] = t.x, cw[2 * i + 1] = t.y;
              }
            }
            double pc[6];
            if (lane < ND)
            {
              // column of the B-operator (evaluate_strain_gradient, calc_lib.hpp:777-826), strain-like rows
              const int ni = lane / 3, ei = lane - 3 * ni;
              const double* Gi = gp + L::GP_G + 3 * ni;
              const double Fe0 = gp[L::GP_F + ei], Fe1 = gp[L::GP_F + ei + 3], Fe2 = gp[L::GP_F + ei + 6];
              const double G0 = Gi[0], G1 = Gi[1], G2 = Gi[2];
              pc[0] = Fe0 * G0, pc[1] = Fe1 * G1, pc[2] = Fe2 * G2;
              pc[3] = Fe0 * G1 + Fe1 * G0;
              pc[4] = Fe1 * G2 + Fe2 * G1;
              pc[5] = Fe2 * G0 + Fe0 * G2;
            }
            else
            {
#pragma unroll
              for (int k = 0; k < 6; ++k) pc[k] = T0[k + 6 * (lane - ND)];
            }
            double qc[6];
#pragma unroll
            for (int I = 0; I < 6; ++I) qc[I] = cw[c21(I, 0)] * pc[0];
#pragma unroll
            for (int Jv = 1; Jv < 6; ++Jv)
#pragma unroll
              for (int I = 0; I < 6; ++I) qc[I] = fma(cw[c21(I, Jv)], pc[Jv], qc[I]);
            if (lane < ND)
            {
#pragma unroll
              for (int I = 0; I < 6; ++I)
              {
                Pm[(6 * gg + I) * LD + lane] = pc[I];
                Qm[(6 * gg + I) * LD + lane] = qc[I];
              }
            }
            else
            {
#pragma unroll
              for (int I = 0; I < 6; ++I) CT[(6 * gg + lane - ND) * 6 + I] = qc[I];
            }
          }
        }
        __syncwarp();
        // columns of Mtilde = detJ0 / detJ T0invT M(xi) (one non-zero per column of M: eas_helpers.hpp:205-275, 352-371):
        // p = mv T0invT(:, row), q = mv (w C T0invT)(:, row); one half warp per Gauss point of the pair, the lane behind
        // the last Mtilde column writes w S into column CF of Q
        {
          const int gg = lane >> 4, g = 2 * pr + gg;
          const double* gp = sm + L::GP + g * L::GP_STRIDE;
          const double sc = gp[L::GP_SCALE];
#pragma unroll
          for (int r = 0; r < MR; ++r)
          {
            double* pd = Pm + 6 * gg * LD + ND + 16 * r + (lane & 15);
            double* qd = Qm + 6 * gg * LD + ND + 16 * r + (lane & 15);
            if (m_row[r] >= 0)
            {
              const double mv = sc * gxi[24 + 8 * m_mono[r] + g];
              const double* t0c = T0 + 6 * m_row[r];
              const double* ctc = CT + (6 * gg + m_row[r]) * 6;
              double pv[6], qv[6];  // all loads before the first store (the compiler has to assume aliasing)
#pragma unroll
              for (int I = 0; I < 6; ++I) pv[I] = t0c[I], qv[I] = ctc[I];
#pragma unroll
              for (int I = 0; I < 6; ++I) pd[I * LD] = mv * pv[I], qd[I * LD] = mv * qv[I];
            }
            else if (m_row[r] == -1)
            {
              double qv[6];
#pragma unroll
              for (int I = 0; I < 6; ++I) qv[I] = gp[L::GP_S + I];
#pragma unroll
              for (int I = 0; I < 6; ++I) qd[I * LD] = qv[I];
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
        {
          const double* pk = Pm + (4 * ks + lr) * LD + lq;
          const double* qk = Qm + (4 * ks + lr) * LD + lq;
          double fa[NT], fb[NT];
#pragma unroll
          for (int t = 0; t < NT; ++t) fa[t] = pk[8 * t], fb[t] = qk[8 * t];
#pragma unroll
          for (int tm = 0; tm < NT; ++tm)
#pragma unroll
            for (int tn = tm; tn < NT; ++tn)
              if (STIFF || tn >= 3) dmma(acc[tm][tn], fa[tm], fb[tn]);
        }
        __syncwarp();  // P / Q are rebuilt by the next pair
      }

      // ---- geometric stiffness kg(i, j) = sum_g G_i . (w S_g) . G_j, an 8 x 8 x 24 product -----------------------
      double kg[2] = {0.0, 0.0};
      if (STIFF)
      {
#pragma unroll
        for (int ks = 0; ks < 6; ++ks)
        {
          const int k = 4 * ks + lr, g = k / 3, a = k - 3 * g;
          const double* gp = sm + L::GP + g * L::GP_STRIDE;
          const double* Gq = gp + L::GP_G + 3 * lq;
          const double av = Gq[a];  // A[i = lq][k] = G_g(i, a)
          // B[k][j = lq] = sum_d w S_g(a, d) G_g(j, d)
          const double bv = fma(gp[L::GP_S + vg(a, 2)], Gq[2], fma(gp[L::GP_S + vg(a, 1)], Gq[1], gp[L::GP_S + vg(a, 0)] * Gq[0]));
          dmma(kg, av, bv);
        }
      }
      __syncwarp();  // the Gauss-loop buffers are dead from here on

      // ---- K_aa, K_da, s, f out of the accumulators ----------------------------------------------------------------
      double* aug = sm + L::AUG;   // K_aa (symmetric, both halves written), row stride LDK
      double* KdaS = sm + L::KDA;  // K_da [r][a], row stride LDK, columns NEAS .. 4 NK - 1 zero
      double* sS = sm + L::SG;
      double* h_s = hist + 2 * NEAS;
      double* h_inv = hist + 3 * NEAS;
      double* h_kda = h_inv + NEAS * NEAS;
      double fr[3] = {0.0, 0.0, 0.0};  // f_int(8 tm + lq) on the lanes that hold column CF
      constexpr int CF_T = L::CF / 8, CF_O = L::CF % 8;
#pragma unroll
      for (int tm = 0; tm < NT; ++tm)
#pragma unroll
        for (int tn = (tm < 3 ? 3 : tm); tn < NT; ++tn)
#pragma unroll
          for (int i = 0; i < 2; ++i)
          {
            const int gr = 8 * tm + lq, gc = 8 * tn + 2 * lr + i;
            const double v = acc[tm][tn][i];
            const int ca = gc - ND;  // EAS column
            if (tm < 3)
            {
              if (ca < NEAS)
              {
                KdaS[gr * LDK + ca] = v;
                h_kda[gr + ND * ca] = v;  // history: column-major 24 x NEAS (eas.hpp:266-275)
              }
              else if (tn == CF_T && 2 * lr + i == CF_O)
                fr[tm] = v;
            }
            else
            {
              const int ra = gr - ND;  // EAS row
              if (ra < NEAS)
              {
                if (ca < NEAS)
                {
                  aug[ra * LDK + ca] = v;
                  if (tn > tm) aug[ca * LDK + ra] = v;
                }
                else if (gc == L::CF)
                  sS[ra] = v;
              }
            }
          }
      if (lane < ND)
#pragma unroll
        for (int c = NEAS; c < 4 * NK; ++c) KdaS[lane * LDK + c] = 0.0;
      __syncwarp();

      // ---- K_aa^-1: rows in registers ------------------------------------------------------------------------------------
      double row[NEAS];
#pragma unroll
      for (int k = 0; k < NEAS; ++k) row[k] = lane < NEAS ? aug[k * LDK + lane] : 0.0;
      __syncwarp();
      warp_spd_inverse_rot<NEAS>(row, sm + L::BUF, lane);
      // history for struct_calc_recover: s, K_aa^-1 (column-major)
      double vr = 0.0;  // (K_aa^-1 s)(lane)
      if (lane < NEAS)
      {
        h_s[lane] = sS[lane];
#pragma unroll
        for (int c = 0; c < NEAS; ++c)
        {
          h_inv[lane + NEAS * c] = row[c];
          vr = fma(row[c], sS[c], vr);
        }
      }
      // B operand of mK = -K_da [K_aa^-1 | K_aa^-1 s | 0]: aug[k][n], 4 NK rows x 8 NAT columns
      if (lane < 8 * NAT)
      {
#pragma unroll
        for (int k = 0; k < 4 * NK; ++k) aug[k * LDK + lane] = (k < NEAS && lane < NEAS) ? row[k < NEAS ? k : 0] : 0.0;
      }
      __syncwarp();
      if (lane < NEAS) aug[lane * LDK + NEAS] = vr;
      __syncwarp();

      // ---- mK = -K_da K_aa^-1 (prepare_condensed_contribution, eas.hpp:188-203), last used column: -K_da K_aa^-1 s ------
      double mk[3][NAT][2];
#pragma unroll
      for (int tm = 0; tm < 3; ++tm)
#pragma unroll
        for (int tn = 0; tn < NAT; ++tn) mk[tm][tn][0] = mk[tm][tn][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < NK; ++ks)
      {
        double fa[3], fb[NAT];
#pragma unroll
        for (int t = 0; t < 3; ++t) fa[t] = -KdaS[(8 * t + lq) * LDK + 4 * ks + lr];
#pragma unroll
        for (int t = 0; t < NAT; ++t) fb[t] = aug[(4 * ks + lr) * LDK + 8 * t + lq];
#pragma unroll
        for (int tm = 0; tm < 3; ++tm)
#pragma unroll
          for (int tn = 0; tn < NAT; ++tn) dmma(mk[tm][tn], fa[tm], fb[tn]);
      }
      // f_r = f_int + mK(r,:) s   (eas.hpp:245-253): both summands sit on the lanes with 2 lr + 1 == CF_O
      {
        constexpr int V_T = NEAS / 8, V_O = NEAS % 8;
        static_assert(V_O == CF_O && (V_O & 1) == 1, "f_int and its condensed part meet on the same lanes");
        if (2 * lr + 1 == CF_O)
#pragma unroll
          for (int tm = 0; tm < 3; ++tm) sm[L::FE + 8 * tm + lq] = fr[tm] + mk[tm][V_T][1];
      }

      if (STIFF)
      {
        // ---- K_dd += mK K_da^T  (eas.hpp:255-263): mK through shared memory into A fragments ---------------------------
        double* MkS = sm + L::MK;
        __syncwarp();  // [K_aa^-1 | K_aa^-1 s] consumed by every lane: mK is staged over it
#pragma unroll
        for (int tm = 0; tm < 3; ++tm)
#pragma unroll
          for (int tn = 0; tn < NAT; ++tn)
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (8 * tn + 2 * lr + i < 4 * NK) MkS[(8 * tm + lq) * LDK + 8 * tn + 2 * lr + i] = mk[tm][tn][i];
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < NK; ++ks)
        {
          double fa[3], fb[3];
#pragma unroll
          for (int t = 0; t < 3; ++t)
          {
            fa[t] = MkS[(8 * t + lq) * LDK + 4 * ks + lr];
            fb[t] = KdaS[(8 * t + lq) * LDK + 4 * ks + lr];  // B[k][n] = K_da(n, k)
          }
#pragma unroll
          for (int tm = 0; tm < 3; ++tm)
#pragma unroll
            for (int tn = tm; tn < 3; ++tn) dmma(acc[tm][tn], fa[tm], fb[tn]);
        }
        __syncwarp();  // mK consumed: K_e is staged over it
        // ---- stage K_e (row stride 26), mirror the off-diagonal tiles, add the geometric part ----------------------------
        double* Ke = sm + L::KE;
#pragma unroll
        for (int tm = 0; tm < 3; ++tm)
#pragma unroll
          for (int tn = tm; tn < 3; ++tn)
#pragma unroll
            for (int i = 0; i < 2; ++i)
            {
              const int gr = 8 * tm + lq, gc = 8 * tn + 2 * lr + i;
              Ke[gr * L::KE_RS + gc] = acc[tm][tn][i];
              if (tn > tm) Ke[gc * L::KE_RS + gr] = acc[tm][tn][i];
            }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int b = 0; b < 3; ++b) Ke[(3 * lq + b) * L::KE_RS + 3 * (2 * lr + i) + b] += kg[i];
      }
      __syncwarp();

      // ---- scatter (as hex8::scatter, one element) -------------------------------------------------------------------
      if (lane < ND && A.scatter_mode == 0)
      {
        const double* Ke = sm + L::KE;
        const int jl = lane / 3, cl = lane - 3 * jl;
        const int32_t fr_ = nrow[jl].frow;
        if (fr_ >= 0) atomicAdd(A.f + fr_ + cl, sm[L::FE + lane]);
        if (STIFF)
#pragma unroll
          for (int h = 0; h < NN; h += 4)
          {
            // the shared-memory reads of four row nodes first, then their REDs (no branch between the loads)
            int64_t off[4];
            int32_t len[4];
            double v[4][3];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
              const NodeRow nr = nrow[h + i];
              off[i] = nr.rowstart >= 0 ? nr.rowstart + pos[(h + i) * NN + jl] + cl : -1;  // -1: row node not owned
              len[i] = nr.rowlen;
#pragma unroll
              for (int b = 0; b < 3; ++b) v[i][b] = Ke[(3 * (h + i) + b) * L::KE_RS + lane];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (off[i] >= 0)
              {
                double* dst = A.vals + off[i];
                atomicAdd(dst, v[i][0]);
                atomicAdd(dst + len[i], v[i][1]);
                atomicAdd(dst + 2 * (int64_t)len[i], v[i][2]);
              }
          }
      }
      __syncwarp();
    }
#endif
  }  // namespace hex8eas
}  // namespace b200s

// hex8_eas_kernel.cuh -- evaluate + assemble for hex8 solid elements with enhanced assumed strains
// (TECH eas_mild: 9 parameters, TECH eas_full: 21): SolidEleCalc<hex8, EASFormulation<hex8, eastype>>
//   src/solid_ele/4C_solid_ele_calc_eas.hpp:44-290            (formulation: prepare / evaluate / condensation)
//   src/solid_ele/4C_solid_ele_calc_eas_helpers.hpp           (M, Mtilde, integrate_eas, update_alpha)
//   src/solid_ele/4C_solid_ele_calc.cpp:137-281,284-314       (driver, recover)
//   src/solid_ele/4C_solid_ele_calc_lib.hpp:1451-1518         (T0^-T at the centroid)
//
// Mapping: one warp = one element.  Lane r < 24 owns dof row r of the element: column r of the
// B-operator, row r of K_dd (24 doubles) and of K_da (NEAS doubles) live in its registers; lane
// a < NEAS owns EAS parameter a: column a of Mtilde, row a of K_aa.  Per Gauss point every lane
// evaluates the (cheap) kinematics and the material redundantly, writes its column of  w C B  and
// w C Mtilde  to shared memory, and after a __syncwarp accumulates its rows against those columns
// (broadcast loads).  After the 8 Gauss points: K_aa is inverted in shared memory by the warp
// (Gauss-Jordan with partial pivoting, the operator LAPACK's getrf/getri gives the reference), the
// element history {s, K_aa^-1, K_da} goes to HBM for struct_calc_recover, the condensed
// contributions  f -= K_da K_aa^-1 s,  K_dd -= K_da K_aa^-1 K_da^T  are applied row-wise, and K_e / f_e
// are scattered exactly like in hex8_kernel.cuh (24 lanes = 24 consecutive columns of one CSR row).
//
// The deformation gradient handed to the material is the displacement-based one: the materials on the
// device path only use the (enhanced) Green-Lagrange strain, so the polar-decomposition based
// "enhanced deformation gradient" of the reference (calc_lib.hpp:1566-1595) is not needed.
#pragma once

#include "hex8_fbar_kernel.cuh"
#include "hex8_kernel.cuh"

namespace b200s
{
  namespace hex8eas
  {
    constexpr int NN = 8, NGP = 8, ND = 24;
    // element history in HBM, per element (device element order), in the order the reference packs it
    // (eas.hpp:266-275): alpha_inc[NEAS], alpha[NEAS], s[NEAS], invKaa[NEAS*NEAS], Kda[24*NEAS],
    // matrices column-major
    B2_HD constexpr int hist_size(int neas) { return 3 * neas + neas * neas + ND * neas; }

#if defined(__CUDACC__)
    // one non-zero per column of the parametric EAS shape function matrix M(xi)
    // (evaluate_eas_shape_functions_parameter_space, eas_helpers.hpp:205-275): Voigt row and monomial
    __device__ __forceinline__ int m_row(int c)
    {
      constexpr int t[21] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 3, 3, 4, 4, 5, 5, 0, 0, 1, 1, 2, 2};
      return t[c];
    }
    __device__ __forceinline__ double m_val(int c, double r, double s, double t)
    {
      constexpr int mono[21] = {0, 1, 2, 0, 1, 1, 2, 0, 2, 4, 5, 3, 4, 3, 5, 3, 4, 3, 5, 4, 5};
      const double v[6] = {r, s, t, r * s, r * t, s * t};
      return v[mono[c]];
    }

    // shared memory of one warp, in doubles
    template <int NEAS>
    struct Layout
    {
      static constexpr int NA = (NEAS + 3) & ~3;          // padded
      static constexpr int XU = 0;                        // X[24], u[24]
      static constexpr int T0 = XU + 48;                  // T0invT 6x6 column-major
      static constexpr int ALPHA = T0 + 36;               // alpha[NA]
      static constexpr int SG = ALPHA + NA;               // s vector after the loop [NA]
      static constexpr int GALL = SG + NA;                // G_k(a) of the current Gauss point [24]
      static constexpr int CB = GALL + 24;                // w C B      6 x 24, [k + 6 r]
      static constexpr int CM = CB + 144;                 // w C Mtilde 6 x NEAS, [k + 6 a]
      static constexpr int WORK = CM + 6 * NA;            // post-loop work area
      //   augmented [Kaa | I] NEAS x 2 NEAS row-major, later invKaa NEAS x NEAS [k*NEAS + a]
      static constexpr int AUG = WORK;
      static constexpr int KDA = AUG + 2 * NEAS * NEAS + (2 * NEAS * NEAS & 1);  // Kda 24 x NEAS [r*NEAS + a]
      static constexpr int KE = KDA + ND * NEAS + (ND * NEAS & 1);               // K_e 24 x 24 row-major, f_e behind
      static constexpr int DOUBLES = KE + 576 + 24;
      static constexpr int BYTES = DOUBLES * 8 + NN * 16 + NN * NN * 4;          // + NodeRow[8] + pos[64]
    };

    // In-place inverse of the n x n matrix in the left half of the row-major augmented [A | I]
    // (n x 2n) in shared memory, by the whole warp: Gauss-Jordan with partial pivoting.  On return the
    // right half holds A^-1.  A singular pivot is used as it is (the reference ignores the LAPACK
    // error, solve_for_inverse_ignoring_errors).
    __device__ __forceinline__ void warp_gauss_jordan(int n, double* aug, int lane)
    {
      const int w = 2 * n;
      for (int c = 0; c < n; ++c)
      {
        // pivot search (every lane reads the same column: uniform result)
        int piv = c;
        double best = fabs(aug[c * w + c]);
        for (int r = c + 1; r < n; ++r)
        {
          const double v = fabs(aug[r * w + c]);
          if (v > best) best = v, piv = r;
        }
        if (piv != c)
          for (int k = lane; k < w; k += 32)
          {
            const double t = aug[c * w + k];
            aug[c * w + k] = aug[piv * w + k];
            aug[piv * w + k] = t;
          }
        __syncwarp();
        const double d = 1.0 / aug[c * w + c];
        __syncwarp();
        for (int k = lane; k < w; k += 32) aug[c * w + k] *= d;
        __syncwarp();
        // eliminate column c from every other row; lane = row, so each lane reads its factor first
        for (int r = lane; r < n; r += 32)
        {
          if (r == c) continue;
          const double f = aug[r * w + c];
          if (f != 0.0)
            for (int k = 0; k < w; ++k) aug[r * w + k] = fma(-f, aug[c * w + k], aug[r * w + k]);
        }
        __syncwarp();
      }
    }

    // The warp's work on one element.  sm: the warp's shared memory (Layout<NEAS>), hist: the element's
    // history block in HBM.
    template <int MAT, bool STIFF, int NEAS>
    __device__ __forceinline__ void element(const EvalArgs& A, double* sm, NodeRow* nrow, int32_t* pos,
        const double* tab_gp /*[g][k*3+a], stride TAB_STRIDE*/, const double* wgt, const double* gxi /*[g][3]*/,
        int lane, int64_t e, double* hist, double& detJ_min, double& detF_min, bool& neg_jac)
    {
      using L = Layout<NEAS>;
      double* xu = sm + L::XU;
      double* T0 = sm + L::T0;
      double* alpha = sm + L::ALPHA;

      // ---- gather ----------------------------------------------------------------------------------
      if (lane < NN)
      {
        const int64_t node = A.conn[e * NN + lane];
#pragma unroll
        for (int d = 0; d < 3; ++d)
        {
          xu[3 * lane + d] = A.X[3 * node + d];
          xu[24 + 3 * lane + d] = A.u[3 * node + d];
        }
        nrow[lane] = A.node_row[node];
      }
      for (int t = lane; t < NN * NN; t += 32) pos[t] = A.pos[e * (NN * NN) + t];
      if (lane < NEAS) alpha[lane] = hist[NEAS + lane];
      __syncwarp();

      // ---- prepare: centroid transformation (eas.hpp:66-70) -------------------------------------------
      double detJ0;
      {
        double J0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) J0[i] = 0.0;
#pragma unroll
        for (int k = 0; k < NN; ++k)
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            const double dn = 0.125 * hex8fbar::sgn(k, a);
#pragma unroll
            for (int b = 0; b < 3; ++b) J0[a + 3 * b] = fma(dn, xu[3 * k + b], J0[a + 3 * b]);
          }
        detJ0 = det3(J0);
        // T^T in strain-like Voigt notation (calc_lib.hpp:1456-1508) into the augmented work area
        double* aug = sm + L::AUG;  // 6 x 12
        if (lane < 6)
        {
          const int c = lane;
          double col[6];
          if (c < 3)
          {
            col[0] = J0[0 + 3 * c] * J0[0 + 3 * c], col[1] = J0[1 + 3 * c] * J0[1 + 3 * c];
            col[2] = J0[2 + 3 * c] * J0[2 + 3 * c], col[3] = 2 * J0[0 + 3 * c] * J0[1 + 3 * c];
            col[4] = 2 * J0[1 + 3 * c] * J0[2 + 3 * c], col[5] = 2 * J0[0 + 3 * c] * J0[2 + 3 * c];
          }
          else
          {
            const int a = (c == 4) ? 1 : 0, b = (c == 3) ? 1 : 2;
            col[0] = J0[0 + 3 * a] * J0[0 + 3 * b], col[1] = J0[1 + 3 * a] * J0[1 + 3 * b];
            col[2] = J0[2 + 3 * a] * J0[2 + 3 * b];
            col[3] = J0[0 + 3 * a] * J0[1 + 3 * b] + J0[1 + 3 * a] * J0[0 + 3 * b];
            col[4] = J0[1 + 3 * a] * J0[2 + 3 * b] + J0[2 + 3 * a] * J0[1 + 3 * b];
            col[5] = J0[0 + 3 * a] * J0[2 + 3 * b] + J0[2 + 3 * a] * J0[0 + 3 * b];
          }
#pragma unroll
          for (int r = 0; r < 6; ++r)
          {
            aug[r * 12 + c] = col[r];
            aug[r * 12 + 6 + c] = (r == c) ? 1.0 : 0.0;
          }
        }
        __syncwarp();
        warp_gauss_jordan(6, aug, lane);
        for (int t = lane; t < 36; t += 32) T0[(t % 6) + 6 * (t / 6)] = aug[(t % 6) * 12 + 6 + (t / 6)];
        __syncwarp();
      }

      // nodal Jacobian check (calc_lib.hpp:543-562): lane < 8 = node
      {
        double J[9];
        const int g = lane & 7;
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

      // ---- Gauss loop ----------------------------------------------------------------------------------
      const int r_ = lane < ND ? lane : 0;      // dof row of this lane (lanes >= 24 shadow row 0, results unused)
      const int ni = r_ / 3, ei = r_ - 3 * ni;  // node and direction of the row
      const int a_ = lane < NEAS ? lane : 0;    // EAS parameter of this lane
      const int arow = m_row(a_);
      double Kdd[ND], Kda[NEAS], Kaa[NEAS], facc = 0.0, sacc = 0.0;
#pragma unroll
      for (int c = 0; c < ND; ++c) Kdd[c] = 0.0;
#pragma unroll
      for (int c = 0; c < NEAS; ++c) Kda[c] = 0.0, Kaa[c] = 0.0;

#pragma unroll 1
      for (int g = 0; g < NGP; ++g)
      {
        const double* dN = tab_gp + g * hex8::TAB_STRIDE;
        double J[9], iJ[9];
        hex8::jacobian(dN, xu, J);
        const double detJ = inv3(iJ, J);
        detJ_min = fmin(detJ_min, detJ);
        // F = I + u G^T (redundant in every lane), G_i of the lane's own node
        double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        double Gi[3] = {0, 0, 0}, Gl[3] = {0, 0, 0};  // of the row's node / of node (lane & 7)
#pragma unroll
        for (int k = 0; k < NN; ++k)
        {
          double Gk[3];
#pragma unroll
          for (int a = 0; a < 3; ++a)
            Gk[a] = iJ[a] * dN[3 * k] + iJ[a + 3] * dN[3 * k + 1] + iJ[a + 6] * dN[3 * k + 2];
          if (k == ni) Gi[0] = Gk[0], Gi[1] = Gk[1], Gi[2] = Gk[2];
          if ((lane & 7) == k) Gl[0] = Gk[0], Gl[1] = Gk[1], Gl[2] = Gk[2];
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            const double ua = xu[24 + 3 * k + a];
#pragma unroll
            for (int b = 0; b < 3; ++b) F[a + 3 * b] += ua * Gk[b];
          }
        }
        detF_min = fmin(detF_min, det3(F));

        // column r of the B-operator (evaluate_strain_gradient, calc_lib.hpp:777-826), strain-like rows
        double Bc[6];
        Bc[0] = F[ei] * Gi[0], Bc[1] = F[ei + 3] * Gi[1], Bc[2] = F[ei + 6] * Gi[2];
        Bc[3] = F[ei] * Gi[1] + F[ei + 3] * Gi[0];
        Bc[4] = F[ei + 3] * Gi[2] + F[ei + 6] * Gi[1];
        Bc[5] = F[ei + 6] * Gi[0] + F[ei] * Gi[2];

        // Mtilde = detJ0/detJ T0invT M(xi): column a of this lane, and Mtilde alpha for the strain
        const double xr = gxi[3 * g], xs = gxi[3 * g + 1], xt = gxi[3 * g + 2];
        const double scale = detJ0 / detJ;
        double Mc[6];
        {
          const double mv = scale * m_val(a_, xr, xs, xt);
#pragma unroll
          for (int k = 0; k < 6; ++k) Mc[k] = mv * T0[k + 6 * arow];
        }
        double C[6];
        right_cauchy_green(F, C);
        {
          double v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int c = 0; c < NEAS; ++c)
          {
            constexpr int rows[21] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 3, 3, 4, 4, 5, 5, 0, 0, 1, 1, 2, 2};
            constexpr int mono[21] = {0, 1, 2, 0, 1, 1, 2, 0, 2, 4, 5, 3, 4, 3, 5, 3, 4, 3, 5, 4, 5};
            const double mon = mono[c] == 0 ? xr : mono[c] == 1 ? xs : mono[c] == 2 ? xt
                             : mono[c] == 3 ? xr * xs : mono[c] == 4 ? xr * xt : xs * xt;
            v[rows[c]] = fma(mon, alpha[c], v[rows[c]]);
          }
          // E_enh = E + stress-like(Mtilde alpha); C_enh = 2 E_enh + I = C + 2 (...)  (shear rows halved)
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
        material_evaluate<MAT>(mp, C, Sv, Cm);
        const double w = detJ * wgt[g];

        // columns of  w C B  and  w C Mtilde  (integrate_eas: cmatM, eas_helpers.hpp:352-371)
        double CBc[6], CMc[6];
#pragma unroll
        for (int I = 0; I < 6; ++I)
        {
          double tb = 0.0, tm = 0.0;
#pragma unroll
          for (int Jv = 0; Jv < 6; ++Jv)
          {
            tb = fma(Cm[c21(I, Jv)], Bc[Jv], tb);
            tm = fma(Cm[c21(I, Jv)], Mc[Jv], tm);
          }
          CBc[I] = w * tb, CMc[I] = w * tm;
        }
        __syncwarp();  // previous Gauss point's CB / CM / G fully consumed
        if (lane < NN) sm[L::GALL + 3 * lane] = Gl[0], sm[L::GALL + 3 * lane + 1] = Gl[1], sm[L::GALL + 3 * lane + 2] = Gl[2];
        if (lane < ND)
#pragma unroll
          for (int I = 0; I < 6; ++I) sm[L::CB + I + 6 * lane] = CBc[I];
        if (lane < NEAS)
#pragma unroll
          for (int I = 0; I < 6; ++I) sm[L::CM + I + 6 * lane] = CMc[I];
        __syncwarp();

        // force and EAS residual: f_r += w B_r . S, s_a += w Mtilde_a . S
        double bs = 0.0, ms = 0.0;
#pragma unroll
        for (int I = 0; I < 6; ++I) bs = fma(Bc[I], Sv[I], bs), ms = fma(Mc[I], Sv[I], ms);
        facc = fma(w, bs, facc);
        sacc = fma(w, ms, sacc);
        // K_da(r, a) += B_r . (w C Mtilde_a);  K_aa(a, a') += Mtilde_a . (w C Mtilde_a')
#pragma unroll
        for (int c = 0; c < NEAS; ++c)
        {
          const double* cm = sm + L::CM + 6 * c;
          double t1 = 0.0, t2 = 0.0;
#pragma unroll
          for (int I = 0; I < 6; ++I) t1 = fma(Bc[I], cm[I], t1), t2 = fma(Mc[I], cm[I], t2);
          Kda[c] += t1, Kaa[c] += t2;
        }
        if (STIFF)
        {
          // elastic: K_dd(r, c) += B_r . (w C B_c); geometric: K_dd(3i+b, 3j+b) += w G_i.S.G_j
          double gi[3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
            gi[d] = w * (Gi[0] * Sv[vg(0, d)] + Gi[1] * Sv[vg(1, d)] + Gi[2] * Sv[vg(2, d)]);
#pragma unroll
          for (int c = 0; c < ND; ++c)
          {
            const double* cb = sm + L::CB + 6 * c;
            double t = 0.0;
#pragma unroll
            for (int I = 0; I < 6; ++I) t = fma(Bc[I], cb[I], t);
            Kdd[c] += t;
          }
#pragma unroll
          for (int j = 0; j < NN; ++j)
          {
            const double* gj = sm + L::GALL + 3 * j;
            const double sg = gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2];
#pragma unroll
            for (int b = 0; b < 3; ++b)
              if (b == ei) Kdd[3 * j + b] += sg;
          }
        }
      }
      __syncwarp();

      // ---- condensation ----------------------------------------------------------------------------------
      double* aug = sm + L::AUG;
      double* KdaS = sm + L::KDA;
      double* sS = sm + L::SG;
      if (lane < NEAS)
      {
#pragma unroll
        for (int c = 0; c < NEAS; ++c)
        {
          aug[lane * 2 * NEAS + c] = Kaa[c];
          aug[lane * 2 * NEAS + NEAS + c] = (c == lane) ? 1.0 : 0.0;
        }
        sS[lane] = sacc;
      }
      if (lane < ND)
#pragma unroll
        for (int c = 0; c < NEAS; ++c) KdaS[lane * NEAS + c] = Kda[c];
      __syncwarp();
      warp_gauss_jordan(NEAS, aug, lane);
      // history for struct_calc_recover: s, invKaa (column-major), Kda (column-major)
      {
        double* h_s = hist + 2 * NEAS;
        double* h_inv = hist + 3 * NEAS;
        double* h_kda = h_inv + NEAS * NEAS;
        if (lane < NEAS) h_s[lane] = sS[lane];
        for (int t = lane; t < NEAS * NEAS; t += 32)  // t = r + NEAS c
          h_inv[t] = aug[(t % NEAS) * 2 * NEAS + NEAS + t / NEAS];
        for (int t = lane; t < ND * NEAS; t += 32)  // t = r + 24 c
          h_kda[t] = KdaS[(t % ND) * NEAS + t / ND];
      }
      // mK(r, a) = -sum_k Kda(r, k) invKaa(k, a)   (prepare_condensed_contribution, eas.hpp:188-203)
      double mK[NEAS];
#pragma unroll
      for (int a = 0; a < NEAS; ++a) mK[a] = 0.0;
#pragma unroll
      for (int k = 0; k < NEAS; ++k)
      {
        const double* row = aug + k * 2 * NEAS + NEAS;
#pragma unroll
        for (int a = 0; a < NEAS; ++a) mK[a] = fma(-Kda[k], row[a], mK[a]);
      }
      // f_r += mK(r,:) s ; K_dd(r, c) += mK(r,:) Kda(c,:)   (eas.hpp:245-263)
#pragma unroll
      for (int a = 0; a < NEAS; ++a) facc = fma(mK[a], sS[a], facc);
      if (STIFF)
#pragma unroll
        for (int c = 0; c < ND; ++c)
        {
          const double* kc = KdaS + c * NEAS;
          double t = 0.0;
#pragma unroll
          for (int a = 0; a < NEAS; ++a) t = fma(mK[a], kc[a], t);
          Kdd[c] += t;
        }
      // ---- stage + scatter (as hex8::scatter, one element) -----------------------------------------------
      double* Ke = sm + L::KE;
      double* fe = Ke + 576;
      if (lane < ND)
      {
        fe[lane] = facc;
        if (STIFF)
#pragma unroll
          for (int c = 0; c < ND; ++c) Ke[lane * ND + c] = Kdd[c];
      }
      __syncwarp();
      if (lane < ND && A.scatter_mode == 0)
      {
        const int jl = lane / 3, cl = lane - 3 * jl;
        const int32_t fr = nrow[jl].frow;
        if (fr >= 0) atomicAdd(A.f + fr + cl, fe[lane]);
        if (STIFF)
#pragma unroll
          for (int i = 0; i < NN; ++i)
          {
            const NodeRow nr = nrow[i];
            if (nr.rowstart < 0) continue;
            double* dst = A.vals + nr.rowstart + pos[i * NN + jl] + cl;
            atomicAdd(dst, Ke[(3 * i) * ND + lane]);
            atomicAdd(dst + nr.rowlen, Ke[(3 * i + 1) * ND + lane]);
            atomicAdd(dst + 2 * (int64_t)nr.rowlen, Ke[(3 * i + 2) * ND + lane]);
          }
      }
      __syncwarp();
    }
#endif
  }  // namespace hex8eas
}  // namespace b200s

// hex8_eas_kernel.cuh -- evaluate + assemble for hex8 solid elements with enhanced assumed strains
// (TECH eas_mild: 9 parameters, TECH eas_full: 21): SolidEleCalc<hex8, EASFormulation<hex8, eastype>>
//   src/solid_ele/4C_solid_ele_calc_eas.hpp:44-290            (formulation: prepare / evaluate / condensation)
//   src/solid_ele/4C_solid_ele_calc_eas_helpers.hpp           (M, Mtilde, integrate_eas, update_alpha)
//   src/solid_ele/4C_solid_ele_calc.cpp:137-281,284-314       (driver, recover)
//   src/solid_ele/4C_solid_ele_calc_lib.hpp:1451-1518         (T0^-T at the centroid)
//
// Mapping: one warp = one element.  Lane r < 24 owns dof row r of the element: column r of the
// B-operator, row r of K_dd (24 doubles) and of K_da (NEAS doubles) live in its registers; lane
// a < NEAS owns EAS parameter a: column a of Mtilde, row a of K_aa.  Phase 1: lane g < 8 evaluates
// Gauss point g once (Jacobian, dN/dX, F, enhanced strain, material) and leaves F, w S, w C, dN/dX in
// shared memory.  Phase 2, per Gauss point: every lane forms its column of B and Mtilde, writes its
// column of  w C B  and  w C Mtilde  to shared memory, and after a __syncwarp accumulates its rows
// against those columns (broadcast loads).  After the 8 Gauss points: K_aa is inverted in shared memory by the warp
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
    // (0 r, 1 s, 2 t, 3 rs, 4 rt, 5 st) of column c
    __constant__ int c_eas_row[21] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 3, 3, 4, 4, 5, 5, 0, 0, 1, 1, 2, 2};
    __constant__ int c_eas_mono[21] = {0, 1, 2, 0, 1, 1, 2, 0, 2, 4, 5, 3, 4, 3, 5, 3, 4, 3, 5, 4, 5};
    // eastype_sh8_7 (TECH shell_eas, the solid-shell interpolation with t = thickness direction, eas_helpers.hpp:330-357):
    // r | s | t, rt, st | r, s in the rows rr | ss | tt | rs
    __constant__ int c_eas_row7[7] = {0, 1, 2, 3, 3, 2, 2};
    __constant__ int c_eas_mono7[7] = {0, 1, 2, 0, 1, 4, 5};
    __device__ __forceinline__ double monomial(int m, double r, double s, double t)
    {
      return m == 0 ? r : m == 1 ? s : m == 2 ? t : m == 3 ? r * s : m == 4 ? r * t : s * t;
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
      static constexpr int CB = SG + NA;                  // w C B      6 x 24, [k + 6 r]
      static constexpr int CM = CB + 144;                 // w C Mtilde 6 x NEAS, [k + 6 a]
      // per Gauss point, written once by lane g before the loop: F[9], w S[6], w Cm[21], G[24], scale
      static constexpr int GP = CM + 6 * NA;
      static constexpr int GP_STRIDE = 62;
      static constexpr int GP_F = 0, GP_S = 9, GP_C = 15, GP_G = 36, GP_SCALE = 60;
      static constexpr int LOOP_END = GP + NGP * GP_STRIDE;
      // K_e 24 x 24 row-major with f_e behind: staged after the Gauss loop, aliases CB / CM / GP (dead then)
      static constexpr int KE = CB;
      static_assert(KE + 576 + 24 <= LOOP_END, "K_e staging fits into the dead Gauss-loop buffers");
      // post-loop work area: Kaa NEAS x NEAS row-major, inverted in place, and Kda 24 x NEAS [r*NEAS + a]
      static constexpr int AUG = LOOP_END;
      static constexpr int KDA = AUG + NEAS * NEAS + (NEAS * NEAS & 1);
      static constexpr int DOUBLES = KDA + ND * NEAS + (ND * NEAS & 1);
      static constexpr int BYTES = DOUBLES * 8 + NN * 16 + NN * NN * 4;          // + NodeRow[8] + pos[64]
    };

    // In-place inverse of the n x n matrix in the left half of the row-major augmented [A | I]
    // (n x 2n, n <= 32) in shared memory, by the whole warp: Gauss-Jordan with partial pivoting (pivot
    // search as a warp arg-max).  On return the right half holds A^-1.  A singular pivot is used as it is
    // (the reference ignores the LAPACK error, solve_for_inverse_ignoring_errors).
    __device__ __forceinline__ void warp_gauss_jordan(int n, double* aug, int lane)
    {
      const int w = 2 * n;
      for (int c = 0; c < n; ++c)
      {
        // pivot: largest |a(r, c)|, r >= c; ties to the lowest row like a sequential search
        double best = (lane >= c && lane < n) ? fabs(aug[lane * w + c]) : -1.0;
        int piv = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int op = __shfl_xor_sync(0xffffffffu, piv, o);
          if (ob > best || (ob == best && op < piv)) best = ob, piv = op;
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
        // eliminate column c from every other row; lane = row
        if (lane < n && lane != c)
        {
          const double f = aug[lane * w + c];
          if (f != 0.0)
            for (int k = 0; k < w; ++k) aug[lane * w + k] = fma(-f, aug[c * w + k], aug[lane * w + k]);
        }
        __syncwarp();
      }
    }

    // In-place inverse of a symmetric positive definite n x n matrix (row-major, n <= 32) in shared memory
    // by the warp, lane = row: Gauss-Jordan without pivoting (two __syncwarp per step).  K_aa = sum w Mt^T C Mt
    // is positive definite whenever the material tangent is; the reference's LU with partial pivoting
    // (solve_for_inverse_ignoring_errors) gives the same operator up to round-off then.
    __device__ __forceinline__ void warp_spd_inverse(int n, double* a, int lane)
    {
      for (int c = 0; c < n; ++c)
      {
        const double d = 1.0 / a[c * n + c];
        if (lane < n && lane != c)
        {
          const double f = a[lane * n + c] * d;
          for (int k = 0; k < n; ++k)
            if (k != c) a[lane * n + k] = fma(-f, a[c * n + k], a[lane * n + k]);
          a[lane * n + c] = -f;
        }
        __syncwarp();
        if (lane < n) a[c * n + lane] = (lane == c) ? d : a[c * n + lane] * d;
        __syncwarp();
      }
    }

    // Shared-memory flavour with a compile-time size: the k loop is unrolled and branch free (the pivot
    // column is overwritten afterwards), the step loop stays rolled to keep the code small.
    template <int N>
    __device__ __forceinline__ void warp_spd_inverse_t(double* a, int lane)
    {
#pragma unroll 1
      for (int c = 0; c < N; ++c)
      {
        const double d = 1.0 / a[c * N + c];
        if (lane < N && lane != c)
        {
          const double f = a[lane * N + c] * d;
          const double* pr = a + c * N;
          double* row = a + lane * N;
#pragma unroll
          for (int k = 0; k < N; ++k) row[k] = fma(-f, pr[k], row[k]);
          row[c] = -f;
        }
        __syncwarp();
        if (lane < N) a[c * N + lane] = (lane == c) ? d : a[c * N + lane] * d;
        __syncwarp();
      }
    }

    // The same inverse with the rows in registers: lane r holds row r of the matrix in a[] (lanes >= N hold
    // anything and are ignored).  Per elimination step the pivot lane publishes its scaled row through a
    // double-buffered 2 N shared-memory scratch (one __syncwarp per step); every index is a compile-time
    // constant after unrolling, so a[] never leaves the register file.  On return a[] is row `lane` of A^-1.
    template <int N>
    __device__ __forceinline__ void warp_spd_inverse_reg(double (&a)[N], double* scratch, int lane)
    {
#pragma unroll
      for (int c = 0; c < N; ++c)
      {
        double* buf = scratch + (c & 1) * N;
        if (lane == c)
        {
          const double d = 1.0 / a[c];
#pragma unroll
          for (int k = 0; k < N; ++k) a[k] = (k == c) ? d : a[k] * d;
#pragma unroll
          for (int k = 0; k < N; ++k) buf[k] = a[k];
        }
        __syncwarp();
        if (lane != c)
        {
          const double f = a[c];
#pragma unroll
          for (int k = 0; k < N; ++k) a[k] = (k == c) ? -f * buf[c] : fma(-f, buf[k], a[k]);
        }
      }
      __syncwarp();
    }

    // Set-up (once per mesh): T0invT (6x6 column-major) and det J0 of one element into t0[0..35], t0[36]
    // (evaluate_centroid_transformation, eas_helpers.hpp:118-133); aug = 72 doubles of shared memory
    __device__ __forceinline__ void centroid_transformation(const double* X, const int32_t* ce, double* aug,
        int lane, double* t0)
    {
      double J0[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) J0[i] = 0.0;
#pragma unroll
      for (int k = 0; k < NN; ++k)
      {
        const double* x = X + 3 * (int64_t)ce[k];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          const double dn = 0.125 * hex8fbar::sgn(k, a);
#pragma unroll
          for (int b = 0; b < 3; ++b) J0[a + 3 * b] = fma(dn, x[b], J0[a + 3 * b]);
        }
      }
      // T^T in strain-like Voigt notation (calc_lib.hpp:1456-1508) into the augmented work area (6 x 12);
      // lane c writes column c (static register indices: the column loop is unrolled)
#pragma unroll
      for (int c = 0; c < 6; ++c)
      {
        if (lane != c) continue;
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
      for (int t = lane; t < 36; t += 32) t0[(t % 6) + 6 * (t / 6)] = aug[(t % 6) * 12 + 6 + (t / 6)];
      if (lane == 0) t0[36] = det3(J0);
      __syncwarp();
    }
    constexpr int T0_STRIDE = 40;  // doubles per element in the precomputed table

    // The warp's work on one element.  sm: the warp's shared memory (Layout<NEAS>), hist: the element's
    // history block in HBM.
    template <int MAT, bool STIFF, int NEAS>
    __device__ __forceinline__ void element(const EvalArgs& A, double* sm, NodeRow* nrow, int32_t* pos,
        const double* tab_gp /*[g][k*3+a], stride TAB_STRIDE*/, const double* wgt, const double* gxi /*[g][3]*/,
        int lane, int64_t e, double* hist, const double* t0tab, double& detJ_min, double& detF_min, bool& neg_jac)
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

      // ---- prepare: centroid transformation (eas.hpp:66-70), precomputed per element at set-up ---------
      for (int t = lane; t < 36; t += 32) T0[t] = t0tab[t];
      const double detJ0 = t0tab[36];
      __syncwarp();

      // ---- phase 1: lane g < 8 = Gauss point g (and element node g for the Jacobian check) -----------------
      if (lane < NGP)
      {
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
        double* gp = sm + L::GP + g * L::GP_STRIDE;
        double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
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
            gp[L::GP_G + 3 * k + a] = Gk[a];
            const double ua = xu[24 + 3 * k + a];
#pragma unroll
            for (int b = 0; b < 3; ++b) F[a + 3 * b] += ua * Gk[b];
          }
        }
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
            constexpr int rows7[7] = {0, 1, 2, 3, 3, 2, 2}, mono7[7] = {0, 1, 2, 0, 1, 4, 5};  // eastype_sh8_7
            const int rw = NEAS == 7 ? rows7[c < 7 ? c : 0] : rows[c], mn = NEAS == 7 ? mono7[c < 7 ? c : 0] : mono[c];
            v[rw] = fma(monomial(mn, xr, xs, xt), alpha[c], v[rw]);
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
        for (int i = 0; i < 9; ++i) gp[L::GP_F + i] = F[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) gp[L::GP_S + i] = w * Sv[i];
#pragma unroll
        for (int i = 0; i < 21; ++i) gp[L::GP_C + i] = w * Cm[i];
        gp[L::GP_SCALE] = scale;
      }
      __syncwarp();

      // ---- phase 2: lane = dof row r (< 24) and EAS parameter a (< NEAS) -------------------------------------
      // Two passes over the Gauss points so that the register-resident rows never coexist: first K_da, K_aa,
      // f and s (42 accumulators), condensation operator mK = -K_da K_aa^-1 (21), then K_dd (24).
      const int r_ = lane < ND ? lane : 0;      // lanes >= 24 shadow row 0, results unused
      const int ni = r_ / 3, ei = r_ - 3 * ni;  // node and direction of the row
      const int a_ = lane < NEAS ? lane : 0;
      const int arow = NEAS == 7 ? c_eas_row7[a_] : c_eas_row[a_], amono = NEAS == 7 ? c_eas_mono7[a_] : c_eas_mono[a_];
      double facc = 0.0;
      double mK[NEAS];
      double* aug = sm + L::AUG;
      double* KdaS = sm + L::KDA;
      double* sS = sm + L::SG;
      {
        double Kda[NEAS], Kaa[NEAS], sacc = 0.0;
#pragma unroll
        for (int c = 0; c < NEAS; ++c) Kda[c] = 0.0, Kaa[c] = 0.0;
#pragma unroll 1
        for (int g = 0; g < NGP; ++g)
        {
          const double* gp = sm + L::GP + g * L::GP_STRIDE;
          const double* Gi = gp + L::GP_G + 3 * ni;
          const double Fe0 = gp[L::GP_F + ei], Fe1 = gp[L::GP_F + ei + 3], Fe2 = gp[L::GP_F + ei + 6];
          const double G0 = Gi[0], G1 = Gi[1], G2 = Gi[2];
          // column r of the B-operator (evaluate_strain_gradient, calc_lib.hpp:777-826), strain-like rows
          double Bc[6];
          Bc[0] = Fe0 * G0, Bc[1] = Fe1 * G1, Bc[2] = Fe2 * G2;
          Bc[3] = Fe0 * G1 + Fe1 * G0;
          Bc[4] = Fe1 * G2 + Fe2 * G1;
          Bc[5] = Fe2 * G0 + Fe0 * G2;
          // column a of Mtilde and of  w C Mtilde  (integrate_eas: cmatM, eas_helpers.hpp:352-371)
          double Mc[6], CMc[6];
          {
            const double mv = gp[L::GP_SCALE] * monomial(amono, gxi[3 * g], gxi[3 * g + 1], gxi[3 * g + 2]);
#pragma unroll
            for (int k = 0; k < 6; ++k) Mc[k] = mv * T0[k + 6 * arow];
#pragma unroll
            for (int I = 0; I < 6; ++I)
            {
              double tm = 0.0;
#pragma unroll
              for (int Jv = 0; Jv < 6; ++Jv) tm = fma(gp[L::GP_C + c21(I, Jv)], Mc[Jv], tm);
              CMc[I] = tm;
            }
          }
          __syncwarp();  // previous Gauss point's CM fully consumed
          if (lane < NEAS)
#pragma unroll
            for (int I = 0; I < 6; ++I) sm[L::CM + I + 6 * lane] = CMc[I];
          __syncwarp();
          // force and EAS residual: f_r += B_r . (w S), s_a += Mtilde_a . (w S)
#pragma unroll
          for (int I = 0; I < 6; ++I)
          {
            const double ws = gp[L::GP_S + I];
            facc = fma(Bc[I], ws, facc), sacc = fma(Mc[I], ws, sacc);
          }
          // K_da(r, a) += B_r . (w C Mtilde_a);  K_aa(a, a') += Mtilde_a . (w C Mtilde_a')
#pragma unroll
          for (int c = 0; c < NEAS; ++c)
          {
            const d2* cm2 = reinterpret_cast<const d2*>(sm + L::CM + 6 * c);
            const d2 q0 = cm2[0], q1 = cm2[1], q2 = cm2[2];
            const double cm[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
            double t1 = 0.0, t2 = 0.0;
#pragma unroll
            for (int I = 0; I < 6; ++I) t1 = fma(Bc[I], cm[I], t1), t2 = fma(Mc[I], cm[I], t2);
            Kda[c] += t1, Kaa[c] += t2;
          }
        }
        __syncwarp();

        // ---- condensation operator ---------------------------------------------------------------------
        if (lane < NEAS) sS[lane] = sacc;
        if (lane < ND)
#pragma unroll
          for (int c = 0; c < NEAS; ++c) KdaS[lane * NEAS + c] = Kda[c];
        // K_aa^-1 into `aug` (row-major).  9 x 9: rows stay in the registers of their lanes (fully unrolled);
        // 21 x 21: in shared memory with a rolled step loop (measured: the unrolled register form is slower
        // there, 34.9 vs 32.9 ms, the 441-FMA update no longer fits the instruction cache together with the rest)
        if (NEAS <= 9)
        {
          warp_spd_inverse_reg<NEAS>(Kaa, aug, lane);  // scratch = the first 2 NEAS doubles of aug
          if (lane < NEAS)
#pragma unroll
            for (int c = 0; c < NEAS; ++c) aug[lane * NEAS + c] = Kaa[c];
          __syncwarp();
        }
        else
        {
          if (lane < NEAS)
#pragma unroll
            for (int c = 0; c < NEAS; ++c) aug[lane * NEAS + c] = Kaa[c];
          __syncwarp();
          warp_spd_inverse_t<NEAS>(aug, lane);
        }
        // history for struct_calc_recover: s, invKaa (column-major), Kda (column-major)
        {
          double* h_s = hist + 2 * NEAS;
          double* h_inv = hist + 3 * NEAS;
          double* h_kda = h_inv + NEAS * NEAS;
          if (lane < NEAS) h_s[lane] = sS[lane];
          for (int t = lane; t < NEAS * NEAS; t += 32)  // t = r + NEAS c
            h_inv[t] = aug[(t % NEAS) * NEAS + t / NEAS];
          for (int t = lane; t < ND * NEAS; t += 32)  // t = r + 24 c
            h_kda[t] = KdaS[(t % ND) * NEAS + t / ND];
        }
        // mK(r, a) = -sum_k Kda(r, k) invKaa(k, a)   (prepare_condensed_contribution, eas.hpp:188-203)
#pragma unroll
        for (int a = 0; a < NEAS; ++a) mK[a] = 0.0;
#pragma unroll
        for (int k = 0; k < NEAS; ++k)
        {
          const double* row = aug + k * NEAS;
#pragma unroll
          for (int a = 0; a < NEAS; ++a) mK[a] = fma(-Kda[k], row[a], mK[a]);
        }
        // f_r += mK(r,:) s   (eas.hpp:245-253)
#pragma unroll
        for (int a = 0; a < NEAS; ++a) facc = fma(mK[a], sS[a], facc);
      }

      // ---- K_dd: second pass over the Gauss points, then K_dd(r, c) += mK(r,:) Kda(c,:)  (eas.hpp:255-263) ---
      double Kdd[ND];
      if (STIFF)
      {
#pragma unroll
        for (int c = 0; c < ND; ++c) Kdd[c] = 0.0;
#pragma unroll 1
        for (int g = 0; g < NGP; ++g)
        {
          const double* gp = sm + L::GP + g * L::GP_STRIDE;
          const double* Gi = gp + L::GP_G + 3 * ni;
          const double Fe0 = gp[L::GP_F + ei], Fe1 = gp[L::GP_F + ei + 3], Fe2 = gp[L::GP_F + ei + 6];
          const double G0 = Gi[0], G1 = Gi[1], G2 = Gi[2];
          double Bc[6], CBc[6];
          Bc[0] = Fe0 * G0, Bc[1] = Fe1 * G1, Bc[2] = Fe2 * G2;
          Bc[3] = Fe0 * G1 + Fe1 * G0;
          Bc[4] = Fe1 * G2 + Fe2 * G1;
          Bc[5] = Fe2 * G0 + Fe0 * G2;
#pragma unroll
          for (int I = 0; I < 6; ++I)
          {
            double tb = 0.0;
#pragma unroll
            for (int Jv = 0; Jv < 6; ++Jv) tb = fma(gp[L::GP_C + c21(I, Jv)], Bc[Jv], tb);
            CBc[I] = tb;
          }
          __syncwarp();  // previous Gauss point's CB fully consumed
          if (lane < ND)
#pragma unroll
            for (int I = 0; I < 6; ++I) sm[L::CB + I + 6 * lane] = CBc[I];
          __syncwarp();
          // elastic: K_dd(r, c) += B_r . (w C B_c); geometric: K_dd(3i+b, 3j+b) += G_i.(w S).G_j
          double gi[3];
#pragma unroll
          for (int d = 0; d < 3; ++d)
            gi[d] = G0 * gp[L::GP_S + vg(0, d)] + G1 * gp[L::GP_S + vg(1, d)] + G2 * gp[L::GP_S + vg(2, d)];
#pragma unroll
          for (int c = 0; c < ND; ++c)
          {
            const d2* cb2 = reinterpret_cast<const d2*>(sm + L::CB + 6 * c);
            const d2 q0 = cb2[0], q1 = cb2[1], q2 = cb2[2];
            // two independent chains per column
            const double t = fma(Bc[2], q1.x, fma(Bc[1], q0.y, Bc[0] * q0.x));
            const double v = fma(Bc[5], q2.y, fma(Bc[4], q2.x, Bc[3] * q1.y));
            Kdd[c] += t + v;
          }
#pragma unroll
          for (int j = 0; j < NN; ++j)
          {
            const double* gj = gp + L::GP_G + 3 * j;
            const double sg = gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2];
#pragma unroll
            for (int b = 0; b < 3; ++b) Kdd[3 * j + b] += (b == ei) ? sg : 0.0;  // select, no divergent branch
          }
        }
#pragma unroll
        for (int c = 0; c < ND; ++c)
        {
          const double* kc = KdaS + c * NEAS;
          double t = 0.0;
#pragma unroll
          for (int a = 0; a < NEAS; ++a) t = fma(mK[a], kc[a], t);
          Kdd[c] += t;
        }
      }
      __syncwarp();  // every lane is done with CB / GP: the K_e staging overwrites them
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
              for (int b = 0; b < 3; ++b) v[i][b] = Ke[(3 * (h + i) + b) * ND + lane];
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

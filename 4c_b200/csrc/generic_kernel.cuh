// generic_kernel.cuh -- general evaluate + assemble for constitutive tensors WITHOUT major symmetry.
//
// The fast element kernels (hex8 / tet10 / hex27) carry the 21 unique entries of a major-symmetric tangent and exploit
// the symmetry of K_e.  A material whose algorithmic tangent is not symmetric -- the non-associative consistent tangent
// of MAT_Struct_DruckerPrager (ETA != ETABAR, src/mat/4C_mat_plasticdruckerprager.cpp:127-150) -- gives a
// non-symmetric K_e = sum_gp B^T C B w, which the reference assembles entry by entry like any other
// (add_stiffness_matrix, src/solid_ele/4C_solid_ele_calc_lib.hpp:1124-1151; linear kinematics:
// 4C_solid_ele_calc_displacement_based_linear_kinematics.hpp:133-150).  This kernel is the plain route for such blocks:
// one thread per (element, row node i) computes the whole block row K_i,: (3 x 3 nn) and f_i and adds them into the CSR
// rows of node i.  Table driven (any cell with nn <= NN nodes), total Lagrangean or linear kinematics; throughput is
// not the point here (every thread of an element repeats the Gauss-point work), correctness and coverage are.
#pragma once

#include "hex8_kernel.cuh"
#include "solid_math.cuh"

namespace b200s
{
  struct GenericArgs
  {
    EvalArgs A;             // mesh, state, graph, outputs as for the fast kernels (pointers already offset)
    const double* dN_gp;    // [ngp][nn][3]
    const double* dN_node;  // [nn][nn][3] (Jacobian check) or nullptr
    const double* w;        // [ngp]
    const double* dN_cen;   // [nn][3] at the element centroid: TECH fbar (nullptr otherwise)
    const double* dN_ans;   // [8][nn][3] at the ANS sampling points: TECH shell_ans on hex8 (nullptr otherwise)
    const double* xi_gp;    // [ngp][3] Gauss point coordinates (ANS interpolation weights)
    int nn, ngp, pos_stride, stiff;
    int hist_pts;           // MULF histories per element: ngp, or ngp + 1 with TECH fbar (the centroid's follows the points')
    double* eas_hist;       // TECH shell_eas_ans: EAS history of the block's elements (alpha_inc | alpha | s | invKaa | Kda,
                            // hex8eas::hist_size(7) doubles per element), nullptr otherwise
    const double* eas_t0;   // ... and the centroid transformation T0^-T[36], detJ0 (stride ANS_EAS_T0_STRIDE)
  };

  // ---- TECH shell_eas_ans on hex8: ShellEASANSFormulation<hex8, eastype_sh8_7> (4C_solid_ele_calc_shell_eas_ans.hpp:70-343).
  // The thread of row node i integrates K_aa, s (the same numbers in the eight threads of an element) and its three rows of
  // K_da next to the ANS block row; the rows meet in the element's history (which the reference keeps too, for recover),
  // then every thread condenses its block row: K_i: -= K_da,i K_aa^-1 K_da^T, f_i -= K_da,i K_aa^-1 s.
  constexpr int ANS_NEAS = 7;
  constexpr int ANS_EAS_HIST = 3 * ANS_NEAS + ANS_NEAS * ANS_NEAS + 24 * ANS_NEAS;
  constexpr int ANS_EAS_T0_STRIDE = 40;
  struct AnsEas
  {
    double Kaa[ANS_NEAS * ANS_NEAS], s[ANS_NEAS], Kda[3][ANS_NEAS];
  };

  // in-place inverse by Gauss-Jordan with partial pivoting (the reference: LAPACK getrf / getri through
  // solve_for_inverse_ignoring_errors, shell_eas_ans.hpp:231)
  template <int N>
  B2_HD void dense_inverse_small(double* A /*r + N c*/)
  {
    int perm[N];
    for (int k = 0; k < N; ++k) perm[k] = k;
    for (int k = 0; k < N; ++k)
    {
      int p = k;
      double best = fabs(A[k + N * k]);
      for (int r = k + 1; r < N; ++r)
        if (fabs(A[r + N * k]) > best) best = fabs(A[r + N * k]), p = r;
      if (p != k)
      {
        for (int c = 0; c < N; ++c)
        {
          const double t = A[k + N * c];
          A[k + N * c] = A[p + N * c], A[p + N * c] = t;
        }
        const int t = perm[k];
        perm[k] = perm[p], perm[p] = t;
      }
      const double piv = 1.0 / A[k + N * k];
      A[k + N * k] = 1.0;
      for (int c = 0; c < N; ++c) A[k + N * c] *= piv;
      for (int r = 0; r < N; ++r)
      {
        if (r == k) continue;
        const double f = A[r + N * k];
        A[r + N * k] = 0.0;
        for (int c = 0; c < N; ++c) A[r + N * c] = fma(-f, A[k + N * c], A[r + N * c]);
      }
    }
    // the row exchanges of A become column exchanges of the inverse: column perm[k] of A^-1 is column k of the result
    double tmp[N * N];
    for (int k = 0; k < N; ++k)
      for (int r = 0; r < N; ++r) tmp[r + N * perm[k]] = A[r + N * k];
    for (int q = 0; q < N * N; ++q) A[q] = tmp[q];
  }

  // thread i of the element: its K_da rows into the history; thread 0 also s and K_aa^-1 (reset_condensed_variable_integration
  // + integrate_eas + prepare_condensed_contribution).  ea.Kaa holds the inverse afterwards.
  B2_HD void ans_eas_store(double* h, int i, AnsEas& ea)
  {
    dense_inverse_small<ANS_NEAS>(ea.Kaa);
    double* Kda = h + 3 * ANS_NEAS + ANS_NEAS * ANS_NEAS;
    for (int d = 0; d < 3; ++d)
      for (int c = 0; c < ANS_NEAS; ++c) Kda[(3 * i + d) + 24 * c] = ea.Kda[d][c];
    if (i != 0) return;
    for (int c = 0; c < ANS_NEAS; ++c) h[2 * ANS_NEAS + c] = ea.s[c];
    for (int q = 0; q < ANS_NEAS * ANS_NEAS; ++q) h[3 * ANS_NEAS + q] = ea.Kaa[q];
  }

  // add_condensed_contribution_to_force_vector / _to_stiffness_matrix (shell_eas_ans.hpp:300-320) for block row i; every
  // K_da row of the element is in the history by now
  B2_HD void ans_eas_condense(const double* h, int i, const AnsEas& ea, bool stiff, double (*Krow)[9], double* fi)
  {
    const double* Kda = h + 3 * ANS_NEAS + ANS_NEAS * ANS_NEAS;
    double m[3][ANS_NEAS];  // -K_da,i K_aa^-1
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < ANS_NEAS; ++c)
      {
        double v = 0.0;
        for (int k = 0; k < ANS_NEAS; ++k) v = fma(ea.Kda[b][k], ea.Kaa[k + ANS_NEAS * c], v);
        m[b][c] = -v;
      }
    for (int b = 0; b < 3; ++b)
    {
      double v = 0.0;
      for (int c = 0; c < ANS_NEAS; ++c) v = fma(m[b][c], ea.s[c], v);
      fi[b] += v;
    }
    if (!stiff) return;
    for (int j = 0; j < 8; ++j)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c)
        {
          double v = 0.0;
          for (int k = 0; k < ANS_NEAS; ++k) v = fma(m[b][k], Kda[(3 * j + c) + 24 * k], v);
          Krow[j][b + 3 * c] += v;
        }
  }

  // S[6], Cm36[I + 6 J] of the block's material at one point (C = tensor components 00,11,22,01,12,02)
  template <int MAT>
  B2_HD void material_evaluate_full(const MatParams& mp, const double* C, double* S, double* Cm36)
  {
    if (MAT == MAT_DP)
      drucker_prager_evaluate_t<true>(mp.ext, mp.hist, C, S, Cm36);
    else
    {
      double Cm[21];
      material_evaluate<MAT>(mp, C, S, Cm);
#pragma unroll
      for (int I = 0; I < 6; ++I)
#pragma unroll
        for (int J = 0; J < 6; ++J) Cm36[I + 6 * J] = Cm[c21(I < J ? I : J, I < J ? J : I)];
    }
  }

  // block row i of K_e and f_i of element e (device element index); Krow[j][b + 3 c] = K_ij(b, c)
  template <int MAT, int NN>
  B2_HD bool generic_block_row(const GenericArgs& Ga, int64_t e, int i, double (*Krow)[9], double* fi, double& detJ_min,
      double& detF_min)
  {
    const EvalArgs& A = Ga.A;
    const int nn = Ga.nn;
    constexpr int vg[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};
    double X[NN][3], u[NN][3], G[NN][3];
    const int32_t* ce = A.conn + e * nn;
    for (int k = 0; k < nn; ++k)
      for (int d = 0; d < 3; ++d) X[k][d] = A.X[3 * (int64_t)ce[k] + d], u[k][d] = A.u[3 * (int64_t)ce[k] + d];
    for (int j = 0; j < nn; ++j)
      for (int q = 0; q < 9; ++q) Krow[j][q] = 0.0;
    fi[0] = fi[1] = fi[2] = 0.0;
    bool neg = false;
    // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562): the thread of row node 0 checks
    // every node of the element, so that offending ELEMENTS are counted once (as the fast kernels and the oracle do)
    if (Ga.dN_node && i == 0)
    {
      for (int q = 0; q < nn; ++q)
      {
        double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const double* dN = Ga.dN_node + (size_t)q * nn * 3;
        for (int k = 0; k < nn; ++k)
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(dN[3 * k + a], X[k][b], J[a + 3 * b]);
        const double dj = det3(J);
        detJ_min = fmin(detJ_min, dj);
        neg = neg || !(dj > 0.0);
      }
    }
    // TECH fbar, FBarFormulation::prepare (4C_solid_ele_calc_fbar.hpp:58-66; MulfFBarFormulation::prepare,
    // 4C_solid_ele_calc_mulf_fbar.hpp:118-149): Jacobian mapping and deformation gradient at the centroid;
    // H0[k][c] = sum_d F0^-1(d,c) N_XYZ_0(d,k) is the centroid half of the H-operator (lib_fbar.hpp:71-104)
    const bool fbar = Ga.dN_cen != nullptr;
    double detF0 = 1.0, H0[NN][3];
    if (fbar)
    {
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ[9], F0[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, iF0[9];
      for (int k = 0; k < nn; ++k)
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(Ga.dN_cen[3 * k + a], X[k][b], J[a + 3 * b]);
      inv3(iJ, J);
      for (int k = 0; k < nn; ++k)
      {
        for (int a = 0; a < 3; ++a)
          G[k][a] = fma(iJ[a + 6], Ga.dN_cen[3 * k + 2], fma(iJ[a + 3], Ga.dN_cen[3 * k + 1], iJ[a] * Ga.dN_cen[3 * k]));
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) F0[a + 3 * b] = fma(u[k][a], G[k][b], F0[a + 3 * b]);
      }
      if (A.gp_hist != nullptr)
        mulf_deformation_gradient<NN>(A.gp_hist + (e * Ga.hist_pts + Ga.ngp) * MULF_HIST, Ga.dN_cen, &u[0][0], F0);
      detF0 = inv3(iF0, F0);
      for (int k = 0; k < nn; ++k)
        for (int c = 0; c < 3; ++c) H0[k][c] = iF0[3 * c] * G[k][0] + iF0[1 + 3 * c] * G[k][1] + iF0[2 + 3 * c] * G[k][2];
    }
    for (int g = 0; g < Ga.ngp; ++g)
    {
      const double* dN = Ga.dN_gp + (size_t)g * nn * 3;
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ[9];
      for (int k = 0; k < nn; ++k)
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) J[a + 3 * b] = fma(dN[3 * k + a], X[k][b], J[a + 3 * b]);
      const double detJ = inv3(iJ, J);
      detJ_min = fmin(detJ_min, detJ);
      double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      for (int k = 0; k < nn; ++k)
      {
        for (int a = 0; a < 3; ++a)
          G[k][a] = fma(iJ[a + 6], dN[3 * k + 2], fma(iJ[a + 3], dN[3 * k + 1], iJ[a] * dN[3 * k]));
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) F[a + 3 * b] = fma(u[k][a], G[k][b], F[a + 3 * b]);
      }
      // PRESTRESS_TECH mulf: F = (I + u (J_hist^-1 dN)^T) F_hist (4C_solid_ele_calc_lib_mulf.hpp:41-90); everything else
      // of the formulation is the displacement-based one
      if (A.gp_hist != nullptr) mulf_deformation_gradient<NN>(A.gp_hist + (e * Ga.hist_pts + g) * MULF_HIST, dN, &u[0][0], F);
      double C[6], S[6], Cm[36], iF[9], Cu[6], fb = 1.0;
      const double detF = fbar ? inv3(iF, F) : det3(F);
      detF_min = fmin(detF_min, detF);
      right_cauchy_green(F, C);
      if (fbar)
      {
        // evaluate_fbar_factor (lib_fbar.hpp:47-51); the material sees F_bar = fb F, i.e. C_bar = fb^2 C (fbar.hpp:103-112)
        fb = pow(detF0 / detF, 1.0 / 3.0);
        double Fb[9];
        for (int q = 0; q < 9; ++q) Fb[q] = fb * F[q];
        for (int q = 0; q < 6; ++q) Cu[q] = C[q];
        right_cauchy_green(Fb, C);
      }
      if (A.linear) linear_kinematics(F, C);  // E = sym(grad u), F = I
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr)
        mp.hist = A.mat_hist + (e * Ga.ngp + g) * material_hist_doubles(MAT);  // every thread of the element writes the same state
      material_evaluate_full<MAT>(mp, C, S, Cm);
      const double w0 = detJ * Ga.w[g];
      // F-bar: force and geometric stiffness with fac / fbar, elastic stiffness with fac * fbar (fbar.hpp:152-176)
      const double w = w0 * fb, ws = w0 / fb;
      // f_i(b) += ws sum_a (F S)(b, a) G_i(a)
      double bS[3];
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int a = 0; a < 3; ++a)
        {
          double fs = 0.0;
          for (int c = 0; c < 3; ++c) fs = fma(F[b + 3 * c], S[vg[c][a]], fs);
          v = fma(fs, G[i][a], v);
        }
        bS[b] = v;
        fi[b] = fma(ws, v, fi[b]);
      }
      if (!Ga.stiff) continue;
      // T1[b][f][d] = sum_{a,e} G_i(a) F(b,e) C[(a e)][(f d)];  V[b][c][d] = w (sum_f T1[b][f][d] F(c,f) + delta_bc sum_a G_i(a) S(a,d))
      double V[3][3][3];
      for (int b = 0; b < 3; ++b)
      {
        double T1[3][3];
        for (int f = 0; f < 3; ++f)
          for (int d = 0; d < 3; ++d)
          {
            double v = 0.0;
            for (int a = 0; a < 3; ++a)
              for (int ee = 0; ee < 3; ++ee) v = fma(G[i][a] * F[b + 3 * ee], Cm[vg[a][ee] + 6 * vg[f][d]], v);
            T1[f][d] = v;
          }
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d)
            V[b][c][d] = w * (T1[0][d] * F[c] + T1[1][d] * F[c + 3] + T1[2][d] * F[c + 6]);
      }
      if (!A.linear)
        for (int d = 0; d < 3; ++d)
        {
          const double sg = ws * (G[i][0] * S[vg[0][d]] + G[i][1] * S[vg[1][d]] + G[i][2] * S[vg[2][d]]);
          for (int b = 0; b < 3; ++b) V[b][b][d] += sg;
        }
      for (int j = 0; j < nn; ++j)
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c)
            Krow[j][b + 3 * c] = fma(V[b][c][2], G[j][2], fma(V[b][c][1], G[j][1], fma(V[b][c][0], G[j][0], Krow[j][b + 3 * c])));
      if (fbar)
      {
        // add_fbar_stiffness_matrix (lib_fbar.hpp:121-160): K_ij(b,c) += Hop_j(c) (B_i^T ccg fac fbar / 3 - B_i^T S fac / fbar / 3)(b),
        // ccg = cmat : C with the Cauchy-Green tensor of the unscaled F, Hop_j(c) = H0_j(c) - sum_d F^-1(d,c) N_XYZ(d,j)
        double ccg[6];
        for (int I = 0; I < 6; ++I)
        {
          double v = 0.0;
          for (int J2 = 0; J2 < 6; ++J2) v = fma(Cm[I + 6 * J2] * (J2 < 3 ? 1.0 : 2.0), Cu[J2], v);
          ccg[I] = v;
        }
        double t[3];
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int a = 0; a < 3; ++a)
          {
            double fc = 0.0;
            for (int c = 0; c < 3; ++c) fc = fma(F[b + 3 * c], ccg[vg[c][a]], fc);
            v = fma(fc, G[i][a], v);
          }
          t[b] = (w / 3.0) * v - (ws / 3.0) * bS[b];
        }
        for (int j = 0; j < nn; ++j)
          for (int c = 0; c < 3; ++c)
          {
            const double hop = H0[j][c] - (iF[3 * c] * G[j][0] + iF[1 + 3 * c] * G[j][1] + iF[2 + 3 * c] * G[j][2]);
            for (int b = 0; b < 3; ++b) Krow[j][b + 3 * c] = fma(hop, t[b], Krow[j][b + 3 * c]);
          }
      }
    }
    return neg;
  }


  // ---- TECH shell_ans on hex8: ShellANSFormulation (src/solid_ele/4C_solid_ele_calc_shell_ans.hpp:52-196,
  // 4C_solid_ele_calc_shell_ans_lib.hpp).  Strains, B-operator and geometric stiffness are formed in the parametric frame;
  // the transverse normal row tt and the transverse shear rows st, rt are interpolated from the sampling points
  // A..D (shears) and E..H (tt) of the mid-surface, T^-T of the Gauss point's Jacobian takes them to the global frame.
  // T(J) is the strain-like Voigt transformation of E' = J E J^T, so T(J)^-1 = T(J^-1) -- formed directly, where the
  // reference inverts numerically (evaluate_voigt_transformation_matrix, 4C_solid_ele_calc_lib.hpp:1451-1517).
  // (voigt_strain_transformation: solid_math.cuh)

  // local (parametric) B-operator columns of node k, dofs d = 0..2: Bl[r][d]; Jc = current Jacobian at the Gauss point,
  // Jcs = current Jacobians at the eight sampling points, dNs = shape derivatives there, w = interpolation weights
  B2_HD void ans_local_bop_node(const double* dN, const double* dNs, int nn, int k, const double* Jc, const double (*Jcs)[9],
      const double* w_tt, const double* w_st, const double* w_rt, double (*Bl)[3])
  {
    const double* D = dN + 3 * k;
#pragma unroll
    for (int d = 0; d < 3; ++d)
    {
      Bl[0][d] = D[0] * Jc[0 + 3 * d];
      Bl[1][d] = D[1] * Jc[1 + 3 * d];
      Bl[3][d] = D[0] * Jc[1 + 3 * d] + D[1] * Jc[0 + 3 * d];
      double tt = 0.0, st = 0.0, rt = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m)
      {
        const double* Ds = dNs + ((size_t)(4 + m) * nn + k) * 3;  // sampling points E..H
        tt = fma(w_tt[m], Ds[2] * Jcs[4 + m][2 + 3 * d], tt);
      }
#pragma unroll
      for (int m = 0; m < 2; ++m)
      {
        const double* Db = dNs + ((size_t)(1 + 2 * m) * nn + k) * 3;  // B, D
        st = fma(w_st[m], Db[1] * Jcs[1 + 2 * m][2 + 3 * d] + Db[2] * Jcs[1 + 2 * m][1 + 3 * d], st);
        const double* Da = dNs + ((size_t)(2 * m) * nn + k) * 3;  // A, C
        rt = fma(w_rt[m], Da[0] * Jcs[2 * m][2 + 3 * d] + Da[2] * Jcs[2 * m][0 + 3 * d], rt);
      }
      Bl[2][d] = tt, Bl[4][d] = st, Bl[5][d] = rt;
    }
  }

  template <int MAT>
  B2_HD bool generic_block_row_ans(const GenericArgs& Ga, int64_t e, int i, double (*Krow)[9], double* fi, double& detJ_min,
      double& detF_min, AnsEas* ea = nullptr)
  {
    constexpr int NN = 8;
    if (ea)
    {
      for (double& v : ea->Kaa) v = 0.0;
      for (double& v : ea->s) v = 0.0;
      for (int d = 0; d < 3; ++d)
        for (int c = 0; c < ANS_NEAS; ++c) ea->Kda[d][c] = 0.0;
    }
    const EvalArgs& A = Ga.A;
    double X[NN][3], x[NN][3];
    const int32_t* ce = A.conn + e * NN;
    for (int k = 0; k < NN; ++k)
      for (int d = 0; d < 3; ++d)
      {
        X[k][d] = A.X[3 * (int64_t)ce[k] + d];
        x[k][d] = X[k][d] + A.u[3 * (int64_t)ce[k] + d];
      }
    for (int j = 0; j < NN; ++j)
      for (int q = 0; q < 9; ++q) Krow[j][q] = 0.0;
    fi[0] = fi[1] = fi[2] = 0.0;
    auto jac = [&](const double* dN, const double (*P)[3], double* J)
    {
      for (int q = 0; q < 9; ++q) J[q] = 0.0;
      for (int k = 0; k < NN; ++k)
        for (int a = 0; a < 3; ++a)
          for (int d = 0; d < 3; ++d) J[a + 3 * d] = fma(dN[3 * k + a], P[k][d], J[a + 3 * d]);
    };
    bool neg = false;
    if (Ga.dN_node && i == 0)
      for (int q = 0; q < NN; ++q)
      {
        double J[9];
        jac(Ga.dN_node + (size_t)q * NN * 3, X, J);
        const double dj = det3(J);
        detJ_min = fmin(detJ_min, dj);
        neg = neg || !(dj > 0.0);
      }
    // evaluate_sampling_point_data (shell_ans_lib.hpp:515-571)
    double Jrs[8][9], Jcs[8][9];
    for (int q = 0; q < 8; ++q)
    {
      jac(Ga.dN_ans + (size_t)q * NN * 3, X, Jrs[q]);
      jac(Ga.dN_ans + (size_t)q * NN * 3, x, Jcs[q]);
    }
    auto rows = [](const double* J, int a, int b) { return J[a] * J[b] + J[a + 3] * J[b + 3] + J[a + 6] * J[b + 6]; };
    for (int g = 0; g < Ga.ngp; ++g)
    {
      const double* dN = Ga.dN_gp + (size_t)g * NN * 3;
      const double* xi = Ga.xi_gp + 3 * g;
      double J[9], Jc[9], iJ[9], T[36];
      jac(dN, X, J), jac(dN, x, Jc);
      const double detJ = inv3(iJ, J);
      detJ_min = fmin(detJ_min, detJ);
      detF_min = fmin(detF_min, det3(Jc) / detJ);
      voigt_strain_transformation(iJ, T);
      const double w_tt[4] = {0.25 * (1 - xi[0]) * (1 - xi[1]), 0.25 * (1 + xi[0]) * (1 - xi[1]), 0.25 * (1 + xi[0]) * (1 + xi[1]),
          0.25 * (1 - xi[0]) * (1 + xi[1])};
      const double w_st[2] = {0.5 * (1 + xi[0]), 0.5 * (1 - xi[0])}, w_rt[2] = {0.5 * (1 - xi[1]), 0.5 * (1 + xi[1])};
      // evaluate_local_glstrain (:169-291), to the global frame, shear rows halved (strain-like -> tensor components)
      double El[6], C[6];
      El[0] = 0.5 * (rows(Jc, 0, 0) - rows(J, 0, 0));
      El[1] = 0.5 * (rows(Jc, 1, 1) - rows(J, 1, 1));
      El[3] = rows(Jc, 0, 1) - rows(J, 0, 1);
      double tt = 0.0, st = 0.0, rt = 0.0;
      for (int m = 0; m < 4; ++m) tt = fma(w_tt[m], rows(Jcs[4 + m], 2, 2) - rows(Jrs[4 + m], 2, 2), tt);
      for (int m = 0; m < 2; ++m)
      {
        st = fma(w_st[m], rows(Jcs[1 + 2 * m], 1, 2) - rows(Jrs[1 + 2 * m], 1, 2), st);
        rt = fma(w_rt[m], rows(Jcs[2 * m], 0, 2) - rows(Jrs[2 * m], 0, 2), rt);
      }
      El[2] = 0.5 * tt, El[4] = st, El[5] = rt;
      // TECH shell_eas_ans: m_tilde = detJ0 / detJ T0^-T M(xi) with the seven modes r | s | t, rt, st | r, s in the rows
      // rr | ss | tt | rs (eastype_sh8_7, eas_helpers.hpp:330-357, :438-453); strain-like rows
      double Mt[6][ANS_NEAS];
      if (ea)
      {
        const double* t0 = Ga.eas_t0 + e * ANS_EAS_T0_STRIDE;
        constexpr int rows7[ANS_NEAS] = {0, 1, 2, 3, 3, 2, 2};
        const double mono7[ANS_NEAS] = {xi[0], xi[1], xi[2], xi[0], xi[1], xi[0] * xi[2], xi[1] * xi[2]};
        const double sc = t0[36] / detJ;
        for (int c = 0; c < ANS_NEAS; ++c)
          for (int r = 0; r < 6; ++r) Mt[r][c] = sc * t0[r + 6 * rows7[c]] * mono7[c];
      }
      const double* alpha = ea ? Ga.eas_hist + e * ANS_EAS_HIST + ANS_NEAS : nullptr;
      for (int r = 0; r < 6; ++r)
      {
        double v = 0.0;
        for (int q = 0; q < 6; ++q) v = fma(T[r + 6 * q], El[q], v);
        if (ea)  // gl_strain_eas = gl_strain + Mtilde alpha (shell_eas_ans.hpp:129-131)
          for (int c = 0; c < ANS_NEAS; ++c) v = fma(Mt[r][c], alpha[c], v);
        C[r] = r < 3 ? 2.0 * v + 1.0 : v;  // C = 2 E + I in tensor components (shear: 2 * (v / 2))
      }
      double S[6], Cm[36];
      MatParams mp = A.mp;
      if (A.per_ele_param != nullptr) mp.p[0] = A.per_ele_param[e];
      if (MAT == MAT_HYPER && A.fibers != nullptr) mp.fiber = A.fibers + 3 * e;
      if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr)
        mp.hist = A.mat_hist + (e * Ga.ngp + g) * material_hist_doubles(MAT);
      material_evaluate_full<MAT>(mp, C, S, Cm);
      const double w = detJ * Ga.w[g];
      // B_i = T^-T Bl_i (6 x 3), f_i += w B_i^T S, A_i = w B_i^T C (3 x 6)
      double Bl[6][3], Bi[6][3], Ai[3][6];
      ans_local_bop_node(dN, Ga.dN_ans, NN, i, Jc, Jcs, w_tt, w_st, w_rt, Bl);
      for (int r = 0; r < 6; ++r)
        for (int d = 0; d < 3; ++d)
        {
          double v = 0.0;
          for (int q = 0; q < 6; ++q) v = fma(T[r + 6 * q], Bl[q][d], v);
          Bi[r][d] = v;
        }
      for (int d = 0; d < 3; ++d)
      {
        double v = 0.0;
        for (int r = 0; r < 6; ++r) v = fma(Bi[r][d], S[r], v);
        fi[d] = fma(w, v, fi[d]);
      }
      if (ea)
      {
        // integrate_eas (eas_helpers.hpp:352-371) with the ANS B-operator: K_aa += w Mt^T C Mt, K_da,i += w B_i^T C Mt,
        // s += w Mt^T S -- for both actions, as the reference does
        double cM[6][ANS_NEAS];
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < ANS_NEAS; ++c)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v = fma(Cm[r + 6 * k], Mt[k][c], v);
            cM[r][c] = w * v;
          }
        for (int c = 0; c < ANS_NEAS; ++c)
        {
          for (int r = 0; r < ANS_NEAS; ++r)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v = fma(Mt[k][r], cM[k][c], v);
            ea->Kaa[r + ANS_NEAS * c] += v;
          }
          for (int d = 0; d < 3; ++d)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v = fma(Bi[k][d], cM[k][c], v);
            ea->Kda[d][c] += v;
          }
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v = fma(Mt[k][c], S[k], v);
          ea->s[c] = fma(w, v, ea->s[c]);
        }
      }
      if (!Ga.stiff) continue;
      for (int b = 0; b < 3; ++b)
        for (int W = 0; W < 6; ++W)
        {
          double v = 0.0;
          for (int V = 0; V < 6; ++V) v = fma(Bi[V][b], Cm[V + 6 * W], v);
          Ai[b][W] = w * v;
        }
      // Sl = T^-T^T S: the geometric term w S . (T^-T G_ij) = w Sl . G_ij (add_ans_geometric_stiffness, :383-454)
      double Sl[6];
      for (int q = 0; q < 6; ++q)
      {
        double v = 0.0;
        for (int r = 0; r < 6; ++r) v = fma(T[r + 6 * q], S[r], v);
        Sl[q] = w * v;
      }
      const double* Di = dN + 3 * i;
      for (int j = 0; j < NN; ++j)
      {
        double Bj[6][3];
        ans_local_bop_node(dN, Ga.dN_ans, NN, j, Jc, Jcs, w_tt, w_st, w_rt, Bl);
        for (int r = 0; r < 6; ++r)
          for (int d = 0; d < 3; ++d)
          {
            double v = 0.0;
            for (int q = 0; q < 6; ++q) v = fma(T[r + 6 * q], Bl[q][d], v);
            Bj[r][d] = v;
          }
        const double* Dj = dN + 3 * j;
        double G[6];
        G[0] = Di[0] * Dj[0], G[1] = Di[1] * Dj[1], G[3] = Di[0] * Dj[1] + Di[1] * Dj[0];
        double gtt = 0.0, gst = 0.0, grt = 0.0;
        for (int m = 0; m < 4; ++m)
          gtt = fma(w_tt[m], Ga.dN_ans[((size_t)(4 + m) * NN + i) * 3 + 2] * Ga.dN_ans[((size_t)(4 + m) * NN + j) * 3 + 2], gtt);
        for (int m = 0; m < 2; ++m)
        {
          const double* Bi_ = Ga.dN_ans + ((size_t)(1 + 2 * m) * NN + i) * 3;
          const double* Bj_ = Ga.dN_ans + ((size_t)(1 + 2 * m) * NN + j) * 3;
          gst = fma(w_st[m], Bi_[1] * Bj_[2] + Bi_[2] * Bj_[1], gst);
          const double* Ai_ = Ga.dN_ans + ((size_t)(2 * m) * NN + i) * 3;
          const double* Aj_ = Ga.dN_ans + ((size_t)(2 * m) * NN + j) * 3;
          grt = fma(w_rt[m], Ai_[0] * Aj_[2] + Ai_[2] * Aj_[0], grt);
        }
        G[2] = gtt, G[4] = gst, G[5] = grt;
        double Gij = 0.0;
        for (int q = 0; q < 6; ++q) Gij = fma(Sl[q], G[q], Gij);
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c)
          {
            double v = (b == c) ? Gij : 0.0;
            for (int W = 0; W < 6; ++W) v = fma(Ai[b][W], Bj[W][c], v);
            Krow[j][b + 3 * c] += v;
          }
      }
    }
    return neg;
  }

#if defined(__CUDACC__)
  __device__ __forceinline__ void atomic_min_double_generic(double* addr, double v)
  {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v)
    {
      const unsigned long long assumed = old;
      old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
      if (old == assumed) break;
    }
  }

  template <int MAT, int NN>
  __global__ void __launch_bounds__(64) k_generic_eval_assemble(const GenericArgs Ga)
  {
    const EvalArgs& A = Ga.A;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / Ga.nn;
    const int i = (int)(t - e * Ga.nn);
    if (e >= A.n_ele) return;
    double Krow[NN][9], fi[3], dj = 1e300, df = 1e300;
    bool neg;
    if constexpr (NN == 8)
    {
      if (Ga.dN_ans != nullptr && Ga.eas_hist != nullptr)
      {
        // TECH shell_eas_ans: the eight threads of an element are eight adjacent lanes of one warp (blockDim 64, nn 8)
        AnsEas ea;
        neg = generic_block_row_ans<MAT>(Ga, e, i, Krow, fi, dj, df, &ea);
        double* h = Ga.eas_hist + e * ANS_EAS_HIST;
        const unsigned mask = 0xFFu << (threadIdx.x & 24);
        ans_eas_store(h, i, ea);
        __syncwarp(mask);
        ans_eas_condense(h, i, ea, Ga.stiff != 0, Krow, fi);
      }
      else
        neg = Ga.dN_ans != nullptr ? generic_block_row_ans<MAT>(Ga, e, i, Krow, fi, dj, df)
                                   : generic_block_row<MAT, NN>(Ga, e, i, Krow, fi, dj, df);
    }
    else
      neg = generic_block_row<MAT, NN>(Ga, e, i, Krow, fi, dj, df);
    atomic_min_double_generic(&A.err->min_detJ, dj);
    atomic_min_double_generic(&A.err->min_detF, df);
    if (neg)
    {
      const long long id = A.ele_id ? (long long)A.ele_id[e] : (long long)e;
      atomicMin(&A.err->first_bad, id);
      atomicAdd(&A.err->n_neg_jac, 1ull);
    }
    const NodeRow nr = A.node_row[A.conn[e * Ga.nn + i]];
    if (nr.rowstart < 0) return;  // row node not owned (lmrowowner != myrank)
    for (int b = 0; b < 3; ++b) atomicAdd(A.f + nr.frow + b, fi[b]);
    if (!Ga.stiff) return;
    const int32_t* pos = A.pos + e * Ga.pos_stride + i * Ga.nn;
    for (int j = 0; j < Ga.nn; ++j)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c)
          atomicAdd(A.vals + nr.rowstart + (int64_t)b * nr.rowlen + pos[j] + c, Krow[j][b + 3 * c]);
  }
#endif
}  // namespace b200s

// post_kernels.cuh -- the element actions around the force/stiffness hot path:
//   struct_calc_stress   SolidEleCalc::calculate_stress          (4C_solid_ele_calc.cpp:416-476)
//   struct_calc_energy   SolidEleCalc::calculate_internal_energy (4C_solid_ele_calc.cpp:365-414)
//   mass matrix / inertia force of struct_calc_nlnstiffmass, struct_calc_nlnstifflmass and
//   struct_calc_internalinertiaforce (4C_solid_ele_evaluate.cpp:116-159; add_mass_matrix,
//   4C_solid_ele_calc_lib.hpp:1221-1238; the branch for a mass rule that differs from the stiffness
//   rule, 4C_solid_ele_calc.cpp:151-153,228-236,265-280)
// None of them is per-Newton-iteration work; they are written for simplicity: one thread per
// (element, Gauss point) resp. (element, row node), generic in the cell type through tables in
// global memory.  The point functions are __host__ __device__ so the CPU suite can check them.
#pragma once

#include <stdint.h>

#include "hex8_kernel.cuh"
#include "solid_math.cuh"

namespace b200s
{
  enum StressType  // FourC::Solid::StressType
  {
    STRESS_NONE = 0,
    STRESS_2PK = 1,
    STRESS_CAUCHY = 2
  };
  enum StrainType  // FourC::Solid::StrainType (strain_log needs an eigen decomposition: CPU path)
  {
    STRAIN_NONE = 0,
    STRAIN_GL = 1,
    STRAIN_EA = 2
  };

  struct PostArgs
  {
    const double* X;              // 3*n_col_nodes
    const double* u;              // 3*n_col_nodes (column map)
    const int32_t* conn;          // n_ele*nn, device element order
    const double* per_ele_param;  // nullptr or n_ele (device order)
    const double* fibers;         // nullptr or 3*n_ele unit fibre vectors (device order)
    const int32_t* ele_id;        // nullptr or device element -> caller's column element
    const uint8_t* ele_is_row;    // nullptr (all) or n_ele flags in the CALLER's numbering
    const double* dN_gp;          // [ngp][nn][3]
    const double* dN_cen;         // [nn][3] (F-bar)
    const double* w;              // [ngp]
    const double* gp_hist;        // nullptr or MULF history (device element order), hist_pts points per element
    int hist_pts;                 // ngp, or ngp + 1 for MULF + F-bar (the centroid history follows the Gauss points)
    double* mat_hist;         // history material: n_ele * ngp * DP_HIST (device element order) or nullptr
    // EAS hex8 (eas_neas = 9 / 21, else 0): element history (alpha at [neas..2 neas)), the precomputed
    // centroid transformation (T0invT[36], detJ0; stride eas_t0_stride) and the Gauss point coordinates
    int eas_neas, eas_hist_size, eas_t0_stride;
    const double* eas_hist;
    const double* eas_t0;
    const double* xi_gp;          // [ngp][3]
    const double* dN_ans;         // TECH shell_ans / shell_eas_ans (hex8): [8][nn][3] at the ANS sampling points, else nullptr
    MatParams mp;
    int64_t n_ele;
    int nn, ngp;
    int fbar, linear;
    int stress_type, strain_type;
    double* stress;  // [caller element][gp][6] or nullptr
    double* strain;  // same
    double* energy;  // one double, accumulated over row elements, or nullptr
  };

  // J, det J and F at one parametric point from the tabulated dN[k*3+a]
  // (evaluate_jacobian_mapping / evaluate_deformation_gradient, calc_lib.hpp:499-516,646-660)
  B2_HD double point_kinematics(int nn, const double* dN, const double* X, const double* u,
      const int32_t* ce, double* F)
  {
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ[9];
    for (int k = 0; k < nn; ++k)
    {
      const double* x = X + 3 * (int64_t)ce[k];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * x[b];
    }
    const double detJ = inv3(iJ, J);
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < nn; ++k)
    {
      const double* uk = u + 3 * (int64_t)ce[k];
      double G[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        G[a] = iJ[a] * dN[3 * k] + iJ[a + 3] * dN[3 * k + 1] + iJ[a + 6] * dN[3 * k + 2];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) F[a + 3 * b] += uk[a] * G[b];
    }
    return detJ;
  }

  B2_HD double point_detJ_rt(int nn, const double* dN, const double* X, const int32_t* ce)
  {
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nn; ++k)
    {
      const double* x = X + 3 * (int64_t)ce[k];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * x[b];
    }
    return det3(J);
  }

  // MULF flavour of point_kinematics with run-time node count: F = (I + u (iJ_hist dN)^T) F_hist, returns
  // det J of the reference mapping; Fd_out (may be null) receives F_delta
  B2_HD double point_kinematics_mulf(int nn, const double* dN, const double* X, const double* u,
      const int32_t* ce, const double* h, double* F, double* Fd_out)
  {
    const double detJ = point_detJ_rt(nn, dN, X, ce);
    double Fd[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Fd[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int k = 0; k < nn; ++k)
    {
      const double* uk = u + 3 * (int64_t)ce[k];
      double Nh[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        Nh[a] = fma(h[a + 6], dN[3 * k + 2], fma(h[a + 3], dN[3 * k + 1], h[a] * dN[3 * k]));
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Fd[a + 3 * b] = fma(uk[a], Nh[b], Fd[a + 3 * b]);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        F[a + 3 * b] = fma(Fd[a + 6], h[9 + 2 + 3 * b], fma(Fd[a + 3], h[9 + 1 + 3 * b], Fd[a] * h[9 + 3 * b]));
    if (Fd_out)
#pragma unroll
      for (int i = 0; i < 9; ++i) Fd_out[i] = Fd[i];
    return detJ;
  }

  // MulfFormulation::update_prestress at one Gauss point (4C_solid_ele_calc_mulf.hpp:124-143):
  // deformation_gradient <- F of the current state, inverse_jacobian <- F_delta^-T inverse_jacobian
  B2_HD void mulf_update_point(int nn, const double* dN, const double* X, const double* u,
      const int32_t* ce, double* h)
  {
    double F[9], Fd[9], iFd[9], niJ[9];
    point_kinematics_mulf(nn, dN, X, u, ce, h, F, Fd);
    inv3(iFd, Fd);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        niJ[a + 3 * b] = iFd[3 * a] * h[3 * b] + iFd[1 + 3 * a] * h[1 + 3 * b] + iFd[2 + 3 * a] * h[2 + 3 * b];
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] = niJ[i], h[9 + i] = F[i];
  }

  // Square root U and inverse square root of a symmetric positive definite 3x3 (Voigt in, full column-major
  // out) by the Denman-Beavers iteration Y <- (Y + Z^-1)/2, Z <- (Z + Y^-1)/2 (Y -> C^1/2, Z -> C^-1/2;
  // quadratic convergence, 12 steps reach round-off for the stretches met here).  The reference takes the
  // root through an eigen decomposition (compute_deformation_gradient_from_gl_strains,
  // 4C_solid_ele_calc_lib.hpp:1566-1595); same tensor.
  B2_HD void sqrt_sym3(const double* Cv, double* U, double* iU)
  {
    double Y[9], Z[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) Y[a + 3 * b] = Cv[vg(a, b)];
    for (int it = 0; it < 12; ++it)
    {
      double iY[9], iZ[9];
      inv3(iY, Y);
      inv3(iZ, Z);
#pragma unroll
      for (int i = 0; i < 9; ++i)
      {
        Y[i] = 0.5 * (Y[i] + iZ[i]);
        Z[i] = 0.5 * (Z[i] + iY[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) U[i] = Y[i], iU[i] = Z[i];
  }

  // EAS kinematics at one Gauss point of a hex8 (evaluate_eas_kinematics, 4C_solid_ele_calc_eas_helpers.hpp:
  // 405-448): C <- C_enh = C + 2 stress-like(Mtilde alpha); with want_F also F <- F_enh = R U_enh,
  // R = F U^-1 (compute_deformation_gradient_from_gl_strains)
  B2_HD void consistent_deformation_gradient(const double* Cd, const double* C, double* F);
  // Cd (may be null): right Cauchy-Green tensor of the incoming F where it is not the incoming C (ANS blocks)
  B2_HD void eas_enhance(const PostArgs& A, int64_t e, int g, double detJ, bool want_F, double* F, double* C,
      const double* Cd_in = nullptr)
  {
    const double* t0 = A.eas_t0 + e * A.eas_t0_stride;
    const double* alpha = A.eas_hist + e * A.eas_hist_size + A.eas_neas;
    const double r = A.xi_gp[3 * g], s = A.xi_gp[3 * g + 1], t = A.xi_gp[3 * g + 2];
    const double scale = t0[36] / detJ;
    const int rows[21] = {0, 1, 2, 3, 3, 4, 4, 5, 5, 3, 3, 4, 4, 5, 5, 0, 0, 1, 1, 2, 2};
    const double mono[21] = {r, s, t, r, s, s, t, r, t, r * t, s * t, r * s, r * t, r * s, s * t, r * s, r * t,
        r * s, s * t, r * t, s * t};
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (A.eas_neas == 7)
    {
      // eastype_sh8_7 (TECH shell_eas): r | s | t, rt, st | r, s in the rows rr | ss | tt | rs
      const int rows7[7] = {0, 1, 2, 3, 3, 2, 2};
      const double mono7[7] = {r, s, t, r, s, r * t, s * t};
      for (int c = 0; c < 7; ++c) v[rows7[c]] += mono7[c] * alpha[c];
    }
    else
      for (int c = 0; c < A.eas_neas; ++c) v[rows[c]] += mono[c] * alpha[c];
    double Cd[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) Cd[k] = Cd_in ? Cd_in[k] : C[k];
#pragma unroll
    for (int k = 0; k < 6; ++k)
    {
      double q = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j) q += t0[k + 6 * j] * v[j];
      C[k] += (k < 3 ? 2.0 : 1.0) * scale * q;
    }
    if (!want_F) return;
    consistent_deformation_gradient(Cd, C, F);
  }

  // compute_deformation_gradient_from_gl_strains (4C_solid_ele_calc_lib.hpp:1566-1595): F <- R U(C), R = F U(Cd)^-1 with Cd the
  // right Cauchy-Green tensor of the incoming F
  B2_HD void consistent_deformation_gradient(const double* Cd, const double* C, double* F)
  {
    double U[9], iU[9], Ue[9], iUe[9], R[9];
    sqrt_sym3(Cd, U, iU);
    sqrt_sym3(C, Ue, iUe);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) R[a + 3 * b] = F[a] * iU[3 * b] + F[a + 3] * iU[1 + 3 * b] + F[a + 6] * iU[2 + 3 * b];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) F[a + 3 * b] = R[a] * Ue[3 * b] + R[a + 3] * Ue[1 + 3 * b] + R[a + 6] * Ue[2 + 3 * b];
  }

  // TECH shell_ans / shell_eas_ans on hex8: right Cauchy-Green tensor C = 2 E_ans + I of the ANS strain at Gauss point g
  // (evaluate_local_glstrain, 4C_solid_ele_calc_shell_ans_lib.hpp:169-291; T^-T as in generic_block_row_ans)
  B2_HD void ans_right_cauchy_green(const PostArgs& A, const int32_t* ce, int g, double* C)
  {
    constexpr int NN = 8;
    double X[NN][3], x[NN][3];
    for (int k = 0; k < NN; ++k)
      for (int d = 0; d < 3; ++d)
      {
        X[k][d] = A.X[3 * (int64_t)ce[k] + d];
        x[k][d] = X[k][d] + A.u[3 * (int64_t)ce[k] + d];
      }
    auto jac = [&](const double* dN, const double (*P)[3], double* J)
    {
      for (int q = 0; q < 9; ++q) J[q] = 0.0;
      for (int k = 0; k < NN; ++k)
        for (int a = 0; a < 3; ++a)
          for (int d = 0; d < 3; ++d) J[a + 3 * d] = fma(dN[3 * k + a], P[k][d], J[a + 3 * d]);
    };
    auto rows = [](const double* J, int a, int b) { return J[a] * J[b] + J[a + 3] * J[b + 3] + J[a + 6] * J[b + 6]; };
    // metric difference (a, b) between the current and the reference configuration at sampling point q
    auto sampled = [&](int q, int a, int b)
    {
      double Jr[9], Jc[9];
      jac(A.dN_ans + (size_t)q * NN * 3, X, Jr), jac(A.dN_ans + (size_t)q * NN * 3, x, Jc);
      return rows(Jc, a, b) - rows(Jr, a, b);
    };
    const double* dN = A.dN_gp + (size_t)g * NN * 3;
    const double* xi = A.xi_gp + 3 * g;
    double J[9], Jc[9], iJ[9], T[36], El[6];
    jac(dN, X, J), jac(dN, x, Jc);
    inv3(iJ, J);
    voigt_strain_transformation(iJ, T);
    El[0] = 0.5 * (rows(Jc, 0, 0) - rows(J, 0, 0));
    El[1] = 0.5 * (rows(Jc, 1, 1) - rows(J, 1, 1));
    El[3] = rows(Jc, 0, 1) - rows(J, 0, 1);
    El[2] = 0.5 * (0.25 * (1 - xi[0]) * (1 - xi[1]) * sampled(4, 2, 2) + 0.25 * (1 + xi[0]) * (1 - xi[1]) * sampled(5, 2, 2) +
                      0.25 * (1 + xi[0]) * (1 + xi[1]) * sampled(6, 2, 2) + 0.25 * (1 - xi[0]) * (1 + xi[1]) * sampled(7, 2, 2));
    El[4] = 0.5 * (1 + xi[0]) * sampled(1, 1, 2) + 0.5 * (1 - xi[0]) * sampled(3, 1, 2);
    El[5] = 0.5 * (1 - xi[1]) * sampled(0, 0, 2) + 0.5 * (1 + xi[1]) * sampled(2, 0, 2);
    for (int r = 0; r < 6; ++r)
    {
      double v = 0.0;
      for (int q = 0; q < 6; ++q) v = fma(T[r + 6 * q], El[q], v);
      C[r] = r < 3 ? 2.0 * v + 1.0 : v;
    }
  }

  // One Gauss point of calculate_stress / calculate_internal_energy.  Returns psi * integration
  // factor (0 when want_energy is false); stress6 / strain6 in the reference's row layout
  // [xx,yy,zz,xy,yz,xz] (assemble_symmetric_tensor_to_matrix_row, calc_lib_io.hpp:41-70).
  template <int MAT>
  B2_HD double post_point(const PostArgs& A, int64_t e, int g, bool want_energy, double* stress6,
      double* strain6)
  {
    const int32_t* ce = A.conn + e * A.nn;
    double F[9];
    const double detJ =
        A.gp_hist ? point_kinematics_mulf(A.nn, A.dN_gp + (size_t)g * A.nn * 3, A.X, A.u, ce,
                        A.gp_hist + (e * A.hist_pts + g) * MULF_HIST, F, nullptr)
                  : point_kinematics(A.nn, A.dN_gp + (size_t)g * A.nn * 3, A.X, A.u, ce, F);
    if (A.fbar)
    {
      // F_bar = (det F_0 / det F)^(1/3) F  (4C_solid_ele_calc_fbar.hpp:103-112, lib_fbar.hpp:47-51)
      double F0[9];
      if (A.gp_hist)  // MulfFBarFormulation: the centroid's own history (mulf_fbar.hpp:41-63)
        point_kinematics_mulf(A.nn, A.dN_cen, A.X, A.u, ce, A.gp_hist + (e * A.hist_pts + A.ngp) * MULF_HIST, F0, nullptr);
      else
        point_kinematics(A.nn, A.dN_cen, A.X, A.u, ce, F0);
      const double fb = pow(det3(F0) / det3(F), 1.0 / 3.0);
#pragma unroll
      for (int i = 0; i < 9; ++i) F[i] *= fb;
    }
    double C[6], S[6], Cm[21];
    right_cauchy_green(F, C);
    if (A.linear) linear_kinematics(F, C);
    const bool want_F =
        !A.linear && ((stress6 != nullptr && A.stress_type == STRESS_CAUCHY) || (strain6 != nullptr && A.strain_type == STRAIN_EA));
    double Cd[6];
    if (A.dN_ans != nullptr)
    {
      // ShellANSFormulation::evaluate (shell_ans.hpp:96-117): the ANS strain and the deformation gradient consistent with it
#pragma unroll
      for (int v = 0; v < 6; ++v) Cd[v] = C[v];
      ans_right_cauchy_green(A, ce, g, C);
      if (!A.eas_neas && want_F) consistent_deformation_gradient(Cd, C, F);
    }
    if (A.eas_neas)  // (KINEM linear: the enhanced deformation gradient is the identity, eas_helpers.hpp:663-665)
      eas_enhance(A, e, g, detJ, want_F, F, C, A.dN_ans != nullptr ? Cd : nullptr);
    MatParams mp = A.mp;
    if (A.per_ele_param != nullptr) mp.p[0] = A.per_ele_param[e];
    if (MAT == MAT_HYPER && A.fibers != nullptr) mp.fiber = A.fibers + 3 * e;
    if (material_hist_doubles(MAT) > 0 && A.mat_hist != nullptr)
      mp.hist = A.mat_hist + (e * A.ngp + g) * material_hist_doubles(MAT);
    if (stress6 != nullptr && A.stress_type != STRESS_NONE)
    {
      material_evaluate<MAT>(mp, C, S, Cm);
      if (A.stress_type == STRESS_2PK)
      {
#pragma unroll
        for (int v = 0; v < 6; ++v) stress6[v] = S[v];
      }
      else
      {
        // sigma = F S F^T / det F  (Solid::Utils::pk2_to_cauchy, 4C_solid_ele_utils.hpp:41-54)
        double P[9];
        first_piola(F, S, P);
        const double idet = 1.0 / det3(F);
#pragma unroll
        for (int v = 0; v < 6; ++v)
        {
          const int a = va(v), b = vb(v);
          stress6[v] = (P[a] * F[b] + P[a + 3] * F[b + 3] + P[a + 6] * F[b + 6]) * idet;
        }
      }
    }
    if (strain6 != nullptr && A.strain_type != STRAIN_NONE)
    {
      double E[6];
#pragma unroll
      for (int v = 0; v < 6; ++v) E[v] = 0.5 * (C[v] - (v < 3 ? 1.0 : 0.0));
      if (A.strain_type == STRAIN_GL)
      {
#pragma unroll
        for (int v = 0; v < 6; ++v) strain6[v] = E[v];
      }
      else
      {
        // e = F^-T E F^-1  (green_lagrange_to_euler_almansi, 4C_solid_ele_utils.hpp:64-79)
        double iF[9], T[9];
        inv3(iF, F);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b)  // T = E iF
            T[a + 3 * b] = E[vg(a, 0)] * iF[0 + 3 * b] + E[vg(a, 1)] * iF[1 + 3 * b] + E[vg(a, 2)] * iF[2 + 3 * b];
#pragma unroll
        for (int v = 0; v < 6; ++v)
        {
          const int a = va(v), b = vb(v);
          strain6[v] = iF[0 + 3 * a] * T[0 + 3 * b] + iF[1 + 3 * a] * T[1 + 3 * b] + iF[2 + 3 * a] * T[2 + 3 * b];
        }
      }
    }
    if (!want_energy) return 0.0;
    return material_strain_energy<MAT>(mp, C) * (detJ * A.w[g]);
  }

  struct MassArgs
  {
    const double* X;
    const int32_t* conn;      // device element order
    const int32_t* pos;       // scatter map of the stiffness graph
    const NodeRow* node_row;
    const double* acc;        // nullptr or 3*n_col_nodes acceleration on the column map
    const double* N_mass;     // [ngp_mass][nn]
    const double* dN_mass;    // [ngp_mass][nn][3]
    const double* w_mass;     // [ngp_mass]
    const double* dN_gp;      // stiffness rule (only read when the rules differ)
    const double* w;
    double density;           // Mat::So3Material::density()
    int64_t n_ele;
    int nn, ngp, ngp_mass, pos_stride;
    int lumped;               // struct_calc_nlnstifflmass: lump_matrix (row sums on the diagonal)
    double* mass_vals;        // nnz or nullptr
    double* f_inertia;        // n_rows or nullptr (needs acc)
  };

  B2_HD double point_detJ(int nn, const double* dN, const double* X, const int32_t* ce)
  {
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < nn; ++k)
    {
      const double* x = X + 3 * (int64_t)ce[k];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * x[b];
    }
    return det3(J);
  }

  // Row node i of the element mass matrix: m[j] = sum_gp rho N_i N_j detJ w (the same scalar on the
  // three dof diagonals of block (i,j)).
  template <int NN>
  B2_HD void mass_row(const MassArgs& A, int64_t e, int i, double* m)
  {
    const int32_t* ce = A.conn + e * NN;
    double rho = A.density;
    if (A.ngp_mass != A.ngp)
    {
      // rules differ: density = element mass / element volume over the stiffness rule
      // (4C_solid_ele_calc.cpp:230-234,277)
      double em = 0.0, ev = 0.0;
      for (int g = 0; g < A.ngp; ++g)
      {
        const double fac = point_detJ(NN, A.dN_gp + (size_t)g * NN * 3, A.X, ce) * A.w[g];
        em += A.density * fac;
        ev += fac;
      }
      rho = em / ev;
    }
#pragma unroll
    for (int j = 0; j < NN; ++j) m[j] = 0.0;
    for (int g = 0; g < A.ngp_mass; ++g)
    {
      const double fac = point_detJ(NN, A.dN_mass + (size_t)g * NN * 3, A.X, ce) * A.w_mass[g];
      const double* N = A.N_mass + (size_t)g * NN;
      const double ifactor = N[i] * fac * rho;
#pragma unroll
      for (int j = 0; j < NN; ++j) m[j] += N[j] * ifactor;
    }
  }

  // scatter of one mass row + inertia force of the row node (owner-computes, as the stiffness)
  template <int NN>
  B2_HD void mass_scatter(const MassArgs& A, int64_t e, int i, const double* m)
  {
    const int32_t* ce = A.conn + e * NN;
    const NodeRow nr = A.node_row[ce[i]];
    if (nr.rowstart < 0) return;  // row node not owned
    if (A.lumped)
    {
      // lump_matrix (4C_solid_ele_utils.cpp): every row keeps its row sum on the diagonal
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < NN; ++j) s += m[j];
      if (A.mass_vals)
      {
        const int64_t base = nr.rowstart + A.pos[e * A.pos_stride + i * NN + i];
#pragma unroll
        for (int d = 0; d < 3; ++d) B2_RED_ADD(A.mass_vals + base + (int64_t)d * nr.rowlen + d, s);
      }
      if (A.f_inertia && A.acc)
      {
        const double* a = A.acc + 3 * (int64_t)ce[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) B2_RED_ADD(A.f_inertia + nr.frow + d, s * a[d]);
      }
      return;
    }
    double fi[3] = {0, 0, 0};
    for (int j = 0; j < NN; ++j)
    {
      if (A.mass_vals)
      {
        const int64_t base = nr.rowstart + A.pos[e * A.pos_stride + i * NN + j];
#pragma unroll
        for (int d = 0; d < 3; ++d) B2_RED_ADD(A.mass_vals + base + (int64_t)d * nr.rowlen + d, m[j]);
      }
      if (A.acc)
      {
        const double* a = A.acc + 3 * (int64_t)ce[j];
#pragma unroll
        for (int d = 0; d < 3; ++d) fi[d] += m[j] * a[d];
      }
    }
    if (A.f_inertia && A.acc)
#pragma unroll
      for (int d = 0; d < 3; ++d) B2_RED_ADD(A.f_inertia + nr.frow + d, fi[d]);
  }

#if defined(__CUDACC__)
  __global__ void __launch_bounds__(128) k_mulf_init(int64_t n_ele, int nn, int ngp, const int32_t* conn,
      const double* X, const double* dN_gp, double* hist)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / ngp;
    const int g = (int)(t - e * ngp);
    if (e >= n_ele) return;
    const int32_t* ce = conn + e * nn;
    const double* dN = dN_gp + (size_t)g * nn * 3;
    double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, iJ[9];
    for (int k = 0; k < nn; ++k)
    {
      const double* x = X + 3 * (int64_t)ce[k];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) J[a + 3 * b] += dN[3 * k + a] * x[b];
    }
    inv3(iJ, J);
    double* h = hist + t * MULF_HIST;
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] = iJ[i], h[9 + i] = (i % 4 == 0) ? 1.0 : 0.0;
  }

  // struct_update_prestress (SolidEleCalc::update_prestress, 4C_solid_ele_calc.cpp:478-519)
  __global__ void __launch_bounds__(128) k_mulf_update(int64_t n_ele, int nn, int ngp, const int32_t* conn,
      const double* X, const double* u, const double* dN_gp, double* hist)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / ngp;
    const int g = (int)(t - e * ngp);
    if (e >= n_ele) return;
    mulf_update_point(nn, dN_gp + (size_t)g * nn * 3, X, u, conn + e * nn, hist + t * MULF_HIST);
  }

  // EAS line-search correction: alpha += dstep * alpha_inc (history block = {alpha_inc[n], alpha[n], ...})
  __global__ void k_eas_correct_alpha(int64_t n_ele, int neas, int hist_size, double dstep, double* hist)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * neas) return;
    double* h = hist + (t / neas) * hist_size;
    const int a = (int)(t % neas);
    h[neas + a] = fma(dstep, h[a], h[neas + a]);
  }

  // history in the caller's element numbering <-> device element order (restart pack / unpack)
  __global__ void k_hist_permute(int64_t n_ele, int per_ele, const int32_t* ele_id, const double* src,
      double* dst, int to_caller)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ele * per_ele) return;
    const int64_t e = t / per_ele;
    const int64_t r = t - e * per_ele;
    const int64_t ec = ele_id ? ele_id[e] : e;
    if (to_caller)
      dst[ec * per_ele + r] = src[t];
    else
      dst[t] = src[ec * per_ele + r];
  }

  template <int MAT>
  __global__ void __launch_bounds__(128) k_post_points(const PostArgs A)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / A.ngp;
    const int g = (int)(t - e * A.ngp);
    double psi = 0.0;
    if (e < A.n_ele)
    {
      const int64_t eid = A.ele_id ? A.ele_id[e] : e;
      const bool row = A.ele_is_row == nullptr || A.ele_is_row[eid] != 0;
      double s6[6], e6[6];
      psi = post_point<MAT>(A, e, g, A.energy != nullptr && row, A.stress ? s6 : nullptr,
          A.strain ? e6 : nullptr);
      if (A.stress && A.stress_type != STRESS_NONE)
#pragma unroll
        for (int v = 0; v < 6; ++v) A.stress[(eid * A.ngp + g) * 6 + v] = s6[v];
      if (A.strain && A.strain_type != STRAIN_NONE)
#pragma unroll
        for (int v = 0; v < 6; ++v) A.strain[(eid * A.ngp + g) * 6 + v] = e6[v];
    }
    if (A.energy != nullptr)
    {
      for (int o = 16; o > 0; o >>= 1) psi += __shfl_xor_sync(0xffffffffu, psi, o);
      if ((threadIdx.x & 31) == 0 && psi != 0.0) atomicAdd(A.energy, psi);
    }
  }

  template <int NN>
  __global__ void __launch_bounds__(128) k_mass_assemble(const MassArgs A)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = t / NN;
    const int i = (int)(t - e * NN);
    if (e >= A.n_ele) return;
    double m[NN];
    mass_row<NN>(A, e, i, m);
    mass_scatter<NN>(A, e, i, m);
  }
#endif

#if defined(__CUDACC__)
  // Mat::PlasticDruckerPrager::update for every Gauss point: last <- current (or, restore = 1, the reverse:
  // reset_to_last_converged discards the iterate)
  __global__ void k_material_history_commit(int64_t n_points, int stride, int half, int restore, double* hist)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_points * half) return;
    double* h = hist + (t / half) * stride;
    const int i = (int)(t % half);
    // restore 0: last <- current; 1: current <- last; 2: last <- current and current <- 0; 3: last <- current and
    // current <- {0, identity} (Mat::PlasticNlnLogNeoHooke::update, 4C_mat_plasticnlnlogneohooke.cpp:359-378)
    if (restore == 1)
      h[half + i] = h[i];
    else
    {
      h[i] = h[half + i];
      if (restore == 2) h[half + i] = 0.0;
      if (restore == 3) h[half + i] = (i >= 1 && i <= 3) ? 1.0 : 0.0;
    }
  }
  // fresh history of MAT_Struct_PlasticNlnLogNeoHooke: {0, I, 0, I} per Gauss point (setup, :337-349)
  __global__ void k_plastic_history_init(int64_t n_points, double* hist)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_points) return;
    double* h = hist + t * PLASTIC_HIST;
#pragma unroll
    for (int i = 0; i < PLASTIC_HIST; ++i) h[i] = ((i % 7) >= 1 && (i % 7) <= 3) ? 1.0 : 0.0;
  }
#endif
}  // namespace b200s

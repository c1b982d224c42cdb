// solid_math.cuh -- register-level FP64 building blocks of the solid element hot path.
//
// Everything is __host__ __device__ so that tests can run the exact product arithmetic on the CPU
// (tests/cpp/emulate_kernels.cpp); the shipped library only instantiates the device side.
//
// Conventions (same as the reference):
//   3x3 tensors column-major m[a + 3*b]; symmetric 2-tensors [00,11,22,01,12,02];
//   constitutive tensor: the 21 unique entries of the 6x6 Voigt matrix (major symmetry of a
//   hyperelastic tangent), row-wise upper triangle, index c21(I,J).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200s
{
  struct alignas(16) d2
  {
    double x, y;
  };

  // Voigt slot of (a,b) and its inverse
  B2_HD constexpr int vg(int a, int b)
  {
    return a == b ? a : ((a + b == 1) ? 3 : ((a + b == 3) ? 4 : 5));
  }
  B2_HD constexpr int va(int I) { return I < 3 ? I : (I == 3 ? 0 : (I == 4 ? 1 : 0)); }
  B2_HD constexpr int vb(int I) { return I < 3 ? I : (I == 3 ? 1 : 2); }
  // index of (I,J) in the packed upper triangle of a symmetric NxN matrix
  B2_HD constexpr int tri(int n, int p, int q)
  {
    return p <= q ? (p * n - p * (p - 1) / 2 + (q - p)) : (q * n - q * (q - 1) / 2 + (p - q));
  }
  B2_HD constexpr int c21(int I, int J) { return tri(6, I, J); }
  B2_HD constexpr int m45(int p, int q) { return tri(9, p, q); }

  // determinant / inverse of a 3x3 (adjugate form of the reference,
  // src/core/linalg/src/dense/4C_linalg_fixedsizematrix.hpp:1306-1335,1443-1457)
  B2_HD double det3(const double* m)
  {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) + m[3] * (m[2] * m[7] - m[1] * m[8]) +
           m[6] * (m[1] * m[5] - m[2] * m[4]);
  }
  B2_HD double inv3(double* o, const double* m)
  {
    o[0] = m[4] * m[8] - m[5] * m[7];
    o[1] = m[2] * m[7] - m[1] * m[8];
    o[2] = m[1] * m[5] - m[2] * m[4];
    const double det = m[0] * o[0] + m[3] * o[1] + m[6] * o[2];
    const double id = 1.0 / det;
    o[0] *= id, o[1] *= id, o[2] *= id;
    o[3] = id * (m[5] * m[6] - m[3] * m[8]);
    o[4] = id * (m[0] * m[8] - m[2] * m[6]);
    o[5] = id * (m[2] * m[3] - m[0] * m[5]);
    o[6] = id * (m[3] * m[7] - m[4] * m[6]);
    o[7] = id * (m[1] * m[6] - m[0] * m[7]);
    o[8] = id * (m[0] * m[4] - m[1] * m[3]);
    return det;
  }
  // inverse + determinant of a symmetric 3x3 in Voigt storage
  B2_HD double inv_sym3(double* o, const double* c)
  {
    const double a0 = c[1] * c[2] - c[4] * c[4];
    const double a3 = c[5] * c[4] - c[3] * c[2];
    const double a5 = c[3] * c[4] - c[5] * c[1];
    const double det = c[0] * a0 + c[3] * a3 + c[5] * a5;
    const double id = 1.0 / det;
    o[0] = a0 * id;
    o[3] = a3 * id;
    o[5] = a5 * id;
    o[1] = id * (c[0] * c[2] - c[5] * c[5]);
    o[4] = id * (c[5] * c[3] - c[0] * c[4]);
    o[2] = id * (c[0] * c[1] - c[3] * c[3]);
    return det;
  }

  enum MatKind
  {
    MAT_COUP_NEOHOOKE = 0,
    MAT_ISO_VOL = 1,
    MAT_SVK = 2,
    MAT_HYPER = 3, // generic MAT_ElastHyper summand list (summands.cuh)
    MAT_DP = 4,    // MAT_Struct_DruckerPrager, a history material (history_materials.cuh); KINEM linear only
    MAT_VISCO = 5, // MAT_VISCONEOHOOKE, a finite-strain history material (history_materials.cuh)
    MAT_PLASTIC = 6 // MAT_Struct_PlasticNlnLogNeoHooke, finite-strain plasticity with history (history_materials.cuh)
  };

  struct HyperMat;
  B2_HD void hyper_evaluate(const HyperMat& hm, const double* fiber, const double* C, double* S, double* Cm);
  B2_HD void drucker_prager_evaluate(const double* q, double* hist, const double* C, double* S, double* Cm);
  B2_HD void visco_neohooke_evaluate(const double* p, double* hist, const double* C, double* S, double* Cm);
  B2_HD void plastic_nlnlog_evaluate(const double* q, double* hist, const double* C, double* S, double* Cm);
  B2_HD double hyper_strain_energy(const HyperMat& hm, const double* fiber, const double* C);

  // Material parameters preprocessed once on the host (b200s_set_material):
  //   COUP_NEOHOOKE: p[0]=c=YOUNG/(4(1+NUE)), p[1]=beta=NUE/(1-2NUE)
  //                  (src/mat/elast/4C_mat_elast_coupneohooke.cpp:17-25)
  //   ISO_VOL      : p[0]=MUE (or per element), p[1]=KAPPA
  //   SVK          : p[0]=lambda, p[1]=2*mue, p[2]=mfac*(1-nu), p[3]=mfac*nu, p[4]=mfac*(1-2nu)/2
  //                  (src/mat/4C_mat_stvenantkirchhoff.hpp:136-170)
  //   HYPER        : p[] unused; hyper points to the summand list (device memory on the device)
  //   VISCO        : p[0]=alpha0=E_s/E_f, p[1]=mue, p[2]=kappa (fast modulus), p[3]=artscalar1, p[4]=artscalar2,
  //                  p[5]=scalarvisco -- functions of the time step size (history_materials.cuh make_visco_neohooke)
  struct MatParams
  {
    double p[6];
    const HyperMat* hyper;
    const double* fiber;  // HYPER with a fibre summand: the element's unit fibre vector (3 doubles)
    const double* ext;    // history materials: preprocessed parameter block in device memory
    double* hist;         // history materials: state of the Gauss point being evaluated (set by the kernel)
  };

  // Host-side preprocessing of the input-file parameters (what the reference's PAR:: constructors
  // do once per material).  Returns false for an unknown kind.
  inline bool make_mat_params(int kind, const double* params, MatParams& mp)
  {
    for (double& v : mp.p) v = 0.0;
    mp.hyper = nullptr, mp.fiber = nullptr, mp.ext = nullptr, mp.hist = nullptr;
    if (kind == MAT_COUP_NEOHOOKE)
    {
      const double young = params[0], nue = params[1];
      mp.p[0] = young / (4.0 * (1.0 + nue));  // c     (4C_mat_elast_coupneohooke.cpp:23)
      mp.p[1] = nue / (1.0 - 2.0 * nue);      // beta  (:24)
    }
    else if (kind == MAT_ISO_VOL)
    {
      mp.p[0] = params[0], mp.p[1] = params[1];
    }
    else if (kind == MAT_SVK)
    {
      const double E = params[0], nu = params[1];
      const double mfac = E / ((1.0 + nu) * (1.0 - 2.0 * nu));
      mp.p[0] = E * nu / ((1 + nu) * (1 - 2 * nu));  // lambda
      mp.p[1] = 2 * (E / (2 * (1 + nu)));            // 2 mue
      mp.p[2] = mfac * (1.0 - nu), mp.p[3] = mfac * nu, mp.p[4] = mfac * 0.5 * (1.0 - 2.0 * nu);
    }
    else
      return false;
    return true;
  }

  // Mat::So3Material::evaluate on the device: right Cauchy-Green C (Voigt, = 2E + I) in,
  // PK2 stress S[6] and tangent Cm[21] out.
  //   ElastHyper isotropic path: src/mat/4C_mat_elasthyper_service.cpp:23-75 (evaluate),
  //   :84-154 (invariant derivatives, convert_mod_to_princ), :156-206 (stress/cmat from
  //   gamma/delta), :417-435 (gamma/delta); only the non-zero delta terms of each summand set are
  //   emitted (COUP_NEOHOOKE: delta5,delta6; ISO_VOL: delta2,delta5,delta6).
  template <int MAT>
  B2_HD void material_evaluate(const MatParams& mp, const double* C, double* S, double* Cm)
  {
    if (MAT == MAT_HYPER)
    {
      hyper_evaluate(*mp.hyper, mp.fiber, C, S, Cm);
    }
    else if (MAT == MAT_DP)
    {
      drucker_prager_evaluate(mp.ext, mp.hist, C, S, Cm);
    }
    else if (MAT == MAT_VISCO)
    {
      visco_neohooke_evaluate(mp.p, mp.hist, C, S, Cm);
    }
    else if (MAT == MAT_PLASTIC)
    {
      plastic_nlnlog_evaluate(mp.ext, mp.hist, C, S, Cm);
    }
    else if (MAT == MAT_SVK)
    {
      // S = lambda tr(E) I + 2 mue E with E = (C - I)/2  (src/mat/4C_mat_stvenantkirchhoff.cpp:105-113)
      const double E0 = 0.5 * (C[0] - 1.0), E1 = 0.5 * (C[1] - 1.0), E2 = 0.5 * (C[2] - 1.0);
      const double ltr = mp.p[0] * (E0 + E1 + E2);
      S[0] = ltr + mp.p[1] * E0;
      S[1] = ltr + mp.p[1] * E1;
      S[2] = ltr + mp.p[1] * E2;
      S[3] = mp.p[1] * (0.5 * C[3]);
      S[4] = mp.p[1] * (0.5 * C[4]);
      S[5] = mp.p[1] * (0.5 * C[5]);
#pragma unroll
      for (int i = 0; i < 21; ++i) Cm[i] = 0.0;
      Cm[c21(0, 0)] = Cm[c21(1, 1)] = Cm[c21(2, 2)] = mp.p[2];
      Cm[c21(0, 1)] = Cm[c21(0, 2)] = Cm[c21(1, 2)] = mp.p[3];
      Cm[c21(3, 3)] = Cm[c21(4, 4)] = Cm[c21(5, 5)] = mp.p[4];
    }
    else
    {
      double iC[6];
      // principal invariants (4C_linalg_fixedsizematrix_voigt_notation.hpp:78-103)
      const double I1 = C[0] + C[1] + C[2];
      const double I3 = inv_sym3(iC, C);
      // dPI / ddPII of the summand set, then gamma/delta
      double g0, g2, d2 = 0.0, d5, d6;
      if (MAT == MAT_COUP_NEOHOOKE)
      {
        const double c = mp.p[0], beta = mp.p[1];
        double dPI2, ddPII2;
        if (I3 > 0)
        {
          const double pw = exp(log(I3) * (-beta - 1.));  // I3^(-beta-1), as the reference
          dPI2 = -c * pw;
          ddPII2 = c * (beta + 1.) * pw / I3;
        }
        else
          dPI2 = ddPII2 = nan("");
        g0 = 2. * c;
        g2 = 2. * I3 * dPI2;
        d5 = 4. * (I3 * dPI2 + I3 * I3 * ddPII2);
        d6 = -4. * I3 * dPI2;
      }
      else
      {
        // IsoNeoHooke + VolSussmanBathe through the modified invariants
        // (4C_mat_elast_isoneohooke.cpp:34-40, 4C_mat_elast_volsussmanbathe.cpp:36-43,
        //  4C_mat_service.cpp:125-134, convert_mod_to_princ :122-154 with the zero terms dropped)
        const double mue = mp.p[0], kappa = mp.p[1];
        const double J = sqrt(I3);         // I3^(1/2)
        const double icb = 1.0 / cbrt(I3);  // I3^(-1/3)
        const double i3 = 1.0 / I3;
        const double iJ = 1.0 / J;  // I3^(-1/2)
        const double dPm0 = 0.5 * mue, dPm2 = kappa * (J - 1.);
        const double dPI0 = icb * dPm0;
        const double dPI2 = 0.5 * iJ * dPm2 - (1. / 3.) * I1 * (icb * i3) * dPm0;
        const double ddPII2 = (4. / 9.) * (icb * i3 * i3) * I1 * dPm0 + 0.25 * i3 * kappa -
                              0.25 * (i3 * iJ) * dPm2;
        const double ddPII4 = -(1. / 3.) * (icb * i3) * dPm0;
        g0 = 2. * dPI0;
        g2 = 2. * I3 * dPI2;
        d2 = 4. * I3 * ddPII4;
        d5 = 4. * (I3 * dPI2 + I3 * I3 * ddPII2);
        d6 = -4. * I3 * dPI2;
      }
      // S = gamma0 I + gamma2 C^-1   (gamma1 = 0 for these summands)
      S[0] = g0 + g2 * iC[0];
      S[1] = g0 + g2 * iC[1];
      S[2] = g0 + g2 * iC[2];
      S[3] = g2 * iC[3];
      S[4] = g2 * iC[4];
      S[5] = g2 * iC[5];
      // Cm = d2 (I (x) C^-1 + C^-1 (x) I) + d5 C^-1 (x) C^-1 + d6 C^-1 (.) C^-1
      // Holzapfel product (4C_linalg_fixedsizematrix_tensor_products.cpp:207-214)
#pragma unroll
      for (int I = 0; I < 6; ++I)
#pragma unroll
        for (int Jv = I; Jv < 6; ++Jv)
        {
          const int a = va(I), b = vb(I), cc = va(Jv), e = vb(Jv);
          double v = d5 * iC[I] * iC[Jv] +
                     (0.5 * d6) * (iC[vg(a, cc)] * iC[vg(b, e)] + iC[vg(a, e)] * iC[vg(b, cc)]);
          if (MAT == MAT_ISO_VOL)
          {
            if (I < 3) v += d2 * iC[Jv];
            if (Jv < 3) v += d2 * iC[I];
          }
          Cm[c21(I, Jv)] = v;
        }
    }
  }

  // Mat::So3Material::strain_energy of the device materials, from C = 2E + I
  //   ElastHyper: 4C_mat_elasthyper.cpp:286-308 with the summands' add_strain_energy
  //   (4C_mat_elast_coupneohooke.cpp:31-48, ..._isoneohooke.cpp:22-31, ..._volsussmanbathe.cpp:25-34)
  //   StVenantKirchhoff: 1/2 E:C:E (4C_mat_stvenantkirchhoff.cpp:119-135)
  template <int MAT>
  B2_HD double material_strain_energy(const MatParams& mp, const double* C)
  {
    if (MAT == MAT_HYPER) return hyper_strain_energy(*mp.hyper, mp.fiber, C);
    if (MAT == MAT_DP || MAT == MAT_VISCO || MAT == MAT_PLASTIC) return 0.0;  // no strain energy in the reference; rejected on the host
    if (MAT == MAT_SVK)
    {
      const double E[6] = {0.5 * (C[0] - 1.0), 0.5 * (C[1] - 1.0), 0.5 * (C[2] - 1.0), 0.5 * C[3],
          0.5 * C[4], 0.5 * C[5]};
      const double tr = E[0] + E[1] + E[2];
      double psi = 0.0;
#pragma unroll
      for (int v = 0; v < 6; ++v)
      {
        const double Sv = (v < 3 ? mp.p[0] * tr : 0.0) + mp.p[1] * E[v];
        psi += 0.5 * Sv * E[v] * (v < 3 ? 1.0 : 2.0);
      }
      return psi;
    }
    const double I1 = C[0] + C[1] + C[2];
    const double I3 = C[0] * C[1] * C[2] + 2 * (C[3] * C[4] * C[5]) - C[1] * C[5] * C[5] -
                      C[2] * C[3] * C[3] - C[0] * C[4] * C[4];
    if (MAT == MAT_COUP_NEOHOOKE)
    {
      const double c = mp.p[0], beta = mp.p[1];
      double psi = c * (I1 - 3.);
      if (beta != 0)
        psi += (c / beta) * (exp(log(I3) * (-beta)) - 1.);
      else
        psi -= c * log(I3);
      return psi;
    }
    // IsoNeoHooke + VolSussmanBathe
    const double m1 = I1 * pow(I3, -1. / 3.), J = pow(I3, 0.5);
    return mp.p[0] * 0.5 * (m1 - 3.) + mp.p[1] * 0.5 * (J - 1.) * (J - 1.);
  }

  // KINEM linear (DisplacementBasedLinearKinematicsFormulation, 4C_solid_ele_calc_displacement_based_
  // linear_kinematics.hpp:44-62,122-150): the material sees E = sym(grad u) (evaluate_linear_gl_strain,
  // calc_lib.hpp:752-761) and the identity as deformation gradient; force = S dN w, stiffness =
  // dN C dN w without geometric term.  In: F = I + grad u.  Out: C = 2 E_lin + I for the material and
  // F = I, so that the nonlinear operators with a zero geometric stress give exactly those terms.
  B2_HD void linear_kinematics(double* F, double* C)
  {
#pragma unroll
    for (int v = 0; v < 6; ++v)
    {
      const int a = va(v), b = vb(v);
      // H = F - I; 2 E_lin = H + H^T
      C[v] = (a == b) ? (2.0 * (F[a + 3 * a] - 1.0) + 1.0) : (F[a + 3 * b] + F[b + 3 * a]);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) F[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }

  // MULF prestress (PRESTRESS_TECH mulf): Gauss-point history = MulfHistoryData without the set-up
  // flag (the device initialises it at b200s_set_element_properties): h[0..8] inverse_jacobian,
  // h[9..17] deformation_gradient, both column-major (a + 3 b).
  constexpr int MULF_HIST = 18;
  // F_delta = I + u (iJ_hist dN)^T  (evaluate_mulf_deformation_gradient_update,
  // 4C_solid_ele_calc_lib_mulf.hpp:41-60); dN[k*3+a], u[k*3+a]
  template <int NN>
  B2_HD void mulf_delta_deformation_gradient(const double* h, const double* dN, const double* u, double* Fd)
  {
#pragma unroll
    for (int i = 0; i < 9; ++i) Fd[i] = (i % 4 == 0) ? 1.0 : 0.0;
#pragma unroll 2
    for (int k = 0; k < NN; ++k)
    {
      double Nh[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        Nh[a] = fma(h[a + 6], dN[3 * k + 2], fma(h[a + 3], dN[3 * k + 1], h[a] * dN[3 * k]));
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Fd[a + 3 * b] = fma(u[3 * k + a], Nh[b], Fd[a + 3 * b]);
    }
  }
  // F = F_delta F_hist (evaluate_mulf_spatial_material_mapping, lib_mulf.hpp:62-90)
  template <int NN>
  B2_HD void mulf_deformation_gradient(const double* h, const double* dN, const double* u, double* F)
  {
    double Fd[9];
    mulf_delta_deformation_gradient<NN>(h, dN, u, Fd);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        F[a + 3 * b] = fma(Fd[a + 6], h[9 + 2 + 3 * b], fma(Fd[a + 3], h[9 + 1 + 3 * b], Fd[a] * h[9 + 3 * b]));
  }

  // Right Cauchy-Green tensor C = F^T F in Voigt storage (calc_lib.hpp:714-722)
  B2_HD void right_cauchy_green(const double* F, double* C)
  {
#pragma unroll
    for (int v = 0; v < 6; ++v)
    {
      const int a = va(v), b = vb(v);
      C[v] = F[3 * a] * F[3 * b] + F[3 * a + 1] * F[3 * b + 1] + F[3 * a + 2] * F[3 * b + 2];
    }
  }

  // P = F S  (first Piola-Kirchhoff, calc_lib.hpp:1116; callers pass S already multiplied by the
  // integration factor)
  B2_HD void first_piola(const double* F, const double* S, double* P)
  {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        P[a + 3 * b] = fma(F[a + 6], S[vg(2, b)], fma(F[a + 3], S[vg(1, b)], F[a] * S[vg(0, b)]));
  }

  // The 45 unique entries of the symmetric 9x9 matrix (S and Cm already carry the integration
  // factor w)
  //   M[(a,b)][(d,c)] = sum_ij F(b,i) Cm(a,i,j,d) F(c,j) + delta_bc * S(a,d)
  // i.e. the reference's FCFTw (einsum<"bi","aijd","cj">, calc_lib.hpp:1131-1133) with the
  // geometric term (:1138,1147-1148) folded in, so that
  //   K_ij(b,c) = sum_ad dN_i(a) M[(a,b)][(d,c)] dN_j(d).
  // p = 3a+b, q = 3d+c; symmetric because of the major symmetry of Cm.
  B2_HD void material_geometric_operator(
      const double* F, const double* S, const double* Cm, double* M)
  {
#pragma unroll
    for (int c = 0; c < 3; ++c)
    {
      // T[V][d] = sum_j Cm(V, (j,d)) F(c,j)
      double T[6][3];
#pragma unroll
      for (int V = 0; V < 6; ++V)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          T[V][d] = fma(Cm[c21(V, vg(2, d))], F[c + 6],
              fma(Cm[c21(V, vg(1, d))], F[c + 3], Cm[c21(V, vg(0, d))] * F[c]));
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
          for (int d = 0; d < 3; ++d)
          {
            const int p = 3 * a + b, q = 3 * d + c;
            if (p <= q)
            {
              const double v0 = (b == c) ? fma(F[b], T[vg(a, 0)][d], S[vg(a, d)]) : F[b] * T[vg(a, 0)][d];
              M[m45(p, q)] = fma(F[b + 6], T[vg(a, 2)][d], fma(F[b + 3], T[vg(a, 1)][d], v0));
            }
          }
    }
  }
  // Strain-like Voigt transformation of E' = M E M^T (rows / columns xx,yy,zz,xy,yz,xz, T[r + 6 c]); with M = J^-1 this is
  // T^-T of evaluate_voigt_transformation_matrix (4C_solid_ele_calc_lib.hpp:1451-1517), which the reference obtains by
  // inverting T(J) numerically
  B2_HD void voigt_strain_transformation(const double* M /*a + 3 b*/, double* T /*r + 6 c*/)
  {
    constexpr int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 2};
#pragma unroll
    for (int r = 0; r < 6; ++r)
    {
      const int a = r < 3 ? r : pa[r - 3], b = r < 3 ? r : pb[r - 3];
#pragma unroll
      for (int c = 0; c < 6; ++c)
      {
        const int cc = c < 3 ? c : pa[c - 3], dd = c < 3 ? c : pb[c - 3];
        double v;
        if (r < 3)
          v = M[a + 3 * cc] * M[a + 3 * dd];
        else
          v = c < 3 ? 2.0 * M[a + 3 * cc] * M[b + 3 * cc] : M[a + 3 * cc] * M[b + 3 * dd] + M[a + 3 * dd] * M[b + 3 * cc];
        T[r + 6 * c] = v;
      }
    }
  }

}  // namespace b200s

#include "summands.cuh"
#include "history_materials.cuh"

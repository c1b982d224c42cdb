// history_materials.cuh -- materials with Gauss-point history on the device GP-state store.
//
// MAT_Struct_DruckerPrager: small-strain elasto-plasticity with a Drucker-Prager yield surface, linear
// isotropic hardening and a non-associative flow rule; return mapping to the smooth cone or to the apex.
//   Mat::PlasticDruckerPrager::evaluate        src/mat/4C_mat_plasticdruckerprager.cpp:168-263
//   setup_cmat_elasto_plastic_cone / _apex     :127-166
//   return_to_cone / _apex_funct_and_deriv     :293-319 (both residuals are linear in the unknown)
//   Core::Utils::solve_local_newton            src/core/utils/src/numerics/4C_utils_local_newton.hpp:194-223
//   history + update()                         :109-125 (strainpl{last,curr}, strainbarpl{last,curr} per point)
//
// History layout per Gauss point (DP_HIST doubles; the pack order of Mat::PlasticDruckerPrager::pack :63-77):
//   [0..5] plastic strain, last converged    [6]  accumulated plastic strain, last converged
//   [7..12] plastic strain, current iterate  [13] accumulated plastic strain, current iterate
// tensor components in the order 00,11,22,01,12,02.
//
// The element kernels keep the constitutive tensor as the 21 entries of a major-symmetric 6x6 matrix, so the
// consistent tangent is available for ETA == ETABAR (associative in the volumetric part: the cone tangent
// eta D(x)1 + etabar 1(x)D is symmetric) and the elastic tangent (TANG elastic) always; the non-associative
// consistent tangent is rejected at b200s_set_material (the converged states are the same with either).
#pragma once

#include "solid_math.cuh"

namespace b200s
{
  constexpr int DP_HIST = 14;
  constexpr int DP_NPAR = 13;
  // preprocessed parameters: G, kappa, Hiso, abstol, cohesion, eta, xi, etabar, itermax, consistent,
  // elastic tangent c11, c12, c44 (StVenantKirchhoff::evaluate_stress_linearization)
  inline bool make_drucker_prager(const double* in /* YOUNG NUE ISOHARD TOL C ETA XI ETABAR MAXITER TANG */, double* q)
  {
    const double young = in[0], nu = in[1];
    q[0] = young / (2.0 * (1.0 + nu)), q[1] = young / (3.0 * (1.0 - 2.0 * nu));
    q[2] = in[2], q[3] = in[3], q[4] = in[4], q[5] = in[5], q[6] = in[6], q[7] = in[7], q[8] = in[8];
    q[9] = in[9] != 0.0 ? 1.0 : 0.0;
    const double mfac = young / ((1.0 + nu) * (1.0 - 2.0 * nu));
    q[10] = mfac * (1.0 - nu), q[11] = mfac * nu, q[12] = mfac * 0.5 * (1.0 - 2.0 * nu);
    // symmetric tangent storage: see the header of this file
    return !(q[9] != 0.0 && in[5] != in[7]) && in[4] != 0.0 && in[5] != 0.0 && in[7] != 0.0;
  }

  // C = 2 E + I (Voigt) in; S[6], Cm[21] out; hist = history of this point or nullptr (virgin, nothing stored)
  B2_HD void drucker_prager_evaluate(const double* q, double* hist, const double* C, double* S, double* Cm)
  {
    const double G = q[0], kappa = q[1], Hiso = q[2], abstol = q[3], coh = q[4], eta = q[5], xi = q[6], etabar = q[7];
    const int itermax = (int)q[8];
    double e[6], bar_last = 0.0;
#pragma unroll
    for (int v = 0; v < 6; ++v) e[v] = 0.5 * (C[v] - (v < 3 ? 1.0 : 0.0));   // total strain
    if (hist != nullptr)
    {
#pragma unroll
      for (int v = 0; v < 6; ++v) e[v] -= hist[v];                            // elastic trial strain
      bar_last = hist[6];
    }
    const double theta = e[0] + e[1] + e[2];
    double d[6];
#pragma unroll
    for (int v = 0; v < 6; ++v) d[v] = e[v] - (v < 3 ? theta / 3.0 : 0.0);
    const double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + 2.0 * (d[3] * d[3] + d[4] * d[4] + d[5] * d[5]);
    const double p_trial = kappa * theta;
    // sqrt(J2) of the trial stress 2 G d:  J2 = 1/2 s:s = 2 G^2 d:d
    const double sqJ2 = sqrt(0.5 * (4.0 * G * G) * dd);
    const double Phi = sqJ2 + eta * p_trial - xi * (coh + Hiso * bar_last);
    double sfac = 2.0 * G;   // deviatoric stress = sfac * d
    double p = p_trial, bar = bar_last, Dgamma = 0.0, dstrainv = 0.0;
    const bool plastic = Phi / fabs(coh) > abstol;
    if (plastic)
    {
      // both return-mapping equations are linear: the reference's local Newton ends after one update
      // (a further one only if round-off leaves the residual above TOL * C)
      const double dc = -G - kappa * etabar * eta - xi * xi * Hiso;
      double res = Phi;
      for (int it = 0; fabs(res) > abstol * coh && it <= itermax; ++it)
      {
        Dgamma -= res / dc;
        res = Phi - Dgamma * (G + eta * kappa * etabar) - (xi * xi * Dgamma * Hiso);
      }
      if ((sqJ2 - G * Dgamma) / fabs(coh) < abstol)
      {
        const double alpha = xi / eta, beta = xi / etabar;
        const double da = xi * xi / eta / etabar * Hiso + kappa;
        res = beta * coh + beta * bar_last * Hiso - p_trial;
        for (int it = 0; fabs(res) > abstol * coh && it <= itermax; ++it)
        {
          dstrainv -= res / da;
          res = beta * coh + beta * bar_last * Hiso - p_trial + dstrainv * (alpha * beta * Hiso + kappa);
        }
        bar = bar_last + xi / eta * dstrainv;
        p = p_trial - kappa * dstrainv;
        sfac = 0.0;
      }
      else
      {
        bar = bar_last + xi * Dgamma;
        sfac = 2.0 * G * (1.0 - G * Dgamma / sqJ2);
        p = p_trial - kappa * etabar * Dgamma;
      }
    }
#pragma unroll
    for (int v = 0; v < 6; ++v) S[v] = sfac * d[v] + (v < 3 ? p : 0.0);
    if (hist != nullptr)
    {
      if (plastic)
      {
        // plastic strain = total strain - elastic strain recovered from the updated stress
#pragma unroll
        for (int v = 0; v < 6; ++v)
          hist[7 + v] = 0.5 * (C[v] - (v < 3 ? 1.0 : 0.0)) - (sfac * d[v] * (1.0 / (2.0 * G)) + (v < 3 ? p / (3.0 * kappa) : 0.0));
      }
      else
      {
#pragma unroll
        for (int v = 0; v < 6; ++v) hist[7 + v] = hist[v];
      }
      hist[13] = bar;
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) Cm[i] = 0.0;
    if (Phi > 0.0 && q[9] != 0.0)
    {
      if (dstrainv != 0.0)
      {
        const double ep = kappa * (1.0 - kappa / (kappa + xi / eta * xi / etabar * Hiso));
        Cm[c21(0, 0)] = Cm[c21(1, 1)] = Cm[c21(2, 2)] = Cm[c21(0, 1)] = Cm[c21(0, 2)] = Cm[c21(1, 2)] = ep;
      }
      else
      {
        const double nd = sqrt(dd);
        const double A = 1.0 / (G + kappa * etabar * eta + xi * xi * Hiso);
        const double r = Dgamma / (sqrt(2.0) * nd);
        const double f1 = 2.0 * G * (1.0 - r), f2 = 2.0 * G * (r - G * A);
        const double f3 = -sqrt(2.0) * G * A * kappa * eta;   // eta == etabar (make_drucker_prager)
        const double f4 = kappa * (1.0 - kappa * eta * etabar * A);
#pragma unroll
        for (int P = 0; P < 6; ++P)
#pragma unroll
          for (int Q = P; Q < 6; ++Q)
          {
            const double iP = P < 3 ? 1.0 : 0.0, iQ = Q < 3 ? 1.0 : 0.0;
            const double Is = (P == Q) ? (P < 3 ? 1.0 : 0.5) : 0.0;
            Cm[c21(P, Q)] = f1 * (Is - iP * iQ / 3.0) + f2 * (d[P] / nd) * (d[Q] / nd) +
                            f3 * ((d[P] / nd) * iQ + iP * (d[Q] / nd)) + f4 * iP * iQ;
          }
      }
    }
    else
    {
      Cm[c21(0, 0)] = Cm[c21(1, 1)] = Cm[c21(2, 2)] = q[10];
      Cm[c21(0, 1)] = Cm[c21(0, 2)] = Cm[c21(1, 2)] = q[11];
      Cm[c21(3, 3)] = Cm[c21(4, 4)] = Cm[c21(5, 5)] = q[12];
    }
  }

}  // namespace b200s

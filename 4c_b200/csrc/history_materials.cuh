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
// The fast element kernels keep the constitutive tensor as the 21 entries of a major-symmetric 6x6 matrix: they
// serve ETA == ETABAR (the cone tangent eta D(x)1 + etabar 1(x)D is symmetric then) and the elastic tangent (TANG
// elastic).  The non-associative consistent tangent (ETA != ETABAR, what the reference's RetToCone ctest uses) has
// no major symmetry: such blocks run through the general kernel of generic_kernel.cuh (full 6x6 tangent, full K_e).
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
    return in[4] != 0.0 && in[5] != 0.0 && in[7] != 0.0;
  }
  // the consistent tangent is major-symmetric only for ETA == ETABAR; otherwise the block is evaluated by the general
  // (full 6x6 tangent, full K_e) kernel of generic_kernel.cuh
  inline bool drucker_prager_tangent_is_symmetric(const double* in) { return !(in[9] != 0.0 && in[5] != in[7]); }

  // C = 2 E + I (Voigt) in; S[6] and the tangent out: FULL = false: Cm[21] (upper triangle, valid for ETA == ETABAR
  // or the elastic tangent), FULL = true: Cm[36], entry [I][J] at I + 6 J like the reference's 6x6 view
  // (setup_cmat_elasto_plastic_cone :127-150: eta D(x)1 + etabar 1(x)D is not symmetric for ETA != ETABAR);
  // hist = history of this point or nullptr (virgin, nothing stored)
  template <bool FULL>
  B2_HD void drucker_prager_evaluate_t(const double* q, double* hist, const double* C, double* S, double* Cm)
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
    // tangent entry (P, Q): the upper triangle of a symmetric matrix or the full 6x6
    auto slot = [](int P, int Q) { return FULL ? P + 6 * Q : c21(P < Q ? P : Q, P < Q ? Q : P); };
#pragma unroll
    for (int i = 0; i < (FULL ? 36 : 21); ++i) Cm[i] = 0.0;
    if (Phi > 0.0 && q[9] != 0.0)
    {
      if (dstrainv != 0.0)
      {
        const double ep = kappa * (1.0 - kappa / (kappa + xi / eta * xi / etabar * Hiso));
#pragma unroll
        for (int P = 0; P < 3; ++P)
#pragma unroll
          for (int Q = FULL ? 0 : P; Q < 3; ++Q) Cm[slot(P, Q)] = ep;
      }
      else
      {
        const double nd = sqrt(dd);
        const double A = 1.0 / (G + kappa * etabar * eta + xi * xi * Hiso);
        const double r = Dgamma / (sqrt(2.0) * nd);
        const double f1 = 2.0 * G * (1.0 - r), f2 = 2.0 * G * (r - G * A);
        const double f3 = -sqrt(2.0) * G * A * kappa;
        const double f4 = kappa * (1.0 - kappa * eta * etabar * A);
#pragma unroll
        for (int P = 0; P < 6; ++P)
#pragma unroll
          for (int Q = FULL ? 0 : P; Q < 6; ++Q)
          {
            const double iP = P < 3 ? 1.0 : 0.0, iQ = Q < 3 ? 1.0 : 0.0;
            const double Is = (P == Q) ? (P < 3 ? 1.0 : 0.5) : 0.0;
            Cm[slot(P, Q)] = f1 * (Is - iP * iQ / 3.0) + f2 * (d[P] / nd) * (d[Q] / nd) +
                             f3 * (eta * (d[P] / nd) * iQ + etabar * iP * (d[Q] / nd)) + f4 * iP * iQ;
          }
      }
    }
    else
    {
#pragma unroll
      for (int P = 0; P < 3; ++P)
#pragma unroll
        for (int Q = FULL ? 0 : P; Q < 3; ++Q) Cm[slot(P, Q)] = P == Q ? q[10] : q[11];
      Cm[slot(3, 3)] = Cm[slot(4, 4)] = Cm[slot(5, 5)] = q[12];
    }
  }
  B2_HD void drucker_prager_evaluate(const double* q, double* hist, const double* C, double* S, double* Cm)
  {
    drucker_prager_evaluate_t<false>(q, hist, C, S, Cm);
  }

  // ---- MAT_VISCONEOHOOKE ----------------------------------------------------------------------------------------
  // Mat::ViscoNeoHooke::evaluate (src/mat/4C_mat_visconeohooke.cpp:222-405, GEN_MAXWELL branch): compressible
  // Neo-Hooke on the fast modulus, W = mue/2 (I1 I3^-1/3 - 3) + kappa/2 (J - 1)^2; the deviatoric stress relaxes
  // through one Maxwell branch integrated with the one-step-theta rule
  //   Q^(n+1) = artscalar2 [ artscalar1 Q^n + S_dev^(n+1) - S_dev^n ],
  //   S = S_vol + alpha0 S_dev + (1 - alpha0) Q,   C = C_vol + scalarvisco C_dev   (major-symmetric).
  // History per Gauss point (VISCO_HIST doubles, the pack order :97-104): histstresslast[6], artstresslast[6],
  // histstresscurr[6], artstresscurr[6]; update() (:197-216): last <- current, current <- 0.  The scalars depend on
  // the time step size (EvaluationContext::time_step_size, b200s_set_time_step_size).
  constexpr int VISCO_HIST = 24;
  // in: {YOUNGS_SLOW, POISSON, DENS, YOUNGS_FAST, RELAX, THETA}; false where setup() throws (:185-189)
  inline bool make_visco_neohooke(const double* in, double dt, double* p6)
  {
    const double E_s = in[0], nue = in[1], E_f = in[3], theta = in[5];
    double tau = in[4];
    if (E_f < E_s || !(tau > 0.0)) return false;
    if (E_f > E_s) tau = tau * E_s / (E_f - E_s);
    const double alpha0 = E_s / E_f, alpha1 = 1.0 - alpha0;
    const double lambda = nue * E_f / ((1.0 + nue) * (1.0 - 2.0 * nue));
    const double mue = E_f / (2.0 * (1.0 + nue));
    p6[0] = alpha0, p6[1] = mue, p6[2] = lambda + 2.0 / 3.0 * mue;
    p6[3] = (tau - dt + theta * dt) / tau;
    p6[4] = tau / (tau + theta * dt);
    p6[5] = alpha0 + alpha1 * tau / (tau + theta * dt);
    return true;
  }

  // C (tensor components 00,11,22,01,12,02) in; S[6], Cm[21] out; hist = state of this point or nullptr
  B2_HD void visco_neohooke_evaluate(const double* p, double* hist, const double* C, double* S, double* Cm)
  {
    const double alpha0 = p[0], mue = p[1], kappa = p[2], art1 = p[3], art2 = p[4], scalarvisco = p[5];
    const double alpha1 = 1.0 - alpha0;
    const double I1 = C[0] + C[1] + C[2];
    const double I3 = C[0] * C[1] * C[2] + 2.0 * C[3] * C[4] * C[5] - C[1] * C[5] * C[5] - C[2] * C[3] * C[3] -
                      C[0] * C[4] * C[4];
    const double J = sqrt(I3);
    const double I3invcubroot = pow(I3, -1.0 / 3.0);
    double Ci[6] = {C[1] * C[2] - C[4] * C[4], C[0] * C[2] - C[5] * C[5], C[0] * C[1] - C[3] * C[3],
        C[5] * C[4] - C[3] * C[2], C[3] * C[5] - C[0] * C[4], C[3] * C[4] - C[5] * C[1]};
    const double inv = 1.0 / I3;
#pragma unroll
    for (int v = 0; v < 6; ++v) Ci[v] *= inv;
#pragma unroll
    for (int v = 0; v < 6; ++v)
    {
      const double idv = v < 3 ? 1.0 : 0.0;
      const double sdev = mue * I3invcubroot * (idv - 1.0 / 3.0 * I1 * Ci[v]);
      const double S_n = hist ? hist[v] : 0.0, Q_n = hist ? hist[6 + v] : 0.0;
      const double Q = art2 * (art1 * Q_n + sdev - S_n);
      if (hist) hist[12 + v] = sdev, hist[18 + v] = Q;
      S[v] = kappa * (J - 1.0) * J * Ci[v] + alpha0 * sdev + alpha1 * Q;
    }
    const double scalar1 = 2.0 * kappa * J * J - kappa * J;
    const double scalar2 = -2.0 * kappa * J * J + 2.0 * kappa * J;
    const double scalar3 = 2.0 / 3.0 * mue * I3invcubroot * I1;
    const double scalar4 = 2.0 / 3.0 * mue * I3invcubroot;
    const double cH = scalar2 + scalarvisco * scalar3, cCC = scalar1 + scalarvisco * scalar3 / 3.0,
                 cIC = -scalarvisco * scalar4;
    constexpr int va[6] = {0, 1, 2, 0, 1, 0}, vb[6] = {0, 1, 2, 1, 2, 2};
    constexpr int vg[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};
#pragma unroll
    for (int P = 0; P < 6; ++P)
#pragma unroll
      for (int Q = P; Q < 6; ++Q)
      {
        // holzapfel_product: 1/2 (Cinv_ac Cinv_bd + Cinv_ad Cinv_bc)
        const double H = 0.5 * (Ci[vg[va[P]][va[Q]]] * Ci[vg[vb[P]][vb[Q]]] + Ci[vg[va[P]][vb[Q]]] * Ci[vg[vb[P]][va[Q]]]);
        const double iP = P < 3 ? 1.0 : 0.0, iQ = Q < 3 ? 1.0 : 0.0;
        Cm[c21(P, Q)] = cH * H + cCC * Ci[P] * Ci[Q] + cIC * (iP * Ci[Q] + iQ * Ci[P]);
      }
  }


  // ---- MAT_Struct_PlasticNlnLogNeoHooke ---------------------------------------------------------------------------
  // Mat::PlasticNlnLogNeoHooke::evaluate (src/mat/4C_mat_plasticnlnlogneohooke.cpp:385-611): finite-strain von Mises
  // plasticity in principal logarithmic strains with nonlinear isotropic hardening and a visco-plastic rate factor;
  // trial state b_e^trial = F Cp_n^-1 F^T, spectral return mapping, S = sum_a tau_a N_a (x) N_a with N_a = F^-1 n_a.
  // The response depends on F only through C (the eigenvalues of b_e^trial are those of Cp^-1 C, N_a is unchanged by a
  // rotation of F), so the device evaluates the reference's algorithm with the upper-triangular F = L^T of the Cholesky
  // factorisation C = L L^T -- what the kernels carry is C, for EAS blocks the enhanced one.
  // Parameter block (device memory, MatParams::ext, PLASTIC_NPAR doubles): {G, kappa, YIELD, ISOHARD, SATHARDENING,
  // HARDEXPO, VISC, RATE_DEPENDENCY, TOL, dt}.  History per Gauss point (PLASTIC_HIST doubles, the pack order :96-104):
  // accplstrainlast, invplrcglast[6], accplstraincurr, invplrcgcurr[6]; a fresh point holds Cp^-1 = I (setup :337-349),
  // update() copies current -> last and resets current to {0, I} (:359-378).  A local Newton that does not converge in 50
  // iterations (the reference throws) poisons the stress with NaN, which the optional finite check reports.
  constexpr int PLASTIC_HIST = 14, PLASTIC_NPAR = 10;

  inline bool make_plastic_nlnlog(const double* in10, double dt, double* q)
  {
    // in10 = {YOUNG, NUE, DENS, YIELD, ISOHARD, SATHARDENING, HARDEXPO, VISC, RATE_DEPENDENCY, TOL}
    const double ym = in10[0], nu = in10[1];
    if (!(ym > 0.0) || !(nu > -1.0 && nu < 0.5) || !(dt > 0.0) || in10[3] == 0.0) return false;
    q[0] = ym / (2.0 * (1.0 + nu)), q[1] = ym / (3.0 * (1.0 - 2.0 * nu));
    for (int i = 0; i < 7; ++i) q[2 + i] = in10[3 + i];
    q[9] = dt;
    return true;
  }

  B2_HD void plastic_nlnlog_evaluate(const double* q, double* hist, const double* C, double* S, double* Cm)
  {
    const double G = q[0], kappa = q[1], yield = q[2], isohard = q[3], infyield = q[4], hardexp = q[5], visc = q[6],
                 eps = q[7], tol = q[8], dt = q[9];
    constexpr int va[6] = {0, 1, 2, 0, 1, 0}, vb[6] = {0, 1, 2, 1, 2, 2};
    constexpr int vg[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};
    // C = L L^T; F := L^T, F(a,b) = L(b,a) (upper triangular), det F = l00 l11 l22
    const double l00 = sqrt(C[0]), l10 = C[3] / l00, l20 = C[5] / l00;
    const double l11 = sqrt(C[1] - l10 * l10), l21 = (C[4] - l20 * l10) / l11;
    const double l22 = sqrt(C[2] - l20 * l20 - l21 * l21);
    const double F[3][3] = {{l00, l10, l20}, {0.0, l11, l21}, {0.0, 0.0, l22}};
    const double logdetF = log(l00 * l11 * l22);
    const double accpl_last = hist ? hist[0] : 0.0;
    double iCp[6] = {1, 1, 1, 0, 0, 0};
    if (hist)
#pragma unroll
      for (int v = 0; v < 6; ++v) iCp[v] = hist[1 + v];
    // Be_trial = F Cp^-1 F^T (symmetric, Voigt)
    double Bt[6];
#pragma unroll
    for (int v = 0; v < 6; ++v)
    {
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int d = 0; d < 3; ++d) acc = fma(F[va[v]][c] * iCp[vg[c][d]], F[vb[v]][d], acc);
      Bt[v] = acc;
    }
    double lam2[3], n[9];
    sym_eig3_voigt(Bt, lam2, n);  // n_a = n[. + 3 a]
    // N_a = F^-1 n_a: back substitution with the upper-triangular F
    double N[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      const double z = n[2 + 3 * a] / l22;
      const double y = (n[1 + 3 * a] - l21 * z) / l11;
      const double x = (n[3 * a] - l10 * y - l20 * z) / l00;
      N[a][0] = x, N[a][1] = y, N[a][2] = z;
    }
    double dev_trial[3], dev[3], lam2_new[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
      dev_trial[a] = G * log(lam2[a]) - 2.0 / 3.0 * G * logdetF;
      lam2_new[a] = lam2[a];
    }
    const double pressure = kappa * logdetF;
    const double abs_trial = sqrt(dev_trial[0] * dev_trial[0] + dev_trial[1] * dev_trial[1] + dev_trial[2] * dev_trial[2]);
    const double y_d = yield + isohard * accpl_last + (infyield - yield) * (1.0 - exp(-hardexp * accpl_last));
    const double f_trial = sqrt(3.0 / 2.0) * abs_trial - y_d;
    double D[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) D[a][b] = kappa + 2.0 * G * ((a == b ? 1.0 : 0.0) - 1.0 / 3.0);
    double Dgamma = 0.0;
    if (f_trial <= 0.0)
    {
#pragma unroll
      for (int a = 0; a < 3; ++a) dev[a] = dev_trial[a];
    }
    else
    {
      // solve_local_newton on residuum_and_jacobian (:63-92, 4C_utils_local_newton.hpp)
      double r, jac, dy;
      int it = 0;
      for (;;)
      {
        const double acc = accpl_last + Dgamma;
        const double ex = exp(-hardexp * acc);
        const double yd = yield + isohard * acc + (infyield - yield) * (1.0 - ex);
        const double pw = pow(visc * Dgamma / dt + 1.0, eps);
        dy = (isohard + (infyield - yield) * hardexp * ex) * pw + yd * eps * pow(visc * Dgamma / dt + 1.0, eps - 1.0) * visc / dt;
        r = sqrt(3.0 / 2.0) * abs_trial - 3.0 * G * Dgamma - yd * pw;
        jac = -3.0 * G - dy;
        if (!(fabs(r) > tol)) break;
        if (it > 50)
        {
          Dgamma = nan("");
          break;
        }
        Dgamma -= r / jac;
        ++it;
      }
      double flow[3];
      const double s23 = sqrt(2.0 / 3.0) * abs_trial;
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        flow[a] = dev_trial[a] / s23;
        dev[a] = (1.0 - 2.0 * G * Dgamma / s23) * dev_trial[a];
        lam2_new[a] = lam2[a] / exp(2.0 * flow[a] * Dgamma);
      }
      const double f1 = -4.0 * G * G * Dgamma / s23;
      const double f2 = 4.0 * G * G * (sqrt(2.0 / 3.0) * Dgamma / abs_trial - 1.0 / (3.0 * G + dy));
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) D[a][b] += f1 * ((a == b ? 1.0 : 0.0) - 1.0 / 3.0) + f2 * flow[a] * flow[b];
    }
    double tau[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) tau[a] = dev[a] + pressure;
    // dyads N_a (x) N_b, symmetrised, in Voigt storage
    double Dy[3][3][6];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int v = 0; v < 6; ++v) Dy[a][b][v] = 0.5 * (N[a][va[v]] * N[b][vb[v]] + N[a][vb[v]] * N[b][va[v]]);
#pragma unroll
    for (int v = 0; v < 6; ++v) S[v] = tau[0] * Dy[0][0][v] + tau[1] * Dy[1][1][v] + tau[2] * Dy[2][2][v];
    // tangent in the principal directions (:520-571); major symmetric
    double fac[3];  // pairs (0,1), (1,2), (0,2)
    constexpr int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 2};
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      const int a = pa[k], b = pb[k];
      fac[k] = (lam2[a] != lam2[b]) ? (tau[a] * lam2[b] - tau[b] * lam2[a]) / (lam2[a] - lam2[b])
                                    : 0.5 * (D[b][b] - D[a][b]) - tau[b];
    }
#pragma unroll
    for (int P = 0; P < 6; ++P)
#pragma unroll
      for (int Q = P; Q < 6; ++Q)
      {
        double c = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          c = fma(-2.0 * tau[a] * Dy[a][a][P], Dy[a][a][Q], c);
#pragma unroll
          for (int b = 0; b < 3; ++b) c = fma(D[a][b] * Dy[a][a][P], Dy[b][b][Q], c);
        }
        // sum over ordered pairs a != b of fac (N_ab (x) N_ab + N_ab (x) N_ab^T) = 4 fac sym(N_ab) (x) sym(N_ab) per pair
#pragma unroll
        for (int k = 0; k < 3; ++k) c = fma(4.0 * fac[k] * Dy[pa[k]][pb[k]][P], Dy[pa[k]][pb[k]][Q], c);
        Cm[c21(P, Q)] = c;
      }
    if (hist)
    {
      hist[7] = accpl_last + Dgamma;
#pragma unroll
      for (int v = 0; v < 6; ++v) hist[8 + v] = lam2_new[0] * Dy[0][0][v] + lam2_new[1] * Dy[1][1][v] + lam2_new[2] * Dy[2][2][v];
    }
  }

  // history doubles per Gauss point of a material kind (0: no history)
  B2_HD constexpr int material_hist_doubles(int mat)
  {
    return mat == MAT_DP ? DP_HIST : (mat == MAT_VISCO ? VISCO_HIST : (mat == MAT_PLASTIC ? PLASTIC_HIST : 0));
  }

}  // namespace b200s

// summands.cuh -- generic isotropic MAT_ElastHyper on the device: a run-time list of strain-energy
// summands evaluated through the principal / modified invariants of C.
//
// Replaces Mat::ElastHyper::evaluate / strain_energy for summand sets that only use
// add_derivatives_principal / add_derivatives_modified / add_strain_energy
// (src/mat/4C_mat_elasthyper.cpp:286-325, src/mat/4C_mat_elasthyper_service.cpp:23-206,417-435,
//  src/mat/4C_mat_service.cpp:110-134; summands in src/mat/elast/4C_mat_elast_*.cpp).
// Stretch-based (IsoOgden, IsoVarga, CoupVarga) and anisotropic summands stay on the CPU path.
//
// __host__ __device__ like solid_math.cuh so the CPU suite can run the product arithmetic.
#pragma once

#include "solid_math.cuh"

namespace b200s
{
  // ids are part of the C ABI (b200s_summand in include/b200_solid.h)
  enum SummandKind
  {
    SUM_COUP_NEOHOOKE = 0,       // YOUNG NUE            -> p = {c, beta}
    SUM_COUP_MOONEYRIVLIN = 1,   // C1 C2 C3
    SUM_COUP_LOGNEOHOOKE = 2,    // (YN|Lame) C1 C2      -> p = {mue, lambda}
    SUM_COUP_LOGMIXNEOHOOKE = 3, // (YN|Lame) C1 C2      -> p = {mue, lambda}
    SUM_COUP_SVK = 4,            // YOUNG NUE            -> p = {lambda, mue}
    SUM_COUP_SIMOPISTER = 5,     // MUE
    SUM_COUP_BLATZKO = 6,        // MUE NUE F            -> p = {mue, beta, f}
    SUM_COUP_EXPPOL = 7,         // A B C
    SUM_COUP_1POW = 8,           // C D
    SUM_COUP_2POW = 9,           // C D
    SUM_COUP_3POW = 10,          // C D
    SUM_COUP_13APOW = 11,        // C D A
    SUM_ISO_NEOHOOKE = 12,       // MUE
    SUM_ISO_MOONEYRIVLIN = 13,   // C1 C2
    SUM_ISO_YEOH = 14,           // C1 C2 C3
    SUM_ISO_1POW = 15,           // C D
    SUM_ISO_2POW = 16,           // C D
    SUM_ISO_EXPOPOW = 17,        // K1 K2 C(=d)
    SUM_VOL_SUSSMANBATHE = 18,   // KAPPA
    SUM_VOL_PENALTY = 19,        // EPSILON GAMMA
    SUM_VOL_OGDEN = 20,          // KAPPA BETA
    SUM_VOL_POW = 21,            // A EXPON
    SUM_COUP_ANISOEXPO = 22,     // K1 K2 K1COMP K2COMP: one exponential fibre family on the element fibre FIBER1
    // formulated in (modified) principal stretches: elast_hyper_add_response_stretches
    SUM_ISO_OGDEN = 23,          // MUE ALPHA
    SUM_COUP_VARGA = 24,         // MUE BETA             -> p = {2 mue - beta, beta}
    SUM_ISO_VARGA = 25,          // MUE BETA             -> p = {2 mue - beta, beta}
    SUM_COUNT = 26
  };

  constexpr int HYPER_MAX_SUMMANDS = 8;
  struct HyperSummand
  {
    int kind;
    int pad;
    double p[4];
  };
  struct HyperMat
  {
    int n;
    int uses_modified;  // any SUM_ISO_* / SUM_VOL_* (summandProperties_.isomod)
    HyperSummand s[HYPER_MAX_SUMMANDS];
  };

  inline bool summand_is_modified(int kind) { return kind >= SUM_ISO_NEOHOOKE && kind <= SUM_VOL_POW; }
  inline bool summand_needs_fiber(int kind) { return kind == SUM_COUP_ANISOEXPO; }
  B2_HD bool summand_in_stretches(int kind) { return kind >= SUM_ISO_OGDEN && kind <= SUM_ISO_VARGA; }

  // What the reference's PAR:: constructors do once per summand: input-file parameters -> p[].
  // mode (only LogNeoHooke / LogMixNeoHooke): 0 = "YN", 1 = "Lame".  Returns false when the
  // reference would FOUR_C_THROW on the parameters.
  inline bool make_summand(int kind, const double* in, int mode, HyperSummand& s)
  {
    s.kind = kind, s.pad = 0;
    for (double& v : s.p) v = 0.0;
    auto lame = [&](double young, double nue, double& lambda, double& mue)
    {
      if (!(nue <= 0.5 && nue > -1.0)) return false;
      lambda = (nue == 0.5) ? 0.0 : young * nue / ((1.0 + nue) * (1.0 - 2.0 * nue));
      mue = young / (2.0 * (1.0 + nue));
      return true;
    };
    switch (kind)
    {
      case SUM_COUP_NEOHOOKE:
        s.p[0] = in[0] / (4.0 * (1.0 + in[1]));
        s.p[1] = in[1] / (1.0 - 2.0 * in[1]);
        return true;
      case SUM_COUP_LOGNEOHOOKE:
      case SUM_COUP_LOGMIXNEOHOOKE:
        if (mode == 1)
        {
          s.p[0] = in[0], s.p[1] = in[1];
          return true;
        }
        return lame(in[0], in[1], s.p[1], s.p[0]);
      case SUM_COUP_SVK:
        return lame(in[0], in[1], s.p[0], s.p[1]);
      case SUM_COUP_BLATZKO:
        s.p[0] = in[0], s.p[1] = in[1] / (1.0 - 2.0 * in[1]), s.p[2] = in[2];
        return true;
      case SUM_COUP_1POW:
      case SUM_COUP_2POW:
      case SUM_COUP_3POW:
        s.p[0] = in[0], s.p[1] = in[1];
        return in[1] >= 2.0;
      case SUM_COUP_13APOW:
        s.p[0] = in[0], s.p[1] = in[1], s.p[2] = in[2];
        return in[1] >= 2.0 && in[2] > 0.0;
      case SUM_ISO_1POW:
      case SUM_ISO_2POW:
        s.p[0] = in[0], s.p[1] = in[1];
        return in[1] >= 1.0;
      case SUM_ISO_EXPOPOW:
        s.p[0] = in[0], s.p[1] = in[1], s.p[2] = in[2];
        return in[2] >= 1.0;
      case SUM_VOL_PENALTY:
        s.p[0] = in[0], s.p[1] = in[1];
        return !(in[0] < 0.0 || in[1] <= 0.0);
      case SUM_COUP_MOONEYRIVLIN:
      case SUM_COUP_EXPPOL:
      case SUM_ISO_YEOH:
        s.p[0] = in[0], s.p[1] = in[1], s.p[2] = in[2];
        return true;
      case SUM_COUP_ANISOEXPO:  // GAMMA / INIT / STR_TENS_ID: element fibres, standard structural tensor only
        s.p[0] = in[0], s.p[1] = in[1], s.p[2] = in[2], s.p[3] = in[3];
        return true;
      case SUM_COUP_SIMOPISTER:
      case SUM_ISO_NEOHOOKE:
      case SUM_VOL_SUSSMANBATHE:
        s.p[0] = in[0];
        return true;
      case SUM_ISO_MOONEYRIVLIN:
      case SUM_VOL_OGDEN:
      case SUM_VOL_POW:
      case SUM_ISO_OGDEN:
        s.p[0] = in[0], s.p[1] = in[1];
        return true;
      case SUM_COUP_VARGA:
      case SUM_ISO_VARGA:
        s.p[0] = 2.0 * in[0] - in[1], s.p[1] = in[1];
        return true;
    }
    return false;
  }
  // number of input parameters of a summand (what b200s_set_material_elasthyper reads per summand)
  inline int summand_num_params(int kind)
  {
    switch (kind)
    {
      case SUM_COUP_SIMOPISTER:
      case SUM_ISO_NEOHOOKE:
      case SUM_VOL_SUSSMANBATHE:
        return 1;
      case SUM_COUP_MOONEYRIVLIN:
      case SUM_COUP_BLATZKO:
      case SUM_COUP_EXPPOL:
      case SUM_COUP_13APOW:
      case SUM_ISO_YEOH:
      case SUM_ISO_EXPOPOW:
        return 3;
      case SUM_COUP_ANISOEXPO:
        return 4;
      default:
        return (kind >= 0 && kind < SUM_COUNT) ? 2 : -1;
    }
  }

  // x^d for the integer exponents of the *Pow summands (the reference calls pow(x, d - 1.) etc.;
  // pow with an integral double exponent is exact to < 1 ulp, so is this)
  B2_HD double ipow(double x, int d) { return pow(x, (double)d); }

  // Accumulates dPI[3], ddPII[6] (w.r.t. principal invariants) of all summands and, when psi is not
  // null, the strain energy.  I = principal invariants of C.
  //   elast_hyper_evaluate_invariant_derivatives (4C_mat_elasthyper_service.cpp:84-120)
  //   convert_mod_to_princ (:122-154), invariants_modified (4C_mat_service.cpp:125-134)
  B2_HD void hyper_invariant_derivatives(
      const HyperMat& hm, const double* I, double* dPI, double* ddPII, double* psi)
  {
#pragma unroll
    for (int i = 0; i < 3; ++i) dPI[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) ddPII[i] = 0.0;
    double dPm[3] = {0, 0, 0}, ddPm[6] = {0, 0, 0, 0, 0, 0};
    double w = 0.0;
    const double I1 = I[0], I2 = I[1], I3 = I[2];
    // modified invariants
    const double m1 = I1 * pow(I3, -1. / 3.), m2 = I2 * pow(I3, -2. / 3.), J = pow(I3, 0.5);
    for (int k = 0; k < hm.n; ++k)
    {
      const double* p = hm.s[k].p;
      switch (hm.s[k].kind)
      {
        case SUM_COUP_NEOHOOKE:  // 4C_mat_elast_coupneohooke.cpp:31-68
        {
          const double c = p[0], beta = p[1];
          if (I3 > 0)
          {
            const double pw = exp(log(I3) * (-beta - 1.));
            dPI[0] += c;
            dPI[2] -= c * pw;
            ddPII[2] += c * (beta + 1.) * pw / I3;
          }
          else
          {
            dPI[0] += c;
            dPI[2] = ddPII[2] = nan("");
          }
          if (psi)
            w += c * (I1 - 3.) + (beta != 0 ? (c / beta) * (exp(log(I3) * (-beta)) - 1.) : -c * log(I3));
          break;
        }
        case SUM_COUP_MOONEYRIVLIN:  // 4C_mat_elast_coupmooneyrivlin.cpp
        {
          const double c1 = p[0], c2 = p[1], c3 = p[2];
          dPI[0] += c1;
          dPI[1] += c2;
          dPI[2] += c3 * (1 - pow(I3, -0.5)) - (c1 + 2. * c2) * pow(I3, -1.);
          ddPII[2] += (c1 + 2 * c2) * pow(I3, -2.) + 0.5 * c3 * pow(I3, -1.5);
          if (psi)
            w += c1 * (I1 - 3.) + c2 * (I2 - 3.) - (2. * c1 + 4. * c2) * log(sqrt(I3)) +
                 c3 * pow((sqrt(I3) - 1.), 2.);
          break;
        }
        case SUM_COUP_LOGNEOHOOKE:  // 4C_mat_elast_couplogneohooke.cpp
        {
          const double mue = p[0], lambda = p[1];
          const double ldf = log(sqrt(I3));
          dPI[0] += mue * 0.5;
          dPI[2] += (lambda * ldf) / (2. * I3) - mue / (2. * I3);
          ddPII[2] += lambda / (4. * I3 * I3) + mue / (2. * I3 * I3) - (lambda * ldf) / (2. * I3 * I3);
          if (psi) w += mue * 0.5 * (I1 - 3.) - mue * ldf + lambda * 0.5 * pow(ldf, 2.);
          break;
        }
        case SUM_COUP_LOGMIXNEOHOOKE:  // 4C_mat_elast_couplogmixneohooke.cpp
        {
          const double mue = p[0], lambda = p[1];
          const double sq = sqrt(I3);
          dPI[0] += mue * 0.5;
          dPI[2] += lambda * (sq - 1.) / (2. * sq) - mue / (2. * I3);
          ddPII[2] += lambda / (4. * I3) + mue / (2. * I3 * I3) - lambda * (sq - 1.) / (4. * sq * sq * sq);
          if (psi) w += mue * 0.5 * (I1 - 3.) - mue * log(sq) + lambda * 0.5 * pow((sq - 1.), 2.);
          break;
        }
        case SUM_COUP_SVK:  // 4C_mat_elast_coupSaintVenantKirchhoff.cpp
        {
          const double lambda = p[0], mue = p[1];
          dPI[0] += (0.5 * mue + 0.25 * lambda) * I1 - 0.75 * lambda - 0.5 * mue;
          dPI[1] -= 0.5 * mue;
          ddPII[0] += 0.5 * mue + 0.25 * lambda;
          if (psi)
            w += (0.25 * mue + 0.125 * lambda) * I1 * I1 - (0.75 * lambda + 0.5 * mue) * I1 -
                 0.5 * mue * I2 + 1.125 * lambda + 0.75 * mue;
          break;
        }
        case SUM_COUP_SIMOPISTER:  // 4C_mat_elast_coupsimopister.cpp
        {
          const double mue = p[0];
          dPI[0] += 0.5 * mue;
          dPI[2] -= 0.5 * mue / I3;
          ddPII[2] += 0.5 * mue / (I3 * I3);
          if (psi) w += 0.5 * mue * (I1 - 3.) - mue * log(pow(I3, 0.5));
          break;
        }
        case SUM_COUP_BLATZKO:  // 4C_mat_elast_coupblatzko.cpp
        {
          const double mue = p[0], beta = p[1], f = p[2];
          dPI[0] += f * mue / 2.;
          dPI[1] += (1. - f) * mue / (2. * I3);
          dPI[2] += -(f * mue) / (2. * pow(I3, beta + 1.)) -
                    (mue * (pow(I3, beta - 1.) - I2 / I3 / I3) * (f - 1.)) * 0.5;
          ddPII[2] += (f * mue * (beta + 1.)) / (2. * pow(I3, beta + 2.)) -
                      (mue * (f - 1.) * ((2. * I2) / I3 / I3 / I3 + pow(I3, beta - 2.) * (beta - 1.))) * 0.5;
          ddPII[3] -= (1. - f) * 0.5 * mue / I3 / I3;
          if (psi)
          {
            const double l3 = log(I3);
            const double neg = beta == 0.0 ? -l3 : expm1(-beta * l3) / beta;
            const double pos = beta == 0.0 ? l3 : expm1(beta * l3) / beta;
            w += 0.5 * mue * (f * (I1 - 3.0 + neg) + (1.0 - f) * (I2 / I3 - 3.0 + pos));
          }
          break;
        }
        case SUM_COUP_EXPPOL:  // 4C_mat_elast_coupexppol.cpp
        {
          const double a = p[0], b = p[1], c = p[2];
          const double sq = sqrt(I3);
          const double ef = exp(b * (I1 - 3.0) - (2.0 * b + c) * log(sq) + c * (sq - 1.0));
          const double t = -0.5 * (2. * b + c) / I3 + 0.5 * c / sq;
          dPI[0] += a * b * ef;
          dPI[2] += a * ef * (c / (2. * sq) - (2. * b + c) / (2. * I3));
          ddPII[0] += a * b * b * ef;
          ddPII[2] += a * ef * ((0.5 * (2. * b + c) / (I3 * I3)) - (0.25 * c / pow(I3, 1.5)) + t * t);
          ddPII[4] += a * b * ef * (c / (2. * sq) - (2. * b + c) / (2. * I3));
          if (psi) w += a * ef - a;
          break;
        }
        case SUM_COUP_1POW:  // 4C_mat_elast_coup1pow.cpp
        case SUM_COUP_2POW:  // 4C_mat_elast_coup2pow.cpp
        {
          const int q = hm.s[k].kind == SUM_COUP_1POW ? 0 : 1;
          const double c = p[0], x = I[q] - 3.;
          const int d = (int)p[1];
          dPI[q] += c * d * ipow(x, d - 1);
          ddPII[q] += (d == 2) ? (c * d * d - c * d) : (c * d * d - c * d) * ipow(x, d - 2);
          if (psi) w += c * ipow(x, d);
          break;
        }
        case SUM_COUP_3POW:  // 4C_mat_elast_coup3pow.cpp
        {
          const double c = p[0];
          const int d = (int)p[1];
          const double x = pow(I3, 1. / 3.) - 1.;
          dPI[2] += 1. / 3. * c * d * pow(I3, -2. / 3.) * ipow(x, d - 1);
          ddPII[2] += -2. / 9. * c * d * pow(I3, -5. / 3.) * ipow(x, d - 1) +
                      1. / 9. * c * d * (d - 1.) * pow(I3, -4. / 3.) * (d == 2 ? 1.0 : ipow(x, d - 2));
          if (psi) w += c * ipow(x, d);
          break;
        }
        case SUM_COUP_13APOW:  // 4C_mat_elast_coup13apow.cpp
        {
          const double c = p[0], a = p[2];
          const int d = (int)p[1];
          const double x = I1 * pow(I3, -a) - 3.;
          const double xd1 = ipow(x, d - 1), xd2 = (d == 2) ? 1.0 : ipow(x, d - 2);
          dPI[0] += c * d * pow(I3, -a) * xd1;
          dPI[2] += -a * c * d * I1 * pow(I3, -a - 1.) * xd1;
          ddPII[0] += c * d * (d - 1.) * pow(I3, -2. * a) * xd2;
          ddPII[2] += a * (a + 1.) * c * d * I1 * pow(I3, -a - 2.) * xd1 +
                      a * a * c * d * (d - 1.) * I1 * I1 * pow(I3, (-2. * a - 2.)) * xd2;
          ddPII[4] += -a * c * d * pow(I3, -a - 1.) * xd1 -
                      a * c * d * (d - 1.) * I1 * pow(I3, (-2. * a - 1.)) * xd2;
          if (psi) w += c * ipow(x, d);
          break;
        }
        case SUM_ISO_NEOHOOKE:  // 4C_mat_elast_isoneohooke.cpp:22-40
          dPm[0] += 0.5 * p[0];
          if (psi) w += p[0] * 0.5 * (m1 - 3.);
          break;
        case SUM_ISO_MOONEYRIVLIN:  // 4C_mat_elast_isomooneyrivlin.cpp
          dPm[0] += p[0];
          dPm[1] += p[1];
          if (psi) w += p[0] * (m1 - 3.) + p[1] * (m2 - 3.);
          break;
        case SUM_ISO_YEOH:  // 4C_mat_elast_isoyeoh.cpp
        {
          const double x = m1 - 3.;
          dPm[0] += p[0] + 2. * p[1] * x + 3. * p[2] * x * x;
          ddPm[0] += 2. * p[1] + 6. * p[2] * x;
          if (psi) w += p[0] * x + p[1] * x * x + p[2] * x * x * x;
          break;
        }
        case SUM_ISO_1POW:  // 4C_mat_elast_iso1pow.cpp
        case SUM_ISO_2POW:  // 4C_mat_elast_iso2pow.cpp
        {
          const int q = hm.s[k].kind == SUM_ISO_1POW ? 0 : 1;
          const double c = p[0], x = (q == 0 ? m1 : m2) - 3.;
          const int d = (int)p[1];
          dPm[q] += (d == 1) ? c * d : c * d * ipow(x, d - 1);
          ddPm[q] += (d == 1) ? 0. : ((d == 2) ? c * d * (d - 1.) : c * d * (d - 1.) * ipow(x, d - 2));
          if (psi) w += c * ipow(x, d);
          break;
        }
        case SUM_ISO_EXPOPOW:  // 4C_mat_elast_isoexpopow.cpp
        {
          const double k1 = p[0], k2 = p[1], x = m1 - 3.;
          const int d = (int)p[2];
          const double ef = exp(k2 * pow(x, (double)d));
          dPm[0] += (d == 1) ? k1 * d * 0.5 * ef : k1 * d * 0.5 * ef * ipow(x, d - 1);
          if (d == 1)
            ddPm[0] += k1 * d * d * k2 * 0.5 * ef;
          else
            ddPm[0] += k1 * d * d * k2 * 0.5 * ef * pow(x, 2. * (d - 1.)) +
                       k1 * d * 0.5 * ef * (d - 1.) * (d == 2 ? 1.0 : ipow(x, d - 2));
          if (psi && k2 != 0) w += k1 / (2. * k2) * (ef - 1.);
          break;
        }
        case SUM_VOL_SUSSMANBATHE:  // 4C_mat_elast_volsussmanbathe.cpp:25-43
          dPm[2] += p[0] * (J - 1.);
          ddPm[2] += p[0];
          if (psi) w += p[0] * 0.5 * (J - 1.) * (J - 1.);
          break;
        case SUM_VOL_PENALTY:  // 4C_mat_elast_volpenalty.cpp
        {
          const double eps = p[0], g = p[1];
          dPm[2] += eps * g * (pow(J, g - 1.) - pow(J, -g - 1.));
          ddPm[2] += eps * g * ((g - 1.) * pow(J, g - 2.) + (g + 1.) * pow(J, -g - 2.));
          if (psi) w += eps * (pow(J, g) + pow(J, -g) - 2.);
          break;
        }
        case SUM_VOL_OGDEN:  // 4C_mat_elast_vologden.cpp
        {
          const double kappa = p[0], beta = p[1];
          if (beta != 0.)
          {
            dPm[2] += kappa / (beta * beta) * (beta / J - beta * pow(J, -beta - 1.));
            ddPm[2] += kappa / (beta * beta) * (-beta / (J * J) + beta * (beta + 1.) * pow(J, -beta - 2.));
            if (psi) w += kappa / (beta * beta) * (beta * log(J) + pow(J, -beta) - 1.);
          }
          else
          {
            dPm[2] += kappa * log(J) / J;
            ddPm[2] += kappa * (1. - log(J)) / (J * J);
            if (psi) w += kappa / 2. * pow(log(J), 2.);
          }
          break;
        }
        case SUM_VOL_POW:  // 4C_mat_elast_volpow.cpp
        {
          const double a = p[0], ex = p[1];
          dPm[2] += -a * (pow(J, -ex) - 1.);
          ddPm[2] += ex * a * pow(J, -(ex + 1.));
          if (psi) w += a / (ex - 1.) * pow(J, -ex + 1.) + a * J;
          break;
        }
        default:
          break;
      }
    }
    if (hm.uses_modified)
    {
      // convert_mod_to_princ (4C_mat_elasthyper_service.cpp:122-154)
      dPI[0] += pow(I3, -1. / 3.) * dPm[0];
      dPI[1] += pow(I3, -2. / 3.) * dPm[1];
      dPI[2] += 0.5 * pow(I3, -0.5) * dPm[2] - 1. / 3. * I1 * pow(I3, -4. / 3.) * dPm[0] -
                2. / 3. * I2 * pow(I3, -5. / 3.) * dPm[1];
      ddPII[0] += pow(I3, -2. / 3.) * ddPm[0];
      ddPII[1] += pow(I3, -4. / 3.) * ddPm[1];
      ddPII[2] += (1. / 9.) * pow(I3, -8. / 3.) * I1 * I1 * ddPm[0] +
                  (4. / 9.) * I1 * I2 * pow(I3, -3.) * ddPm[5] -
                  (1. / 3.) * pow(I3, -11. / 6.) * I1 * ddPm[4] +
                  (4. / 9.) * pow(I3, -7. / 3.) * I1 * dPm[0] +
                  (4. / 9.) * pow(I3, -10. / 3.) * I2 * I2 * ddPm[1] -
                  (2. / 3.) * pow(I3, -13. / 6.) * I2 * ddPm[3] +
                  (10. / 9.) * pow(I3, -8. / 3.) * I2 * dPm[1] + 0.25 * pow(I3, -1.) * ddPm[2] -
                  0.25 * pow(I3, -1.5) * dPm[2];
      ddPII[3] += -(1. / 3.) * pow(I3, -2.) * I1 * ddPm[5] - (2. / 3.) * pow(I3, -7. / 3.) * I2 * ddPm[1] +
                  0.5 * pow(I3, -7. / 6.) * ddPm[3] - (2. / 3.) * pow(I3, -5. / 3.) * dPm[1];
      ddPII[4] += -(1. / 3.) * pow(I3, -5. / 3.) * I1 * ddPm[0] - (2. / 3.) * pow(I3, -2.) * I2 * ddPm[5] +
                  0.5 * pow(I3, -5. / 6.) * ddPm[4] - (1. / 3.) * pow(I3, -4. / 3.) * dPm[0];
      ddPII[5] += pow(I3, -1.) * ddPm[5];
    }
    if (psi) *psi = w;
  }

  // principal invariants of a symmetric tensor in Voigt storage
  // (4C_linalg_fixedsizematrix_voigt_notation.hpp:78-103)
  B2_HD void invariants_principal(const double* C, double* I)
  {
    I[0] = C[0] + C[1] + C[2];
    I[1] = 0.5 * (I[0] * I[0] - C[0] * C[0] - C[1] * C[1] - C[2] * C[2]) - C[3] * C[3] - C[4] * C[4] -
           C[5] * C[5];
    I[2] = C[0] * C[1] * C[2] + 2 * (C[3] * C[4] * C[5]) - C[1] * C[5] * C[5] - C[2] * C[3] * C[3] -
           C[0] * C[4] * C[4];
  }

  // ELAST_CoupAnisoExpo summands (src/mat/elast/4C_mat_elast_coupanisoexpobase.cpp:36-62,128-160): one
  // fibre family on the element's unit fibre a, structural tensor A = a (x) a, I4 = C : A, scalar product 1.
  // Adds to S / Cm (when not null) and to psi (when not null).
  B2_HD void hyper_add_fibers(const HyperMat& hm, const double* fiber, const double* C, double* S, double* Cm,
      double* psi)
  {
    if (fiber == nullptr) return;
    for (int k = 0; k < hm.n; ++k)
    {
      if (hm.s[k].kind != SUM_COUP_ANISOEXPO) continue;
      const double* p = hm.s[k].p;
      double Av[6];
#pragma unroll
      for (int v = 0; v < 6; ++v) Av[v] = fiber[va(v)] * fiber[vb(v)];
      double I4 = 0.0;
#pragma unroll
      for (int v = 0; v < 6; ++v) I4 += C[v] * Av[v] * (v < 3 ? 1.0 : 2.0);
      const double k1 = I4 < 1.0 ? p[2] : p[0], k2 = I4 < 1.0 ? p[3] : p[1];
      const double x = I4 - 1.0, ef = exp(k2 * x * x);
      if (psi) *psi += (k1 / (2.0 * k2)) * (ef - 1.0);
      if (S)
      {
        const double gamma = 2. * (k1 * x * ef);
        const double delta = 2. * (1. + 2. * k2 * x * x) * 2. * k1 * ef;
#pragma unroll
        for (int v = 0; v < 6; ++v) S[v] += gamma * Av[v];
#pragma unroll
        for (int P = 0; P < 6; ++P)
#pragma unroll
          for (int Q = P; Q < 6; ++Q) Cm[c21(P, Q)] += delta * Av[P] * Av[Q];
      }
    }
  }

  // Eigen decomposition of a symmetric 3x3 tensor in Voigt storage by cyclic Jacobi rotations
  // (the reference calls Core::LinAlg::eig = LAPACK dsyev): C = sum_a w[a] n_a (x) n_a, n_a = V[. + 3a].
  B2_HD void sym_eig3_voigt(const double* Cv, double* w, double* V)
  {
    double A[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) A[a + 3 * b] = Cv[vg(a, b)], V[a + 3 * b] = a == b ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep)
    {
      if (fabs(A[3]) + fabs(A[6]) + fabs(A[7]) < 1e-300) break;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q)
        {
          const double apq = A[p + 3 * q];
          if (apq == 0.0) continue;
          const double theta = (A[q + 3 * q] - A[p + 3 * p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
          for (int k = 0; k < 3; ++k)
          {
            const double akp = A[k + 3 * p], akq = A[k + 3 * q];
            A[k + 3 * p] = c * akp - sn * akq, A[k + 3 * q] = sn * akp + c * akq;
          }
          for (int k = 0; k < 3; ++k)
          {
            const double apk = A[p + 3 * k], aqk = A[q + 3 * k];
            A[p + 3 * k] = c * apk - sn * aqk, A[q + 3 * k] = sn * apk + c * aqk;
          }
          for (int k = 0; k < 3; ++k)
          {
            const double vkp = V[k + 3 * p], vkq = V[k + 3 * q];
            V[k + 3 * p] = c * vkp - sn * vkq, V[k + 3 * q] = sn * vkp + c * vkq;
          }
        }
    }
    w[0] = A[0], w[1] = A[4], w[2] = A[8];
  }

  // Summands in principal stretches (ELAST_CoupVarga) and modified principal stretches (ELAST_IsoOgden,
  // ELAST_IsoVarga): Mat::elast_hyper_add_response_stretches (4C_mat_elasthyper_service.cpp:208-386) with the
  // coefficient functions of src/mat/elast/4C_mat_elast_{isoogden,coupvarga,isovarga}.cpp.  With the
  // eigen-dyads D^a = n_a (x) n_a and T^ab = n_a (x) n_b + n_b (x) n_a the elasticity tensor of Holzapfel
  // (6.180),(6.196) is  sum_ab f1_ab D^a (x) D^b + sum_{a<b} f2_ab T^ab (x) T^ab.  Adds to S[6], Cm[21].
  B2_HD void hyper_add_stretches(const HyperMat& hm, const double* C, double* S, double* Cm)
  {
    bool princ = false, mod = false;
    for (int k = 0; k < hm.n; ++k)
    {
      princ = princ || hm.s[k].kind == SUM_COUP_VARGA;
      mod = mod || hm.s[k].kind == SUM_ISO_OGDEN || hm.s[k].kind == SUM_ISO_VARGA;
    }
    if (!princ && !mod) return;
    double ev[3], N[9], lam[3];
    sym_eig3_voigt(C, ev, N);
#pragma unroll
    for (int a = 0; a < 3; ++a) lam[a] = sqrt(ev[a]);
    double gam[3] = {0, 0, 0}, del[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (princ)
      for (int k = 0; k < hm.n; ++k)
      {
        if (hm.s[k].kind != SUM_COUP_VARGA) continue;
        const double al = hm.s[k].p[0], be = hm.s[k].p[1];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
          gam[a] += al - be / (lam[a] * lam[a]);
          del[a][a] += 2.0 * be / (lam[a] * lam[a] * lam[a]);
        }
      }
    if (mod)
    {
      const double j13 = pow(lam[0] * lam[1] * lam[2], -1.0 / 3.0);
      double ml[3], mg[3] = {0, 0, 0}, md[3] = {0, 0, 0};  // modified stretches, dPsi/dml, d2Psi/dml2 (diagonal)
#pragma unroll
      for (int a = 0; a < 3; ++a) ml[a] = j13 * lam[a];
      for (int k = 0; k < hm.n; ++k)
      {
        const double* p = hm.s[k].p;
        if (hm.s[k].kind == SUM_ISO_OGDEN)
        {
          const double f = 2.0 * p[0] / p[1];
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            mg[a] += f * pow(ml[a], p[1] - 1.0);
            md[a] += f * (p[1] - 1.0) * pow(ml[a], p[1] - 2.0);
          }
        }
        else if (hm.s[k].kind == SUM_ISO_VARGA)
        {
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            mg[a] += p[0] - p[1] / (ml[a] * ml[a]);
            md[a] += 2.0 * p[1] / (ml[a] * ml[a] * ml[a]);
          }
        }
      }
      // d ml_a / d lam_b = J^(-1/3)/3 (3 delta_ab - ml_a / ml_b)
      double B[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) B[a][b] = (j13 / 3.0) * ((a == b ? 3.0 : 0.0) - ml[a] / ml[b]);
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) gam[b] += B[a][b] * mg[a];
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int g = 0; g < 3; ++g)
        {
          double v = 0.0;
#pragma unroll
          for (int a = 0; a < 3; ++a)
          {
            v += B[a][b] * md[a] * B[a][g];                  // Psi_,mm  m_,l  m_,l   (md is diagonal)
            v += -mg[a] * B[a][b] / (3.0 * lam[g]);          // Psi_,m   m_,ll
            if (g == a) v += -mg[a] * j13 / (3.0 * lam[b]);
            if (b == g) v += mg[a] * j13 * lam[a] / (3.0 * lam[b] * lam[b]);
          }
          del[b][g] += v;
        }
    }
    double prs[3], f1[3][3], f2[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) prs[a] = gam[a] / lam[a];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
      {
        // the reference reads the upper triangle of delta (Voigt pairs 01, 12, 20)
        const double d = (a == b) ? del[a][a] : ((a + b == 2) ? del[2][0] : del[a < b ? a : b][a < b ? b : a]);
        f1[a][b] = d / (lam[a] * lam[b]) - (a == b ? gam[a] / (lam[a] * lam[a] * lam[a]) : 0.0);
      }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
      {
        f2[a][b] = 0.0;
        if (a == b) continue;
        // Voigt pair (al, be) of the reference: (0,1), (1,2), (2,0)
        const int al = (a + b == 2) ? 2 : (a < b ? a : b), be = (a + b == 2) ? 0 : (a < b ? b : a);
        if (fabs(lam[al] - lam[be]) < 1e-6)
          f2[a][b] = (f1[be][be] - f1[al][be]) / 2.0;
        else
          f2[a][b] = (prs[be] - prs[al]) / (lam[be] * lam[be] - lam[al] * lam[al]);
      }
    double D[3][6], T[3][6];  // T: pairs (0,1), (1,2), (2,0)
#pragma unroll
    for (int v = 0; v < 6; ++v)
    {
      const int i = va(v), j = vb(v);
#pragma unroll
      for (int a = 0; a < 3; ++a) D[a][v] = N[i + 3 * a] * N[j + 3 * a];
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        const int b = (a + 1) % 3;
        T[a][v] = N[i + 3 * a] * N[j + 3 * b] + N[i + 3 * b] * N[j + 3 * a];
      }
    }
#pragma unroll
    for (int v = 0; v < 6; ++v) S[v] += prs[0] * D[0][v] + prs[1] * D[1][v] + prs[2] * D[2][v];
#pragma unroll
    for (int P = 0; P < 6; ++P)
#pragma unroll
      for (int Q = P; Q < 6; ++Q)
      {
        double c = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
#pragma unroll
          for (int b = 0; b < 3; ++b) c += f1[a][b] * D[a][P] * D[b][Q];
          c += f2[a][(a + 1) % 3] * T[a][P] * T[a][Q];
        }
        Cm[c21(P, Q)] += c;
      }
  }

  // Mat::ElastHyper::evaluate for the generic summand list: S[6], Cm[21] from C (Voigt)
  // (elast_hyper_add_isotropic_stress_cmat, 4C_mat_elasthyper_service.cpp:156-206;
  //  calculate_gamma_delta :417-435)
  B2_HD void hyper_evaluate(const HyperMat& hm, const double* fiber, const double* C, double* S, double* Cm)
  {
    double I[3], iC[6], dPI[3], ddPII[6];
    invariants_principal(C, I);
    inv_sym3(iC, C);
    hyper_invariant_derivatives(hm, I, dPI, ddPII, nullptr);
    const double g0 = 2. * (dPI[0] + I[0] * dPI[1]), g1 = -2. * dPI[1], g2 = 2. * I[2] * dPI[2];
    const double d0 = 4. * (ddPII[0] + 2. * I[0] * ddPII[5] + dPI[1] + I[0] * I[0] * ddPII[1]);
    const double d1 = -4. * (ddPII[5] + I[0] * ddPII[1]);
    const double d2 = 4. * (I[2] * ddPII[4] + I[0] * I[2] * ddPII[3]);
    const double d3 = 4. * ddPII[1];
    const double d4 = -4. * I[2] * ddPII[3];
    const double d5 = 4. * (I[2] * dPI[2] + I[2] * I[2] * ddPII[2]);
    const double d6 = -4. * I[2] * dPI[2];
    const double d7 = -4. * dPI[1];
#pragma unroll
    for (int v = 0; v < 6; ++v) S[v] = (v < 3 ? g0 : 0.0) + g1 * C[v] + g2 * iC[v];
#pragma unroll
    for (int P = 0; P < 6; ++P)
#pragma unroll
      for (int Q = P; Q < 6; ++Q)
      {
        const int a = va(P), b = vb(P), c = va(Q), e = vb(Q);
        const double iP = P < 3 ? 1.0 : 0.0, iQ = Q < 3 ? 1.0 : 0.0;
        double v = d0 * iP * iQ + d1 * (iP * C[Q] + C[P] * iQ) + d2 * (iP * iC[Q] + iC[P] * iQ) +
                   d3 * C[P] * C[Q] + d4 * (C[P] * iC[Q] + iC[P] * C[Q]) + d5 * iC[P] * iC[Q] +
                   (0.5 * d6) * (iC[vg(a, c)] * iC[vg(b, e)] + iC[vg(a, e)] * iC[vg(b, c)]);
        if (P == Q) v += P < 3 ? d7 : 0.5 * d7;  // symmetric identity 1/2 (d_ac d_be + d_ae d_bc)
        Cm[c21(P, Q)] = v;
      }
    hyper_add_stretches(hm, C, S, Cm);
    hyper_add_fibers(hm, fiber, C, S, Cm, nullptr);
  }

  // Mat::ElastHyper::strain_energy (4C_mat_elasthyper.cpp:286-308)
  B2_HD double hyper_strain_energy(const HyperMat& hm, const double* fiber, const double* C)
  {
    double I[3], dPI[3], ddPII[6], psi = 0.0;
    invariants_principal(C, I);
    hyper_invariant_derivatives(hm, I, dPI, ddPII, &psi);
    hyper_add_fibers(hm, fiber, C, nullptr, nullptr, &psi);
    return psi;
  }
}  // namespace b200s

/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See solid_oracle.h for the contract.
 *
 * Plain C++ restatement of the 4C solid element evaluate + assemble path.  Citations are
 * relative to /root/reference (4C @ 96b9fa91).  Nothing here includes reference sources; the
 * numerical constants of the quadrature tables are data taken verbatim (incl. the truncated
 * hex27 decimals) because parity at 1e-12 depends on them.
 */
#include "solid_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>
#include <thread>
#include <vector>

namespace
{
  constexpr int MAXN = 27;

  // ---- node coordinates in parameter space
  // (src/core/fem/src/general/utils/4C_fem_general_utils_local_connectivity_matrices.hpp:291,314)
  const double hex27_ref[27][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1},
      {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {0, -1, -1}, {1, 0, -1}, {0, 1, -1}, {-1, 0, -1},
      {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1},
      {0, 0, -1}, {0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, 1}, {0, 0, 0}};
  const double tet10_ref[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.5, 0, 0},
      {0.5, 0.5, 0}, {0, 0.5, 0}, {0, 0, 0.5}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};

  // wedge6 / pyramid5: nodal parameter coordinates as the reference tabulates them
  // (4C_fem_general_utils_local_connectivity_matrices.hpp:318-325,564-598: wedge6 reads the first six rows of the wedge18
  // table, whose corner order (0,0),(1,0),(0,1) is NOT the order of the wedge6 shape functions r, s, 1-r-s -- the
  // nodal Jacobian check therefore visits the corners in a permuted order, which changes nothing for the check)
  const double wedge6_ref[6][3] = {{0, 0, -1}, {1, 0, -1}, {0, 1, -1}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}};
  const double pyramid5_ref[5][3] = {{-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}, {0, 0, 1}};
  // hex18: quad9 in the plane (corners, mid-sides, centre) at t = -1, then the same at t = +1 (:174-179)
  const double hex18_ref[18][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {0, -1, -1}, {1, 0, -1}, {0, 1, -1},
      {-1, 0, -1}, {0, 0, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1},
      {0, 0, 1}};

  int nnodes(int ct)
  {
    if (ct == ORC_WEDGE6) return 6;
    if (ct == ORC_PYRAMID5) return 5;
    if (ct == ORC_HEX18) return 18;
    return ct == ORC_HEX8 ? 8 : (ct == ORC_HEX27 ? 27 : (ct == ORC_TET4 ? 4 : (ct == ORC_HEX20 ? 20 : 10)));
  }
  // stiffness rule (get_gauss_rule_stiffness_matrix, 4C_solid_ele_calc_lib_integration.hpp:43-64): hex_8point,
  // hex_27point (hex27 AND hex20), tet_4point (tet10), tet_1point (tet4)
  // wedge6: wedge_6point, pyramid5: pyramid_8point (DisTypeToOptGaussRule,
  // 4C_fem_general_element_integration_select.hpp:68-82), for stiffness and mass
  int ngauss(int ct)
  {
    if (ct == ORC_WEDGE6) return 6;
    if (ct == ORC_PYRAMID5) return 8;
    if (ct == ORC_HEX18) return 18;  // hex_18point: quad_9point in the plane x line_2point (integration.cpp:112-129)
    return ct == ORC_HEX8 ? 8 : ((ct == ORC_HEX27 || ct == ORC_HEX20) ? 27 : (ct == ORC_TET4 ? 1 : 4));
  }

  void node_ref(int ct, int k, double* xi)
  {
    // hex8 / hex20 = first 8 / 20 of hex27, tet4 = first 4 of tet10
    const double* p = ct == ORC_HEX18 ? hex18_ref[k]
                      : (ct == ORC_TET10 || ct == ORC_TET4) ? tet10_ref[k]
                      : (ct == ORC_WEDGE6 ? wedge6_ref[k] : (ct == ORC_PYRAMID5 ? pyramid5_ref[k] : hex27_ref[k]));
    xi[0] = p[0], xi[1] = p[1], xi[2] = p[2];
  }

  // ---- shape functions and first derivatives
  // (src/core/fem/src/general/utils/4C_fem_general_utils_fem_shapefunctions.hpp
  //  hex8 :61-79,409-450; hex27 :198-238,706-830; tet10 :252-267,831-875)
  void shape(int ct, const double* xi, double* N, double* dN)
  {
    const double r = xi[0], s = xi[1], t = xi[2];
    if (ct == ORC_HEX8)
    {
      for (int k = 0; k < 8; ++k)
      {
        const double rk = hex27_ref[k][0], sk = hex27_ref[k][1], tk = hex27_ref[k][2];
        const double fr = 1.0 + rk * r, fs = 1.0 + sk * s, ft = 1.0 + tk * t;
        N[k] = 0.125 * fr * fs * ft;
        dN[0 + 3 * k] = rk * 0.125 * fs * ft;
        dN[1 + 3 * k] = sk * 0.125 * ft * fr;
        dN[2 + 3 * k] = tk * 0.125 * fr * fs;
      }
    }
    else if (ct == ORC_HEX27)
    {
      // 1-D quadratic Lagrange functions indexed by node coordinate -1,0,+1 -> slot 0,1,2
      auto lag = [](double x, double* L, double* dL)
      {
        L[0] = 0.5 * x * (x - 1.0), L[1] = 1.0 - x * x, L[2] = 0.5 * x * (x + 1.0);
        dL[0] = x - 0.5, dL[1] = -2.0 * x, dL[2] = x + 0.5;
      };
      double Lr[3], Ls[3], Lt[3], dLr[3], dLs[3], dLt[3];
      lag(r, Lr, dLr), lag(s, Ls, dLs), lag(t, Lt, dLt);
      for (int k = 0; k < 27; ++k)
      {
        const int a = (int)hex27_ref[k][0] + 1, b = (int)hex27_ref[k][1] + 1,
                  c = (int)hex27_ref[k][2] + 1;
        N[k] = Lr[a] * Ls[b] * Lt[c];
        dN[0 + 3 * k] = Ls[b] * Lt[c] * dLr[a];
        dN[1 + 3 * k] = Lr[a] * Lt[c] * dLs[b];
        dN[2 + 3 * k] = Lr[a] * Ls[b] * dLt[c];
      }
    }
    else if (ct == ORC_HEX18)
    {
      // hex18 (shapefunctions.hpp:112-145, 531-600): the quad9 functions of (r, s) times (1 -+ t) / 2
      auto lag = [](double x, double* L, double* dL)
      {
        L[0] = 0.5 * x * (x - 1.0), L[1] = 1.0 - x * x, L[2] = 0.5 * x * (x + 1.0);
        dL[0] = x - 0.5, dL[1] = -2.0 * x, dL[2] = x + 0.5;
      };
      double Lr[3], Ls[3], dLr[3], dLs[3];
      lag(r, Lr, dLr), lag(s, Ls, dLs);
      for (int k = 0; k < 18; ++k)
      {
        const int a = (int)hex18_ref[k][0] + 1, b = (int)hex18_ref[k][1] + 1;
        const double sg = hex18_ref[k][2], lin = 0.5 * (1.0 + sg * t);
        N[k] = Lr[a] * Ls[b] * lin;
        dN[0 + 3 * k] = dLr[a] * Ls[b] * lin;
        dN[1 + 3 * k] = Lr[a] * dLs[b] * lin;
        dN[2 + 3 * k] = Lr[a] * Ls[b] * 0.5 * sg;
      }
    }
    else if (ct == ORC_WEDGE6)
    {
      // linear wedge (shapefunctions.hpp:268-279, 877-903): triangle coordinates r, s, 1-r-s times (1 -+ t)/2
      const double tri[3] = {r, s, 1.0 - r - s};
      const double dtri[3][2] = {{1, 0}, {0, 1}, {-1, -1}};
      for (int k = 0; k < 6; ++k)
      {
        const int c = k % 3;
        const double sg = k < 3 ? -1.0 : 1.0, lin = 0.5 * (1.0 + sg * t);
        N[k] = tri[c] * lin;
        dN[0 + 3 * k] = dtri[c][0] * lin;
        dN[1 + 3 * k] = dtri[c][1] * lin;
        dN[2 + 3 * k] = 0.5 * sg * tri[c];
      }
    }
    else if (ct == ORC_PYRAMID5)
    {
      // pyramid with the rational term r s t / (1 - t) (shapefunctions.hpp:320-348, 976-1018); at the apex (|t - 1| <= 1e-14)
      // the reference sets the term to 0 in the functions and its three derivatives to 1
      const bool apex = !(std::fabs(t - 1.0) > 1e-14);
      const double q = apex ? 0.0 : r * s * t / (1.0 - t);
      const double qr = apex ? 1.0 : s * t / (1.0 - t), qs = apex ? 1.0 : r * t / (1.0 - t),
                   qt = apex ? 1.0 : r * s / ((1.0 - t) * (1.0 - t));
      for (int k = 0; k < 4; ++k)
      {
        const double rk = pyramid5_ref[k][0], sk = pyramid5_ref[k][1], sg = rk * sk;  // +1, -1, +1, -1
        N[k] = 0.25 * ((1.0 + rk * r) * (1.0 + sk * s) - t + sg * q);
        dN[0 + 3 * k] = 0.25 * (rk * (1.0 + sk * s) + sg * qr);
        dN[1 + 3 * k] = 0.25 * (sk * (1.0 + rk * r) + sg * qs);
        dN[2 + 3 * k] = 0.25 * (sg * qt - 1.0);
      }
      N[4] = t;
      dN[0 + 12] = 0.0, dN[1 + 12] = 0.0, dN[2 + 12] = 1.0;
    }
    else if (ct == ORC_TET4)
    {
      // linear tetrahedron (shapefunctions.hpp:239-251, 813-830)
      N[0] = 1.0 - r - s - t, N[1] = r, N[2] = s, N[3] = t;
      const double D[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      for (int k = 0; k < 4; ++k)
        for (int a = 0; a < 3; ++a) dN[a + 3 * k] = D[k][a];
    }
    else if (ct == ORC_HEX20)
    {
      // 20-node serendipity hexahedron (shapefunctions.hpp:146-197, 607-705): corner k
      //   N = 1/8 (1 + r_k r)(1 + s_k s)(1 + t_k t)((1 + r_k r) + (1 + s_k s) + (1 + t_k t) - 5),
      // mid-edge node on an edge along direction a: 1/4 (1 - xi_a^2) (1 + x_b x_b^k)(1 + x_c x_c^k)
      for (int k = 0; k < 20; ++k)
      {
        const double c[3] = {hex27_ref[k][0], hex27_ref[k][1], hex27_ref[k][2]};
        const double x[3] = {r, s, t};
        if (k < 8)
        {
          const double f[3] = {1.0 + c[0] * r, 1.0 + c[1] * s, 1.0 + c[2] * t};
          const double sum = f[0] + f[1] + f[2] - 5.0;
          N[k] = 0.125 * f[0] * f[1] * f[2] * sum;
          for (int a = 0; a < 3; ++a)
          {
            const int b = (a + 1) % 3, e = (a + 2) % 3;
            // d/dx_a [f_a f_b f_e (f_a + f_b + f_e - 5)] = c_a f_b f_e (2 f_a + f_b + f_e - 5)
            dN[a + 3 * k] = 0.125 * c[a] * f[b] * f[e] * (2.0 * f[a] + f[b] + f[e] - 5.0);
          }
        }
        else
        {
          int a = 0;  // the direction with node coordinate 0
          for (int d = 0; d < 3; ++d)
            if (c[d] == 0.0) a = d;
          const int b = (a + 1) % 3, e = (a + 2) % 3;
          const double q = 1.0 - x[a] * x[a], fb = 1.0 + c[b] * x[b], fe = 1.0 + c[e] * x[e];
          N[k] = 0.25 * q * fb * fe;
          dN[a + 3 * k] = -0.5 * x[a] * fb * fe;
          dN[b + 3 * k] = 0.25 * q * c[b] * fe;
          dN[e + 3 * k] = 0.25 * q * fb * c[e];
        }
      }
    }
    else
    {
      const double u = 1.0 - r - s - t;
      N[0] = u * (2.0 * u - 1.0), N[1] = r * (2.0 * r - 1.0), N[2] = s * (2.0 * s - 1.0),
      N[3] = t * (2.0 * t - 1.0);
      N[4] = 4.0 * r * u, N[5] = 4.0 * r * s, N[6] = 4.0 * s * u, N[7] = 4.0 * t * u,
      N[8] = 4.0 * r * t, N[9] = 4.0 * s * t;
      const double d0 = -4.0 * u + 1.0;
      const double D[10][3] = {{d0, d0, d0}, {4.0 * r - 1.0, 0, 0}, {0, 4.0 * s - 1.0, 0},
          {0, 0, 4.0 * t - 1.0}, {4.0 * (u - r), -4.0 * r, -4.0 * r}, {4.0 * s, 4.0 * r, 0},
          {-4.0 * s, 4.0 * (u - s), -4.0 * s}, {-4.0 * t, -4.0 * t, 4.0 * (u - t)},
          {4.0 * t, 0, 4.0 * r}, {0, 4.0 * t, 4.0 * s}};
      for (int k = 0; k < 10; ++k)
        for (int a = 0; a < 3; ++a) dN[a + 3 * k] = D[k][a];
    }
  }

  // ---- Gauss rules for force/stiffness
  // (rule selection src/solid_ele/4C_solid_ele_calc_lib_integration.hpp:43-64; tables
  //  src/core/fem/src/general/utils/4C_fem_general_utils_integration.cpp hex_8point :74-104,
  //  hex_27point :130-245 (truncated decimals!), tet_4point :331-352)
  void gauss(int ct, double* xi, double* w)
  {
    if (ct == ORC_HEX8)
    {
      const double g = 1.0 / std::sqrt(3.0);
      for (int q = 0; q < 8; ++q)
      {
        for (int d = 0; d < 3; ++d) xi[3 * q + d] = hex27_ref[q][d] * g;
        w[q] = 1.0;
      }
    }
    else if (ct == ORC_WEDGE6)
    {
      // wedge_6point (4C_fem_general_utils_integration.cpp:5992-6026): the 3-point triangle rule at t = +-1/sqrt(3)
      const double g = 1.0 / std::sqrt(3.0), a = 2.0 / 3.0, b = 1.0 / 6.0;
      const double T[3][2] = {{a, b}, {b, a}, {b, b}};
      for (int q = 0; q < 6; ++q)
      {
        xi[3 * q] = T[q % 3][0], xi[3 * q + 1] = T[q % 3][1], xi[3 * q + 2] = q < 3 ? g : -g;
        w[q] = b;
      }
    }
    else if (ct == ORC_PYRAMID5)
    {
      // pyramid_8point (4C_fem_general_utils_integration.cpp:6092-6132), decimals as tabulated there
      const double a1 = 0.26318405556971, a2 = 0.50661630334979, t1 = 0.54415184401122, t2 = 0.12251482265544;
      const double w1 = 0.10078588207983, w2 = 0.23254745125351;
      for (int q = 0; q < 8; ++q)
      {
        const double a = (q & 1) ? a2 : a1;
        xi[3 * q] = (q < 4 ? -a : a);
        xi[3 * q + 1] = ((q & 2) ? a : -a);
        xi[3 * q + 2] = (q & 1) ? t2 : t1;
        w[q] = (q & 1) ? w2 : w1;
      }
    }
    else if (ct == ORC_HEX18)
    {
      // hex_18point (4C_fem_general_utils_integration.cpp:112-129): quad_9point (:6185-6233, corners, mid-sides, centre,
      // decimals as tabulated) for every point of line_2point (:6946-6953)
      const double g = 0.7745966692415, w1 = 0.5555555555556, w2 = 0.8888888888889, d = 0.5773502691896;
      for (int q = 0; q < 18; ++q)
      {
        const int p = q % 9;
        xi[3 * q] = hex18_ref[p][0] * g, xi[3 * q + 1] = hex18_ref[p][1] * g, xi[3 * q + 2] = q < 9 ? -d : d;
        const double wr = hex18_ref[p][0] == 0.0 ? w2 : w1, ws = hex18_ref[p][1] == 0.0 ? w2 : w1;
        w[q] = 1.0 * (wr * ws);  // line_2point weights are 1
      }
    }
    else if (ct == ORC_TET4)
    {
      // tet_1point (4C_fem_general_utils_integration.cpp:322-330)
      xi[0] = xi[1] = xi[2] = 0.25;
      w[0] = 1.0 / 6.0;
    }
    else if (ct == ORC_HEX27 || ct == ORC_HEX20)
    {
      const double xi3 = 0.7745966692415;
      const double w1 = 0.5555555555556, w2 = 0.8888888888889;
      for (int q = 0; q < 27; ++q)
      {
        for (int d = 0; d < 3; ++d) xi[3 * q + d] = hex27_ref[q][d] * xi3;
        const double wx = hex27_ref[q][0] == 0.0 ? w2 : w1, wy = hex27_ref[q][1] == 0.0 ? w2 : w1,
                     wz = hex27_ref[q][2] == 0.0 ? w2 : w1;
        w[q] = wx * wy * wz;
      }
    }
    else
    {
      const double pa = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0, pb = (5.0 - std::sqrt(5.0)) / 20.0;
      const double P[4][3] = {{pb, pb, pb}, {pa, pb, pb}, {pb, pa, pb}, {pb, pb, pa}};
      for (int q = 0; q < 4; ++q)
      {
        for (int d = 0; d < 3; ++d) xi[3 * q + d] = P[q][d];
        w[q] = 1.0 / 6.0 / 4.0;
      }
    }
  }

  // ---- 3x3 helpers, column-major m[a + 3b]
  // (src/core/linalg/src/dense/4C_linalg_fixedsizematrix.hpp invert3x3 :1306-1335, determinant :1443-1457)
  double det3(const double* m)
  {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) + m[3] * (m[2] * m[7] - m[1] * m[8]) +
           m[6] * (m[1] * m[5] - m[2] * m[4]);
  }
  double inv3(double* o, const double* m)
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

  // Voigt slot of (a,b): [00,11,22,01,12,02]
  const int VG[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};
  const int VA[6] = {0, 1, 2, 0, 1, 0}, VB[6] = {0, 1, 2, 1, 2, 2};
  // doubles of Gauss-point history of the finite-strain history materials (0: none)
  inline int finite_strain_hist(int kind)
  {
    return kind == ORC_MAT_VISCONEOHOOKE ? ORC_VISCO_HIST : (kind == ORC_MAT_PLASTIC_NLNLOG ? ORC_PLASTIC_HIST : 0);
  }

  // ---- Jacobian mapping at one parameter point (src/solid_ele/4C_solid_ele_calc_lib.hpp:499-516)
  struct JacMap
  {
    double J[9], iJ[9], det;
    double G[MAXN][3];  // N_XYZ
  };
  void jacobian_mapping(int n, const double* dN, const double* X, JacMap& jm)
  {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int k = 0; k < n; ++k) v += dN[a + 3 * k] * X[3 * k + b];
        jm.J[a + 3 * b] = v;
      }
    inv3(jm.iJ, jm.J);
    jm.det = det3(jm.J);
    for (int k = 0; k < n; ++k)
      for (int a = 0; a < 3; ++a)
      {
        double v = 0.0;
        for (int b = 0; b < 3; ++b) v += jm.iJ[a + 3 * b] * dN[b + 3 * k];
        jm.G[k][a] = v;
      }
  }

  // F = I + u N_XYZ^T  (src/solid_ele/4C_solid_ele_calc_lib.hpp:646-660)
  void deformation_gradient(int n, const double* u, const JacMap& jm, double* F)
  {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = (a == b) ? 1.0 : 0.0;
        for (int k = 0; k < n; ++k) v += u[3 * k + a] * jm.G[k][b];
        F[a + 3 * b] = v;
      }
  }

  // ---- materials ------------------------------------------------------------------------------
  // calculate_gamma_delta (src/mat/4C_mat_elasthyper_service.cpp:417-435)
  void gamma_delta(const double* I, const double* dPI, const double* ddPII, double* g, double* d)
  {
    g[0] = 2. * (dPI[0] + I[0] * dPI[1]);
    g[1] = -2. * dPI[1];
    g[2] = 2. * I[2] * dPI[2];
    d[0] = 4. * (ddPII[0] + 2. * I[0] * ddPII[5] + dPI[1] + I[0] * I[0] * ddPII[1]);
    d[1] = -4. * (ddPII[5] + I[0] * ddPII[1]);
    d[2] = 4. * (I[2] * ddPII[4] + I[0] * I[2] * ddPII[3]);
    d[3] = 4. * ddPII[1];
    d[4] = -4. * I[2] * ddPII[3];
    d[5] = 4. * (I[2] * dPI[2] + I[2] * I[2] * ddPII[2]);
    d[6] = -4. * I[2] * dPI[2];
    d[7] = -4. * dPI[1];
  }

  // elast_hyper_evaluate_invariant_derivatives for the supported summand sets
  // (src/mat/4C_mat_elasthyper_service.cpp:84-154; summands
  //  src/mat/elast/4C_mat_elast_coupneohooke.cpp:17-25,50-68, ..._isoneohooke.cpp:34-40,
  //  ..._volsussmanbathe.cpp:36-43; invariants_modified src/mat/4C_mat_service.cpp:125-134)
  void invariant_derivatives(int kind, const double* par, const double* I, double* dPI, double* ddPII)
  {
    for (int i = 0; i < 3; ++i) dPI[i] = 0.0;
    for (int i = 0; i < 6; ++i) ddPII[i] = 0.0;
    if (kind == ORC_MAT_COUP_NEOHOOKE)
    {
      const double youngs = par[0], nue = par[1];
      const double c = youngs / (4.0 * (1.0 + nue));
      const double beta = nue / (1.0 - 2.0 * nue);
      dPI[0] += c;
      if (I[2] > 0)
      {
        const double p = std::exp(std::log(I[2]) * (-beta - 1.));
        dPI[2] -= c * p;
        ddPII[2] += c * (beta + 1.) * p / I[2];
      }
      else
        dPI[2] = ddPII[2] = std::numeric_limits<double>::quiet_NaN();
    }
    else if (kind == ORC_MAT_ISO_VOL)
    {
      const double mue = par[0], kappa = par[1];
      double modinv[3], dPm[3] = {0, 0, 0}, ddPm[6] = {0, 0, 0, 0, 0, 0};
      modinv[0] = I[0] * std::pow(I[2], -1. / 3.);
      modinv[1] = I[1] * std::pow(I[2], -2. / 3.);
      modinv[2] = std::pow(I[2], 1. / 2.);
      dPm[0] += 0.5 * mue;                  // IsoNeoHooke
      dPm[2] += kappa * (modinv[2] - 1.);   // VolSussmanBathe
      ddPm[2] += kappa;
      // convert_mod_to_princ (:122-154)
      dPI[0] += std::pow(I[2], -1. / 3.) * dPm[0];
      dPI[1] += std::pow(I[2], -2. / 3.) * dPm[1];
      dPI[2] += 0.5 * std::pow(I[2], -0.5) * dPm[2] -
                1. / 3. * I[0] * std::pow(I[2], -4. / 3.) * dPm[0] -
                2. / 3. * I[1] * std::pow(I[2], -5. / 3.) * dPm[1];
      ddPII[0] += std::pow(I[2], -2. / 3.) * ddPm[0];
      ddPII[1] += std::pow(I[2], -4. / 3.) * ddPm[1];
      ddPII[2] += (1. / 9.) * std::pow(I[2], -8. / 3.) * I[0] * I[0] * ddPm[0] +
                  (4. / 9.) * I[0] * I[1] * std::pow(I[2], -3.) * ddPm[5] -
                  (1. / 3.) * std::pow(I[2], -11. / 6.) * I[0] * ddPm[4] +
                  (4. / 9.) * std::pow(I[2], -7. / 3.) * I[0] * dPm[0] +
                  (4. / 9.) * std::pow(I[2], -10. / 3.) * I[1] * I[1] * ddPm[1] -
                  (2. / 3.) * std::pow(I[2], -13. / 6.) * I[1] * ddPm[3] +
                  (10. / 9.) * std::pow(I[2], -8. / 3.) * I[1] * dPm[1] +
                  0.25 * std::pow(I[2], -1.) * ddPm[2] - 0.25 * std::pow(I[2], -1.5) * dPm[2];
      ddPII[3] += -(1. / 3.) * std::pow(I[2], -2.) * I[0] * ddPm[5] -
                  (2. / 3.) * std::pow(I[2], -7. / 3.) * I[1] * ddPm[1] +
                  0.5 * std::pow(I[2], -7. / 6.) * ddPm[3] -
                  (2. / 3.) * std::pow(I[2], -5. / 3.) * dPm[1];
      ddPII[4] += -(1. / 3.) * std::pow(I[2], -5. / 3.) * I[0] * ddPm[0] -
                  (2. / 3.) * std::pow(I[2], -2.) * I[1] * ddPm[5] +
                  0.5 * std::pow(I[2], -5. / 6.) * ddPm[4] -
                  (1. / 3.) * std::pow(I[2], -4. / 3.) * dPm[0];
      ddPII[5] += std::pow(I[2], -1.) * ddPm[5];
    }
  }


  // ---- generic MAT_ElastHyper: list of isotropic summands ----------------------------------------
  // Packed parameter block (oracle.py pack_elasthyper): par[0] = number of summands, then per summand
  // 6 doubles {kind (ORC_SUM_*), mode, p0, p1, p2, p3} with the INPUT-FILE parameters; what the
  // PAR:: constructors derive from them is restated in place.
  //   Mat::ElastHyper::evaluate / strain_energy   src/mat/4C_mat_elasthyper.cpp:286-325
  //   elast_hyper_evaluate_invariant_derivatives  src/mat/4C_mat_elasthyper_service.cpp:84-120
  //   convert_mod_to_princ                        :122-154
  //   invariants_modified                         src/mat/4C_mat_service.cpp:125-134
  void lame_from_input(double mode, double c1, double c2, double& lambda, double& mue)
  {
    if (mode == 1.0)  // "Lame": C1 = mue, C2 = lambda (4C_mat_elast_couplogneohooke.cpp:33-37)
    {
      mue = c1, lambda = c2;
      return;
    }
    lambda = (c2 == 0.5) ? 0.0 : c1 * c2 / ((1.0 + c2) * (1.0 - 2.0 * c2));
    mue = c1 / (2.0 * (1.0 + c2));
  }

  // one summand: adds to dPI/ddPII (principal) or dPm/ddPm (modified) and to psi
  void summand_add(int kind, double mode, const double* p, const double* I, const double* m,
      double* dPI, double* ddPII, double* dPm, double* ddPm, double& psi)
  {
    using std::exp;
    using std::log;
    using std::pow;
    using std::sqrt;
    switch (kind)
    {
      case ORC_SUM_COUP_NEOHOOKE:  // src/mat/elast/4C_mat_elast_coupneohooke.cpp:17-68
      {
        const double c = p[0] / (4.0 * (1.0 + p[1])), beta = p[1] / (1.0 - 2.0 * p[1]);
        double psiadd = c * (I[0] - 3.);
        if (beta != 0)
          psiadd += (c / beta) * (exp(log(I[2]) * (-beta)) - 1.);
        else
          psiadd -= c * log(I[2]);
        psi += psiadd;
        dPI[0] += c;
        if (I[2] > 0)
        {
          const double prinv2_to_beta_m1 = exp(log(I[2]) * (-beta - 1.));
          dPI[2] -= c * prinv2_to_beta_m1;
          ddPII[2] += c * (beta + 1.) * prinv2_to_beta_m1 / I[2];
        }
        else
          dPI[2] = ddPII[2] = std::numeric_limits<double>::quiet_NaN();
        break;
      }
      case ORC_SUM_COUP_MOONEYRIVLIN:  // 4C_mat_elast_coupmooneyrivlin.cpp:29-64
      {
        const double c1 = p[0], c2 = p[1], c3 = p[2];
        psi += c1 * (I[0] - 3.) + c2 * (I[1] - 3.) - (2. * c1 + 4. * c2) * log(sqrt(I[2])) +
               c3 * pow((sqrt(I[2]) - 1.), 2.);
        dPI[0] += c1;
        dPI[1] += c2;
        dPI[2] += c3 * (1 - pow(I[2], -0.5)) - (c1 + 2. * c2) * pow(I[2], -1.);
        ddPII[2] += (c1 + 2 * c2) * pow(I[2], -2.) + 0.5 * c3 * pow(I[2], -1.5);
        break;
      }
      case ORC_SUM_COUP_LOGNEOHOOKE:  // 4C_mat_elast_couplogneohooke.cpp:17-82
      {
        double lambda, mue;
        lame_from_input(mode, p[0], p[1], lambda, mue);
        psi += mue * 0.5 * (I[0] - 3.) - mue * log(sqrt(I[2])) + lambda * 0.5 * pow(log(sqrt(I[2])), 2.);
        const double logdetf = log(sqrt(I[2]));
        dPI[0] += mue * 0.5;
        dPI[2] += (lambda * logdetf) / (2. * I[2]) - mue / (2. * I[2]);
        ddPII[2] += lambda / (4. * I[2] * I[2]) + mue / (2. * I[2] * I[2]) -
                    (lambda * logdetf) / (2. * I[2] * I[2]);
        break;
      }
      case ORC_SUM_COUP_LOGMIXNEOHOOKE:  // 4C_mat_elast_couplogmixneohooke.cpp:17-82
      {
        double lambda, mue;
        lame_from_input(mode, p[0], p[1], lambda, mue);
        const double sq = sqrt(I[2]);
        psi += mue * 0.5 * (I[0] - 3.) - mue * log(sq) + lambda * 0.5 * pow((sq - 1.), 2.);
        dPI[0] += mue * 0.5;
        dPI[2] += lambda * (sq - 1.) / (2. * sq) - mue / (2. * I[2]);
        ddPII[2] += lambda / (4. * I[2]) + mue / (2. * I[2] * I[2]) - lambda * (sq - 1.) / (4. * sq * sq * sq);
        break;
      }
      case ORC_SUM_COUP_SVK:  // 4C_mat_elast_coupSaintVenantKirchhoff.cpp:17-62
      {
        double lambda, mue;
        lame_from_input(0.0, p[0], p[1], lambda, mue);
        psi += (0.25 * mue + 0.125 * lambda) * I[0] * I[0] - (0.75 * lambda + 0.5 * mue) * I[0] -
               0.5 * mue * I[1] + 1.125 * lambda + 0.75 * mue;
        dPI[0] += (0.5 * mue + 0.25 * lambda) * I[0] - 0.75 * lambda - 0.5 * mue;
        dPI[1] -= 0.5 * mue;
        ddPII[0] += 0.5 * mue + 0.25 * lambda;
        break;
      }
      case ORC_SUM_COUP_SIMOPISTER:  // 4C_mat_elast_coupsimopister.cpp
      {
        const double mue = p[0];
        psi += 0.5 * mue * (I[0] - 3.) - mue * log(pow(I[2], 0.5));
        dPI[0] += 0.5 * mue;
        dPI[2] -= 0.5 * mue / I[2];
        ddPII[2] += 0.5 * mue / (I[2] * I[2]);
        break;
      }
      case ORC_SUM_COUP_BLATZKO:  // 4C_mat_elast_coupblatzko.cpp
      {
        const double mue = p[0], nue = p[1], f = p[2];
        const double beta = nue / (1. - 2. * nue);
        const double log_i3 = log(I[2]);
        const double neg = beta == 0.0 ? -log_i3 : std::expm1(-beta * log_i3) / beta;
        const double pos = beta == 0.0 ? log_i3 : std::expm1(beta * log_i3) / beta;
        psi += 0.5 * mue * (f * (I[0] - 3.0 + neg) + (1.0 - f) * (I[1] / I[2] - 3.0 + pos));
        dPI[0] += f * mue / 2.;
        dPI[1] += (1. - f) * mue / (2. * I[2]);
        dPI[2] += -(f * mue) / (2. * pow(I[2], beta + 1.)) -
                  (mue * (pow(I[2], beta - 1.) - I[1] / I[2] / I[2]) * (f - 1.)) * 0.5;
        ddPII[2] += (f * mue * (beta + 1.)) / (2. * pow(I[2], beta + 2.)) -
                    (mue * (f - 1.) *
                        ((2. * I[1]) / I[2] / I[2] / I[2] + pow(I[2], beta - 2.) * (beta - 1.))) *
                        0.5;
        ddPII[3] -= (1. - f) * 0.5 * mue / I[2] / I[2];
        break;
      }
      case ORC_SUM_COUP_EXPPOL:  // 4C_mat_elast_coupexppol.cpp
      {
        const double a = p[0], b = p[1], c = p[2];
        psi += a * exp(b * (I[0] - 3.0) - (2.0 * b + c) * log(sqrt(I[2])) + c * (sqrt(I[2]) - 1.0)) - a;
        const double logdetf = log(sqrt(I[2]));
        const double expfunc = exp(b * (I[0] - 3.0) - (2.0 * b + c) * logdetf + c * (sqrt(I[2]) - 1.0));
        dPI[0] += a * b * expfunc;
        dPI[2] += a * expfunc * (c / (2. * sqrt(I[2])) - (2. * b + c) / (2. * I[2]));
        ddPII[0] += a * b * b * expfunc;
        ddPII[2] += a * expfunc *
                    ((0.5 * (2. * b + c) / (I[2] * I[2])) - (0.25 * c / pow(I[2], 1.5)) +
                        (-0.5 * (2. * b + c) / I[2] + 0.5 * c / sqrt(I[2])) *
                            (-0.5 * (2. * b + c) / I[2] + 0.5 * c / sqrt(I[2])));
        ddPII[4] += a * b * expfunc * (c / (2. * sqrt(I[2])) - (2. * b + c) / (2. * I[2]));
        break;
      }
      case ORC_SUM_COUP_1POW:  // 4C_mat_elast_coup1pow.cpp
      case ORC_SUM_COUP_2POW:  // 4C_mat_elast_coup2pow.cpp
      {
        const int q = kind == ORC_SUM_COUP_1POW ? 0 : 1;
        const double c = p[0];
        const int d = (int)p[1];
        psi += c * pow((I[q] - 3.), d);
        dPI[q] += c * d * pow((I[q] - 3.), d - 1.);
        if (d == 2)
          ddPII[q] += (c * d * d - c * d);
        else
          ddPII[q] += (c * d * d - c * d) * pow((I[q] - 3.), d - 2.);
        break;
      }
      case ORC_SUM_COUP_3POW:  // 4C_mat_elast_coup3pow.cpp
      {
        const double c = p[0];
        const int d = (int)p[1];
        psi += c * pow((pow(I[2], 1. / 3.) - 1.), d);
        dPI[2] += 1. / 3. * c * d * pow(I[2], -2. / 3.) * pow((pow(I[2], 1. / 3.) - 1.), d - 1.);
        if (d == 2)
          ddPII[2] += -2. / 9. * c * d * pow(I[2], -5. / 3.) * pow((pow(I[2], 1. / 3.) - 1.), d - 1.) +
                      1. / 9. * c * d * (d - 1.) * pow(I[2], -4. / 3.);
        else
          ddPII[2] += -2. / 9. * c * d * pow(I[2], -5. / 3.) * pow((pow(I[2], 1. / 3.) - 1.), d - 1.) +
                      1. / 9. * c * d * (d - 1.) * pow(I[2], -4. / 3.) *
                          pow((pow(I[2], 1. / 3.) - 1.), d - 2.);
        break;
      }
      case ORC_SUM_COUP_13APOW:  // 4C_mat_elast_coup13apow.cpp
      {
        const double c = p[0], a = p[2];
        const int d = (int)p[1];
        psi += c * pow((I[0] * (pow(I[2], -a)) - 3.), d);
        const double I1I3a3 = I[0] * pow(I[2], -a) - 3.;
        dPI[0] += c * d * pow(I[2], -a) * pow(I1I3a3, d - 1.);
        dPI[2] += -a * c * d * I[0] * pow(I[2], -a - 1.) * pow(I1I3a3, d - 1.);
        if (d == 2)
        {
          ddPII[0] += c * d * (d - 1.) * pow(I[2], -2. * a);
          ddPII[2] += a * (a + 1.) * c * d * I[0] * pow(I[2], -a - 2.) * pow(I1I3a3, d - 1.) +
                      a * a * c * d * (d - 1.) * I[0] * I[0] * pow(I[2], (-2. * a - 2.));
          ddPII[4] += -a * c * d * pow(I[2], -a - 1.) * pow(I1I3a3, d - 1.) -
                      a * c * d * (d - 1.) * I[0] * pow(I[2], (-2. * a - 1.));
        }
        else
        {
          ddPII[0] += c * d * (d - 1.) * pow(I[2], -2. * a) * pow(I1I3a3, d - 2.);
          ddPII[2] += a * (a + 1.) * c * d * I[0] * pow(I[2], -a - 2.) * pow(I1I3a3, d - 1.) +
                      a * a * c * d * (d - 1.) * I[0] * I[0] * pow(I[2], (-2. * a - 2.)) *
                          pow(I1I3a3, d - 2.);
          ddPII[4] += -a * c * d * pow(I[2], -a - 1.) * pow(I1I3a3, d - 1.) -
                      a * c * d * (d - 1.) * I[0] * pow(I[2], (-2. * a - 1.)) * pow(I1I3a3, d - 2.);
        }
        break;
      }
      case ORC_SUM_ISO_NEOHOOKE:  // 4C_mat_elast_isoneohooke.cpp:22-40
        psi += p[0] * 0.5 * (m[0] - 3.);
        dPm[0] += 0.5 * p[0];
        break;
      case ORC_SUM_ISO_MOONEYRIVLIN:  // 4C_mat_elast_isomooneyrivlin.cpp
        psi += p[0] * (m[0] - 3.) + p[1] * (m[1] - 3.);
        dPm[0] += p[0];
        dPm[1] += p[1];
        break;
      case ORC_SUM_ISO_YEOH:  // 4C_mat_elast_isoyeoh.cpp
      {
        const double c1 = p[0], c2 = p[1], c3 = p[2];
        psi += c1 * (m[0] - 3.) + c2 * (m[0] - 3.) * (m[0] - 3.) +
               c3 * (m[0] - 3.) * (m[0] - 3.) * (m[0] - 3.);
        dPm[0] += c1 + 2. * c2 * (m[0] - 3.) + 3. * c3 * (m[0] - 3.) * (m[0] - 3.);
        ddPm[0] += 2. * c2 + 6. * c3 * (m[0] - 3.);
        break;
      }
      case ORC_SUM_ISO_1POW:  // 4C_mat_elast_iso1pow.cpp
      case ORC_SUM_ISO_2POW:  // 4C_mat_elast_iso2pow.cpp
      {
        const int q = kind == ORC_SUM_ISO_1POW ? 0 : 1;
        const double c = p[0];
        const int d = (int)p[1];
        psi += c * pow((m[q] - 3.), d);
        if (d == 1)
          dPm[q] += c * d;
        else
          dPm[q] += c * d * pow(m[q] - 3., d - 1.);
        if (d == 1)
          ddPm[q] += 0.;
        else if (d == 2)
          ddPm[q] += c * d * (d - 1.);
        else
          ddPm[q] += c * d * (d - 1.) * pow(m[q] - 3., d - 2.);
        break;
      }
      case ORC_SUM_ISO_EXPOPOW:  // 4C_mat_elast_isoexpopow.cpp
      {
        const double k1 = p[0], k2 = p[1];
        const int d = (int)p[2];
        if (k2 != 0) psi += k1 / (2. * k2) * (exp(k2 * pow(m[0] - 3., d)) - 1.);
        const double expf = exp(k2 * pow(m[0] - 3., (double)d));
        if (d == 1)
          dPm[0] += k1 * d * 0.5 * expf;
        else
          dPm[0] += k1 * d * 0.5 * expf * pow(m[0] - 3., d - 1.);
        if (d == 1)
          ddPm[0] += k1 * d * d * k2 * 0.5 * expf;
        else if (d == 2)
          ddPm[0] += k1 * d * d * k2 * 0.5 * expf * pow((m[0] - 3.), 2. * (d - 1.)) +
                     k1 * d * 0.5 * expf * (d - 1.);
        else
          ddPm[0] += k1 * d * d * k2 * 0.5 * expf * pow((m[0] - 3.), 2. * (d - 1.)) +
                     k1 * d * 0.5 * expf * (d - 1.) * pow((m[0] - 3.), (d - 2.));
        break;
      }
      case ORC_SUM_VOL_SUSSMANBATHE:  // 4C_mat_elast_volsussmanbathe.cpp:25-43
        psi += p[0] * 0.5 * (m[2] - 1.) * (m[2] - 1.);
        dPm[2] += p[0] * (m[2] - 1.);
        ddPm[2] += p[0];
        break;
      case ORC_SUM_VOL_PENALTY:  // 4C_mat_elast_volpenalty.cpp
      {
        const double eps = p[0], game = p[1];
        psi += eps * (pow(m[2], game) + pow(m[2], -game) - 2.);
        dPm[2] += eps * game * (pow(m[2], game - 1.) - pow(m[2], -game - 1.));
        ddPm[2] += eps * game * ((game - 1.) * pow(m[2], game - 2.) + (game + 1.) * pow(m[2], -game - 2.));
        break;
      }
      case ORC_SUM_VOL_OGDEN:  // 4C_mat_elast_vologden.cpp
      {
        const double kappa = p[0], beta = p[1];
        if (beta != 0)
          psi += kappa / (beta * beta) * (beta * log(m[2]) + pow(m[2], -beta) - 1.);
        else
          psi += kappa / 2. * pow(log(m[2]), 2.);
        if (beta != 0.)
        {
          dPm[2] += kappa / (beta * beta) * (beta / m[2] - beta * pow(m[2], -beta - 1.));
          ddPm[2] += kappa / (beta * beta) *
                     (-beta / (m[2] * m[2]) + beta * (beta + 1.) * pow(m[2], -beta - 2.));
        }
        else
        {
          dPm[2] += kappa * log(m[2]) / m[2];
          ddPm[2] += kappa * (1. - log(m[2])) / (m[2] * m[2]);
        }
        break;
      }
      case ORC_SUM_VOL_POW:  // 4C_mat_elast_volpow.cpp
      {
        const double a = p[0], expon = p[1];
        psi += a / (expon - 1.) * pow(m[2], -expon + 1.) + a * m[2];
        dPm[2] += -a * (pow(m[2], -expon) - 1.);
        ddPm[2] += expon * a * pow(m[2], -(expon + 1.));
        break;
      }
      default:
        break;
    }
  }

  void convert_mod_to_princ(const double* I, const double* dPm, const double* ddPm, double* dPI, double* ddPII)
  {
    using std::pow;
    dPI[0] += pow(I[2], -1. / 3.) * dPm[0];
    dPI[1] += pow(I[2], -2. / 3.) * dPm[1];
    dPI[2] += 0.5 * pow(I[2], -0.5) * dPm[2] - 1. / 3. * I[0] * pow(I[2], -4. / 3.) * dPm[0] -
              2. / 3. * I[1] * pow(I[2], -5. / 3.) * dPm[1];
    ddPII[0] += pow(I[2], -2. / 3.) * ddPm[0];
    ddPII[1] += pow(I[2], -4. / 3.) * ddPm[1];
    ddPII[2] += (1. / 9.) * pow(I[2], -8. / 3.) * I[0] * I[0] * ddPm[0] +
                (4. / 9.) * I[0] * I[1] * pow(I[2], -3.) * ddPm[5] -
                (1. / 3.) * pow(I[2], -11. / 6.) * I[0] * ddPm[4] +
                (4. / 9.) * pow(I[2], -7. / 3.) * I[0] * dPm[0] +
                (4. / 9.) * pow(I[2], -10. / 3.) * I[1] * I[1] * ddPm[1] -
                (2. / 3.) * pow(I[2], -13. / 6.) * I[1] * ddPm[3] +
                (10. / 9.) * pow(I[2], -8. / 3.) * I[1] * dPm[1] + 0.25 * pow(I[2], -1.) * ddPm[2] -
                0.25 * pow(I[2], -1.5) * dPm[2];
    ddPII[3] += -(1. / 3.) * pow(I[2], -2.) * I[0] * ddPm[5] -
                (2. / 3.) * pow(I[2], -7. / 3.) * I[1] * ddPm[1] + 0.5 * pow(I[2], -7. / 6.) * ddPm[3] -
                (2. / 3.) * pow(I[2], -5. / 3.) * dPm[1];
    ddPII[4] += -(1. / 3.) * pow(I[2], -5. / 3.) * I[0] * ddPm[0] -
                (2. / 3.) * pow(I[2], -2.) * I[1] * ddPm[5] + 0.5 * pow(I[2], -5. / 6.) * ddPm[4] -
                (1. / 3.) * pow(I[2], -4. / 3.) * dPm[0];
    ddPII[5] += pow(I[2], -1.) * ddPm[5];
  }

  // ELAST_CoupAnisoExpo with one element fibre (ORACLE ONLY: the reference's MULF known answers use
  // it, so the oracle needs it to pin its MULF restatement; the device path does not have it).
  // Packed as two consecutive entries: {ORC_SUM_COUP_ANISOEXPO, 0, K1, K2, K1COMP, K2COMP} and
  // {ORC_SUM_FIBER, 0, a0, a1, a2, 0} with the element's FIBER1.  Structural tensor A = a (x) a
  // ("Standard" strategy), scalar product 1
  // (src/mat/elast/4C_mat_elast_coupanisoexpobase.cpp:36-62,128-160, ..._coupanisoexpo.cpp:18-21).
  // CoupAnisoExpoBase at one point for a general structural tensor A (Voigt, tensor components) and scalar
  // product (src/mat/elast/4C_mat_elast_coupanisoexpobase.cpp:46-113): I4 = C : A, compression parameters for
  // I4 < 1; returns psi, dPsi/dI4 and d2Psi/dI4^2
  void aniso_expo_point(const double* k /*K1 K2 K1COMP K2COMP*/, const double* A, double scalarProduct, const double* C,
      double* psi, double* dPI, double* ddPII)
  {
    double I4 = 0.0;
    for (int v = 0; v < 6; ++v) I4 += C[v] * A[v] * (v < 3 ? 1.0 : 2.0);
    double k1 = k[0], k2 = k[1];
    if (I4 < 1.0) k1 = k[2], k2 = k[3];
    *psi = (k1 / (2.0 * k2)) * (std::exp(k2 * (I4 - scalarProduct) * (I4 - scalarProduct)) - 1.0);
    *dPI = k1 * (I4 - scalarProduct) * std::exp(k2 * (I4 - scalarProduct) * (I4 - scalarProduct));
    *ddPII = (1.0 + 2.0 * k2 * std::pow((I4 - scalarProduct), 2)) * k1 * std::exp(k2 * std::pow((I4 - scalarProduct), 2));
  }

  void aniso_expo_add(const double* par, const double* C, double* S, double* Cm, double* psi)
  {
    const int n = (int)par[0];
    for (int k = 0; k + 1 < n; ++k)
    {
      const double* q = par + 1 + 6 * k;
      if ((int)q[0] != ORC_SUM_COUP_ANISOEXPO) continue;
      const double* fq = q + 6;
      if ((int)fq[0] != ORC_SUM_FIBER) continue;
      // read_fibers normalises the input vector (src/solid_ele/4C_solid_ele_fiber.cpp:66-69); element fibres give the
      // structural tensor a (x) a and scalar product 1 (unittests/mat/elast/4C_elast_coupanisoexpo_elementfibers_test.cpp)
      const double nrm = std::sqrt(fq[2] * fq[2] + fq[3] * fq[3] + fq[4] * fq[4]);
      const double a[3] = {fq[2] / nrm, fq[3] / nrm, fq[4] / nrm};
      double A[6];
      for (int v = 0; v < 6; ++v) A[v] = a[VA[v]] * a[VB[v]];
      double w, dPI, ddPII;
      aniso_expo_point(q + 2, A, 1.0, C, &w, &dPI, &ddPII);
      if (psi) *psi += w;
      if (S)
      {
        // add_stress_aniso_principal (:115-160): S += 2 dPI A, C += 4 ddPII A (x) A
        const double gamma = 2. * dPI, delta = 4. * ddPII;
        for (int v = 0; v < 6; ++v) S[v] += gamma * A[v];
        for (int I = 0; I < 6; ++I)
          for (int J = 0; J < 6; ++J) Cm[I + 6 * J] += delta * A[I] * A[J];
      }
    }
  }

  // returns psi; dPI/ddPII may be null (strain energy only)
  double hyper_generic(const double* par, const double* I, double* dPI, double* ddPII)
  {
    const int n = (int)par[0];
    double d1[3] = {0, 0, 0}, d2[6] = {0, 0, 0, 0, 0, 0}, dPm[3] = {0, 0, 0}, ddPm[6] = {0, 0, 0, 0, 0, 0};
    const double m[3] = {I[0] * std::pow(I[2], -1. / 3.), I[1] * std::pow(I[2], -2. / 3.), std::pow(I[2], 1. / 2.)};
    double psi = 0.0;
    bool isomod = false;
    for (int k = 0; k < n; ++k)
    {
      const double* q = par + 1 + 6 * k;
      const int kind = (int)q[0];
      if (kind >= ORC_SUM_ISO_NEOHOOKE && kind <= ORC_SUM_VOL_POW) isomod = true;
      summand_add(kind, q[1], q + 2, I, m, d1, d2, dPm, ddPm, psi);
    }
    if (isomod) convert_mod_to_princ(I, dPm, ddPm, d1, d2);
    if (dPI)
      for (int i = 0; i < 3; ++i) dPI[i] = d1[i];
    if (ddPII)
      for (int i = 0; i < 6; ++i) ddPII[i] = d2[i];
    return psi;
  }

  // eigen decomposition of a symmetric 3x3 by cyclic Jacobi rotations (the reference calls
  // Core::LinAlg::eig = LAPACK dsyev; same decomposition): A = V diag(w) V^T, V column-major
  void sym_eig3(const double* Afull, double* w, double* V)
  {
    double A[9];
    for (int i = 0; i < 9; ++i) A[i] = Afull[i], V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep)
    {
      const double off = std::fabs(A[3]) + std::fabs(A[6]) + std::fabs(A[7]);
      if (off < 1e-300) break;
      for (int p = 0; p < 2; ++p)
        for (int q = p + 1; q < 3; ++q)
        {
          const double apq = A[p + 3 * q];
          if (apq == 0.0) continue;
          const double theta = (A[q + 3 * q] - A[p + 3 * p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < 3; ++k)
          {
            const double akp = A[k + 3 * p], akq = A[k + 3 * q];
            A[k + 3 * p] = c * akp - s * akq, A[k + 3 * q] = s * akp + c * akq;
          }
          for (int k = 0; k < 3; ++k)
          {
            const double apk = A[p + 3 * k], aqk = A[q + 3 * k];
            A[p + 3 * k] = c * apk - s * aqk, A[q + 3 * k] = s * apk + c * aqk;
          }
          for (int k = 0; k < 3; ++k)
          {
            const double vkp = V[k + 3 * p], vkq = V[k + 3 * q];
            V[k + 3 * p] = c * vkp - s * vkq, V[k + 3 * q] = s * vkp + c * vkq;
          }
        }
    }
    for (int i = 0; i < 3; ++i) w[i] = A[i + 3 * i];
  }

  // add_coefficients_stretches_modified of ELAST_IsoOgden (4C_mat_elast_isoogden.cpp:24-50) and
  // ELAST_IsoVarga (4C_mat_elast_isovarga.cpp): first / second derivatives w.r.t. the modified stretches
  void stretch_coefficients_modified(int kind, const double* p, const double* modstr, double* mg, double* md)
  {
    if (kind == ORC_SUM_ISO_OGDEN)
    {
      const double mue = p[0], alpha = p[1];
      for (int i = 0; i < 3; ++i)
      {
        mg[i] += 2 * mue / alpha * std::pow(modstr[i], alpha - 1);
        md[i] += 2 * mue / alpha * (alpha - 1) * std::pow(modstr[i], alpha - 2);
      }
    }
    else if (kind == ORC_SUM_ISO_VARGA)
    {
      const double alpha = 2.0 * p[0] - p[1], beta = p[1];
      for (int i = 0; i < 3; ++i)
      {
        mg[i] += alpha - beta / (modstr[i] * modstr[i]);
        md[i] += 2.0 * beta / (modstr[i] * modstr[i] * modstr[i]);
      }
    }
  }

  // Mat::elast_hyper_add_response_stretches (src/mat/4C_mat_elasthyper_service.cpp:208-386) for the summands
  // formulated in principal stretches (ELAST_CoupVarga, 4C_mat_elast_coupvarga.cpp) or modified principal
  // stretches (ELAST_IsoOgden, 4C_mat_elast_isoogden.cpp:24-50; ELAST_IsoVarga, 4C_mat_elast_isovarga.cpp);
  // stretches_modified: src/mat/4C_mat_service.cpp:156-164.  Adds to S and Cm; returns false if the block
  // holds no such summand.
  bool stretch_response_add(const double* par, const double* C, double* S, double* Cm)
  {
    const int n = (int)par[0];
    bool princ = false, mod = false;
    for (int k = 0; k < n; ++k)
    {
      const int kind = (int)par[1 + 6 * k];
      princ = princ || kind == ORC_SUM_COUP_VARGA;
      mod = mod || kind == ORC_SUM_ISO_OGDEN || kind == ORC_SUM_ISO_VARGA;
    }
    if (!princ && !mod) return false;
    double Cf[9], ev[3], N[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Cf[a + 3 * b] = C[VG[a][b]];
    sym_eig3(Cf, ev, N);  // N(i, al) = N[i + 3 * al]
    double prstr[3], modstr[3];
    for (int i = 0; i < 3; ++i) prstr[i] = std::sqrt(ev[i]);
    const double detF = prstr[0] * prstr[1] * prstr[2];
    for (int i = 0; i < 3; ++i) modstr[i] = std::pow(detF, -1.0 / 3.0) * prstr[i];
    double gamma[3] = {0, 0, 0}, delta[6] = {0, 0, 0, 0, 0, 0};
    if (princ)
      for (int k = 0; k < n; ++k)
      {
        const double* q = par + 1 + 6 * k;
        if ((int)q[0] != ORC_SUM_COUP_VARGA) continue;
        const double alpha = 2.0 * q[2] - q[3], beta = q[3];
        for (int i = 0; i < 3; ++i)
        {
          gamma[i] += alpha - beta / (prstr[i] * prstr[i]);
          delta[i] += 2.0 * beta / (prstr[i] * prstr[i] * prstr[i]);
        }
      }
    if (mod)
    {
      const double detF13 = std::pow(detF, -1.0 / 3.0);
      double mg[3] = {0, 0, 0}, md[6] = {0, 0, 0, 0, 0, 0};
      for (int k = 0; k < n; ++k)
      {
        const double* q = par + 1 + 6 * k;
        const int kind = (int)q[0];
        stretch_coefficients_modified(kind, q + 2, modstr, mg, md);
      }
      double modbypr[3][3];
      for (int al = 0; al < 3; ++al)
      {
        for (int be = 0; be < 3; ++be) modbypr[al][be] = -modstr[al] / modstr[be];
        modbypr[al][al] += 3.0;
      }
      for (int al = 0; al < 3; ++al)
        for (int be = 0; be < 3; ++be) modbypr[al][be] *= detF13 / 3.0;
      for (int be = 0; be < 3; ++be)
        for (int al = 0; al < 3; ++al) gamma[be] += modbypr[al][be] * mg[al];
      const double mdt[3][3] = {{md[0], md[3], md[5]}, {md[3], md[1], md[4]}, {md[5], md[4], md[2]}};
      double aux[3][3], dt[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
          aux[i][j] = 0.0;
          for (int k = 0; k < 3; ++k) aux[i][j] += modbypr[k][i] * mdt[k][j];
        }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
          dt[i][j] = 0.0;
          for (int k = 0; k < 3; ++k) dt[i][j] += aux[i][k] * modbypr[k][j];
        }
      for (int be = 0; be < 3; ++be)
        for (int ga = 0; ga < 3; ++ga)
        {
          double v = 0.0;
          for (int al = 0; al < 3; ++al)
          {
            v += -mg[al] * modbypr[al][be] / (3.0 * prstr[ga]);
            if (ga == al) v += -mg[al] * detF13 / (3.0 * prstr[be]);
            if (be == ga) v += mg[al] * detF13 * prstr[al] / (3.0 * prstr[be] * prstr[be]);
          }
          dt[be][ga] += v;
        }
      delta[0] += dt[0][0], delta[1] += dt[1][1], delta[2] += dt[2][2];
      delta[3] += dt[0][1], delta[4] += dt[1][2], delta[5] += dt[2][0];
    }
    // principal 2nd Piola-Kirchhoff stresses and the stress tensor
    double prsts[3];
    for (int al = 0; al < 3; ++al)
    {
      prsts[al] = gamma[al] / prstr[al];
      for (int v = 0; v < 6; ++v) S[v] += prsts[al] * N[VA[v] + 3 * al] * N[VB[v] + 3 * al];
    }
    double f1[6], f2[6] = {0, 0, 0, 0, 0, 0};
    for (int albe = 0; albe < 6; ++albe)
    {
      const int al = VA[albe], be = VB[albe];
      f1[albe] = delta[albe] / (prstr[al] * prstr[be]);
      if (albe < 3) f1[albe] -= gamma[al] / (prstr[be] * prstr[al] * prstr[al]);
    }
    for (int albe = 3; albe < 6; ++albe)
    {
      const int al = VA[albe], be = VB[albe];
      if (std::fabs(prstr[al] - prstr[be]) < 1e-6)
        f2[albe] = (f1[be] - f1[albe]) / 2.0;
      else
        f2[albe] = (prsts[be] - prsts[al]) / (prstr[be] * prstr[be] - prstr[al] * prstr[al]);
    }
    auto n_ = [&](int i, int al) { return N[i + 3 * al]; };
    for (int kl = 0; kl < 6; ++kl)
      for (int ij = 0; ij < 6; ++ij)
      {
        const int k = VA[kl], l = VB[kl], i = VA[ij], j = VB[ij];
        double c = 0.0;
        for (int albe = 0; albe < 6; ++albe)
        {
          const int al = VA[albe], be = VB[albe];
          c += f1[albe] * n_(i, al) * n_(j, al) * n_(k, be) * n_(l, be);
          if (albe >= 3)
          {
            c += f1[albe] * n_(i, be) * n_(j, be) * n_(k, al) * n_(l, al);
            c += f2[albe] * (n_(i, al) * n_(j, be) * n_(k, al) * n_(l, be) + n_(i, al) * n_(j, be) * n_(k, be) * n_(l, al) +
                             n_(i, be) * n_(j, al) * n_(k, be) * n_(l, al) + n_(i, be) * n_(j, al) * n_(k, al) * n_(l, be));
          }
        }
        Cm[ij + 6 * kl] += c;
      }
    return true;
  }

  // Mat::ElastHyper::evaluate -> elast_hyper_evaluate (src/mat/4C_mat_elasthyper.cpp:314-325,
  // src/mat/4C_mat_elasthyper_service.cpp:23-75,156-206)
  void elasthyper(int kind, const double* par, const double* E, double* S, double* Cm)
  {
    double C[6], Cfull[9], iCfull[9], iC[6];
    for (int i = 0; i < 6; ++i) C[i] = 2.0 * E[i] + (i < 3 ? 1.0 : 0.0);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Cfull[a + 3 * b] = C[VG[a][b]];
    inv3(iCfull, Cfull);
    for (int i = 0; i < 6; ++i) iC[i] = iCfull[VA[i] + 3 * VB[i]];
    // invariants_principal (src/core/linalg/src/dense/4C_linalg_fixedsizematrix_voigt_notation.hpp:78-103)
    double I[3];
    I[0] = C[0] + C[1] + C[2];
    I[1] = 0.5 * (I[0] * I[0] - C[0] * C[0] - C[1] * C[1] - C[2] * C[2]) - C[3] * C[3] -
           C[4] * C[4] - C[5] * C[5];
    I[2] = C[0] * C[1] * C[2] + 2 * (C[3] * C[4] * C[5]) - C[1] * C[5] * C[5] -
           C[2] * C[3] * C[3] - C[0] * C[4] * C[4];
    double dPI[3], ddPII[6], g[3], d[8];
    if (kind == ORC_MAT_ELASTHYPER)
      hyper_generic(par, I, dPI, ddPII);
    else
      invariant_derivatives(kind, par, I, dPI, ddPII);
    gamma_delta(I, dPI, ddPII, g, d);
    const double Id[6] = {1, 1, 1, 0, 0, 0};
    for (int i = 0; i < 6; ++i) S[i] = g[0] * Id[i] + g[1] * C[i] + g[2] * iC[i];
    for (int Iv = 0; Iv < 6; ++Iv)
      for (int Jv = 0; Jv < 6; ++Jv)
      {
        const int a = VA[Iv], b = VB[Iv], c = VA[Jv], e = VB[Jv];
        double v = 0.0;
        v += d[0] * Id[Iv] * Id[Jv];
        v += d[1] * (Id[Iv] * C[Jv] + C[Iv] * Id[Jv]);
        v += d[2] * (Id[Iv] * iC[Jv] + iC[Iv] * Id[Jv]);
        v += d[3] * C[Iv] * C[Jv];
        v += d[4] * (C[Iv] * iC[Jv] + iC[Iv] * C[Jv]);
        v += d[5] * iC[Iv] * iC[Jv];
        // Holzapfel product 1/2 (A_ac A_be + A_ae A_bc)
        // (src/core/linalg/src/dense/4C_linalg_fixedsizematrix_tensor_products.cpp:207-214)
        v += d[6] * 0.5 * (iC[VG[a][c]] * iC[VG[b][e]] + iC[VG[a][e]] * iC[VG[b][c]]);
        // symmetric identity 1/2 (d_ac d_be + d_ae d_bc)
        v += d[7] * 0.5 * ((a == c && b == e ? 1.0 : 0.0) + (a == e && b == c ? 1.0 : 0.0));
        Cm[Iv + 6 * Jv] = v;
      }
    if (kind == ORC_MAT_ELASTHYPER) stretch_response_add(par, C, S, Cm);
    if (kind == ORC_MAT_ELASTHYPER) aniso_expo_add(par, C, S, Cm, nullptr);
  }

  // Mat::StVenantKirchhoff::evaluate (src/mat/4C_mat_stvenantkirchhoff.cpp:105-113,
  // src/mat/4C_mat_stvenantkirchhoff.hpp:136-170)
  void svk(const double* par, const double* E, double* S, double* Cm)
  {
    const double Ey = par[0], nu = par[1];
    const double mue = Ey / (2 * (1 + nu));
    const double lambda = Ey * nu / ((1 + nu) * (1 - 2 * nu));
    const double tr = E[0] + E[1] + E[2];
    for (int i = 0; i < 6; ++i) S[i] = lambda * tr * (i < 3 ? 1.0 : 0.0) + 2 * mue * E[i];
    for (int i = 0; i < 36; ++i) Cm[i] = 0.0;
    const double mfac = Ey / ((1.0 + nu) * (1.0 - 2.0 * nu));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Cm[i + 6 * j] = (i == j) ? mfac * (1.0 - nu) : mfac * nu;
    for (int i = 3; i < 6; ++i) Cm[i + 6 * i] = mfac * 0.5 * (1.0 - 2.0 * nu);
  }

  // Mat::PlasticDruckerPrager::evaluate (src/mat/4C_mat_plasticdruckerprager.cpp:168-263) with
  // setup_cmat_elasto_plastic_cone / _apex (:127-166), the return-mapping residuals (:293-319) and
  // Core::Utils::solve_local_newton (src/core/utils/src/numerics/4C_utils_local_newton.hpp:194-223).
  // E, S: tensor components [00,11,22,01,12,02]; Cm 6x6 (NOT major-symmetric for ETA != ETABAR).
  // hist: {strainpllast[6], strainbarpllast, strainplcurr[6], strainbarplcurr} or null (virgin point).
  void drucker_prager(const double* par, double* hist, const double* E, double* S, double* Cm)
  {
    const double young = par[0], nu = par[1], Hiso = par[2], abstol = par[3], cohesion = par[4], eta = par[5],
                 xi = par[6], etabar = par[7];
    const int itermax = (int)par[8];
    const bool consistent = par[9] != 0.0;
    const double G = young / (2.0 * (1.0 + nu)), kappa = young / (3.0 * (1.0 - 2.0 * nu));
    const double id2[6] = {1, 1, 1, 0, 0, 0};
    double pl_last[6] = {0, 0, 0, 0, 0, 0}, bar_last = 0.0;
    if (hist)
    {
      for (int i = 0; i < 6; ++i) pl_last[i] = hist[i];
      bar_last = hist[6];
    }
    double trial[6], dev[6], devstress[6];
    for (int i = 0; i < 6; ++i) trial[i] = E[i] - pl_last[i];
    const double tr = trial[0] + trial[1] + trial[2];
    for (int i = 0; i < 6; ++i) dev[i] = trial[i] - id2[i] * (tr / 3.0);
    double p = kappa * tr;
    const double p_trial = p;
    for (int i = 0; i < 6; ++i) devstress[i] = 2.0 * G * dev[i];
    auto ddot = [](const double* a, const double* b)
    {
      double v = 0.0;
      for (int i = 0; i < 6; ++i) v += a[i] * b[i] * (i < 3 ? 1.0 : 2.0);
      return v;
    };
    const double J2 = 0.5 * ddot(devstress, devstress);
    double strainbar_p = bar_last;
    const double Phi_trial = std::sqrt(J2) + eta * p_trial - xi * (cohesion + Hiso * strainbar_p);
    double Dgamma = 0.0, dstrainv = 0.0;
    double pl_curr[6], bar_curr;
    auto newton = [&](auto fun, double x0, double tol)
    {
      double x = x0;
      auto rd = fun(x);
      int it = 0;
      while (std::fabs(rd.first) > tol)
      {
        if (it > itermax) break;  // the reference throws
        x -= rd.first / rd.second;
        rd = fun(x);
        ++it;
      }
      return x;
    };
    if (Phi_trial / std::fabs(cohesion) > abstol)
    {
      Dgamma = newton(
          [&](double x)
          {
            return std::make_pair(Phi_trial - x * (G + eta * kappa * etabar) - (xi * xi * x * Hiso),
                -G - (kappa * etabar * eta) - (xi * xi * Hiso));
          },
          Dgamma, abstol * cohesion);
      if ((std::sqrt(J2) - G * Dgamma) / std::fabs(cohesion) < abstol)
      {
        strainbar_p = bar_last;
        const double alpha = xi / eta, beta = xi / etabar;
        dstrainv = newton(
            [&](double x)
            {
              return std::make_pair(beta * cohesion + beta * strainbar_p * Hiso - p_trial + x * (alpha * beta * Hiso + kappa),
                  xi * xi / eta / etabar * Hiso + kappa);
            },
            dstrainv, abstol * cohesion);
        strainbar_p = bar_last + xi / eta * dstrainv;
        p = p_trial - kappa * dstrainv;
        for (int i = 0; i < 6; ++i) devstress[i] = 0.0;
      }
      else
      {
        strainbar_p = bar_last + xi * Dgamma;
        const double f = 1.0 - G * Dgamma / std::sqrt(J2);
        for (int i = 0; i < 6; ++i) devstress[i] *= f;
        p = p_trial - kappa * etabar * Dgamma;
      }
      for (int i = 0; i < 6; ++i) S[i] = devstress[i] + p * id2[i];
      for (int i = 0; i < 6; ++i) pl_curr[i] = E[i] - (devstress[i] * (1.0 / (2.0 * G)) + id2[i] * (p / (3.0 * kappa)));
      bar_curr = strainbar_p;
    }
    else
    {
      for (int i = 0; i < 6; ++i) S[i] = devstress[i] + p * id2[i], pl_curr[i] = pl_last[i];
      bar_curr = bar_last;
    }
    if (hist)
    {
      for (int i = 0; i < 6; ++i) hist[7 + i] = pl_curr[i];
      hist[13] = bar_curr;
    }
    if (Phi_trial > 0.0 && consistent)
    {
      // tensors in Voigt rows/columns of tensor components: Is(I,J) = 1 (normal), 1/2 (shear)
      if (dstrainv != 0.0)
      {
        const double epfac = kappa * (1.0 - kappa / (kappa + xi / eta * xi / etabar * Hiso));
        for (int I = 0; I < 6; ++I)
          for (int J = 0; J < 6; ++J) Cm[I + 6 * J] = epfac * id2[I] * id2[J];
      }
      else
      {
        const double nd = std::sqrt(ddot(dev, dev));
        double D[6];
        for (int i = 0; i < 6; ++i) D[i] = dev[i] / nd;
        const double A = 1.0 / (G + kappa * etabar * eta + xi * xi * Hiso);
        const double epfac = 2.0 * G * (1.0 - Dgamma / (std::sqrt(2.0) * nd));
        const double epfac2 = 2.0 * G * (Dgamma / (std::sqrt(2.0) * nd) - G * A);
        const double epfac3 = -std::sqrt(2.0) * G * A * kappa;
        const double epfac4 = kappa * (1.0 - kappa * eta * etabar * A);
        for (int I = 0; I < 6; ++I)
          for (int J = 0; J < 6; ++J)
          {
            const double Is = (I == J) ? (I < 3 ? 1.0 : 0.5) : 0.0;
            const double Idev = Is - id2[I] * id2[J] / 3.0;
            Cm[I + 6 * J] = epfac * Idev + epfac2 * D[I] * D[J] + epfac3 * (eta * D[I] * id2[J] + etabar * id2[I] * D[J]) +
                            epfac4 * id2[I] * id2[J];
          }
      }
    }
    else
    {
      double Sdummy[6];
      svk(par, E, Sdummy, Cm);  // StVenantKirchhoff::evaluate_stress_linearization(young, nu)
    }
  }

  // Mat::ViscoNeoHooke::evaluate (src/mat/4C_mat_visconeohooke.cpp:222-405, the GEN_MAXWELL branch that is
  // compiled in): compressible Neo-Hooke W = mue/2 (I1 I3^-1/3 - 3) + kappa/2 (J - 1)^2 on the FAST modulus, the
  // deviatoric stress relaxing through one Maxwell branch integrated with the one-step-theta rule
  //   Q^(n+1) = tau/(tau + theta dt) [ (tau - dt + theta dt)/tau Q^n + S_dev^(n+1) - S_dev^n ].
  // par: {YOUNGS_SLOW, POISSON, DENS, YOUNGS_FAST, RELAX, THETA, dt (EvaluationContext::time_step_size)}.
  // E, S: tensor components [00,11,22,01,12,02]; Cm 6x6 column-major in the same (stress-like) index pairs.
  // hist (ORC_VISCO_HIST = 24): {histstresslast[6], artstresslast[6], histstresscurr[6], artstresscurr[6]}, the
  // pack order (:97-104); null = virgin point, nothing stored.
  void visco_neohooke(const double* par, double* hist, const double* E, double* S, double* Cm)
  {
    const double E_s = par[0], nue = par[1], E_f = par[3], theta = par[5], dt = par[6];
    double tau = par[4];
    if (E_f > E_s) tau = tau * E_s / (E_f - E_s);  // algorithmic relaxation time (:244-249)
    const double alpha0 = E_s / E_f, alpha1 = 1.0 - alpha0;
    const double lambda = nue * E_f / ((1.0 + nue) * (1.0 - 2.0 * nue));
    const double mue = E_f / (2.0 * (1.0 + nue));
    const double kappa = lambda + 2.0 / 3.0 * mue;
    const double artscalar1 = (tau - dt + theta * dt) / tau;
    const double artscalar2 = tau / (tau + theta * dt);
    const double scalarvisco = alpha0 + alpha1 * tau / (tau + theta * dt);
    const double Id[6] = {1, 1, 1, 0, 0, 0};
    double c[6];  // right Cauchy-Green, tensor components (the reference's strain-like C(3..5) are 2 c[3..5])
    for (int v = 0; v < 6; ++v) c[v] = 2.0 * E[v] + Id[v];
    const double I1 = c[0] + c[1] + c[2];
    const double I3 = c[0] * c[1] * c[2] + 2.0 * c[3] * c[4] * c[5] - c[1] * c[5] * c[5] - c[2] * c[3] * c[3] -
                      c[0] * c[4] * c[4];
    const double J = std::sqrt(I3);
    const double I3invcubroot = std::pow(I3, -1.0 / 3.0);
    double Cinv[6] = {c[1] * c[2] - c[4] * c[4], c[0] * c[2] - c[5] * c[5], c[0] * c[1] - c[3] * c[3],
        c[5] * c[4] - c[3] * c[2], c[3] * c[5] - c[0] * c[4], c[3] * c[4] - c[5] * c[1]};
    for (double& v : Cinv) v *= 1.0 / I3;
    double SDevEla[6], Q[6];
    for (int v = 0; v < 6; ++v)
    {
      S[v] = kappa * (J - 1.0) * J * Cinv[v];                                    // volumetric part
      SDevEla[v] = mue * I3invcubroot * (Id[v] - 1.0 / 3.0 * I1 * Cinv[v]);       // 2 dW_dev/dC
      const double S_n = hist ? hist[v] : 0.0, Q_n = hist ? hist[6 + v] : 0.0;
      Q[v] = artscalar2 * (artscalar1 * Q_n + SDevEla[v] - S_n);
      if (hist) hist[12 + v] = SDevEla[v], hist[18 + v] = Q[v];
      S[v] += alpha0 * SDevEla[v] + alpha1 * Q[v];
    }
    const double scalar1 = 2.0 * kappa * J * J - kappa * J;
    const double scalar2 = -2.0 * kappa * J * J + 2.0 * kappa * J;
    const double scalar3 = 2.0 / 3.0 * mue * I3invcubroot * I1;
    const double scalar4 = 2.0 / 3.0 * mue * I3invcubroot;
    for (int I = 0; I < 6; ++I)
      for (int Jv = 0; Jv < 6; ++Jv)
      {
        const int a = VA[I], b = VB[I], cc = VA[Jv], d = VB[Jv];
        // holzapfel_product: 1/2 (Cinv_ac Cinv_bd + Cinv_ad Cinv_bc)
        const double H = 0.5 * (Cinv[VG[a][cc]] * Cinv[VG[b][d]] + Cinv[VG[a][d]] * Cinv[VG[b][cc]]);
        Cm[I + 6 * Jv] = (scalar2 + scalarvisco * scalar3) * H + scalar1 * Cinv[I] * Cinv[Jv] +
                         scalarvisco * (-scalar4) * (Id[I] * Cinv[Jv] + Id[Jv] * Cinv[I]) +
                         scalarvisco * scalar3 * Cinv[I] * Cinv[Jv] / 3.0;
      }
  }


  // Mat::PlasticNlnLogNeoHooke::evaluate (src/mat/4C_mat_plasticnlnlogneohooke.cpp:385-611): finite-strain von Mises
  // plasticity in principal logarithmic strains (de Souza Neto et al., ch. 14) -- trial elastic left Cauchy-Green
  // tensor b_e^trial = F Cp_n^-1 F^T, its spectral decomposition, return mapping of the deviatoric Kirchhoff stress with
  // the nonlinear isotropic hardening y = YIELD + ISOHARD ep + (SATHARDENING - YIELD)(1 - exp(-HARDEXPO ep)) and the
  // visco-plastic factor (VISC Dgamma / dt + 1)^RATE_DEPENDENCY (local Newton :117-147, solve_local_newton of
  // 4C_utils_local_newton.hpp: |r| > TOL, at most 50 iterations), S = sum_a tau_a N_a (x) N_a with N_a = F^-1 n_a and the
  // tangent in the principal directions (:520-571).
  // par: {YOUNG, NUE, DENS, YIELD, ISOHARD, SATHARDENING, HARDEXPO, VISC, RATE_DEPENDENCY, TOL, dt}; HARDENING_FUNC = 0.
  // hist (ORC_PLASTIC_HIST = 14): {accplstrainlast, invplrcglast[6], accplstraincurr, invplrcgcurr[6]} (pack order
  // :96-104; tensor components [00,11,22,01,12,02]); a fresh point holds Cp^-1 = I (setup, :337-349); null = virgin
  // point, nothing stored.  Returns 1 if the local Newton did not converge (the reference throws).
  int plastic_nlnlog_neohooke(const double* par, double* hist, const double* F, double* S, double* Cm)
  {
    const double ym = par[0], nu = par[1], yield = par[3], isohard = par[4], infyield = par[5], hardexp = par[6],
                 visc = par[7], eps = par[8], tol = par[9], dt = par[10];
    const double G = ym / (2.0 * (1.0 + nu)), kappa = ym / (3.0 * (1.0 - 2.0 * nu));
    const double detF = det3(F);
    double iF[9];
    inv3(iF, F);
    const double Id[6] = {1, 1, 1, 0, 0, 0};
    const double accpl_last = hist ? hist[0] : 0.0;
    const double* iCp = hist ? hist + 1 : Id;
    // Be_trial = F Cp^-1 F^T
    double Bt[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d) v += F[a + 3 * c] * iCp[VG[c][d]] * F[b + 3 * d];
        Bt[a + 3 * b] = v;
      }
    for (int a = 0; a < 3; ++a)
      for (int b = a + 1; b < 3; ++b) Bt[a + 3 * b] = Bt[b + 3 * a] = 0.5 * (Bt[a + 3 * b] + Bt[b + 3 * a]);
    double lam2[3], n[9];
    sym_eig3(Bt, lam2, n);  // n_a = n[. + 3 a]
    double N[3][3];          // N_a = F^-1 n_a (:449-452)
    for (int a = 0; a < 3; ++a)
      for (int i = 0; i < 3; ++i) N[a][i] = iF[i] * n[3 * a] + iF[i + 3] * n[1 + 3 * a] + iF[i + 6] * n[2 + 3 * a];
    double dev_trial[3], dev[3];
    for (int a = 0; a < 3; ++a) dev_trial[a] = G * std::log(lam2[a]) - 2.0 / 3.0 * G * std::log(detF);
    const double pressure = kappa * std::log(detF);
    const double abs_trial = std::sqrt(dev_trial[0] * dev_trial[0] + dev_trial[1] * dev_trial[1] + dev_trial[2] * dev_trial[2]);
    const double y_d = yield + isohard * accpl_last + (infyield - yield) * (1.0 - std::exp(-hardexp * accpl_last));
    const double f_trial = std::sqrt(3.0 / 2.0) * abs_trial - y_d;  // (visc Dgamma / dt + 1)^eps = 1 at Dgamma = 0
    double D[3][3];  // D_ep_principal = kappa ones + 2 G Idev
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) D[a][b] = kappa + 2.0 * G * ((a == b ? 1.0 : 0.0) - 1.0 / 3.0);
    double Dgamma = 0.0, lam2_new[3] = {lam2[0], lam2[1], lam2[2]};
    int fail = 0;
    if (f_trial <= 0.0)
      for (int a = 0; a < 3; ++a) dev[a] = dev_trial[a];
    else
    {
      // residuum_and_jacobian (:63-92) inside solve_local_newton
      auto res_jac = [&](double Dg, double& r, double& jac, double& dy)
      {
        const double acc = accpl_last + Dg;
        const double yd = yield + isohard * acc + (infyield - yield) * (1.0 - std::exp(-hardexp * acc));
        const double ydv = yd * std::pow(visc * Dg / dt + 1.0, eps);
        dy = (isohard + (infyield - yield) * hardexp * std::exp(-hardexp * acc)) * std::pow(visc * Dg / dt + 1.0, eps) +
             yd * eps * std::pow(visc * Dg / dt + 1.0, eps - 1.0) * visc / dt;
        r = std::sqrt(3.0 / 2.0) * abs_trial - 3.0 * G * Dg - ydv;
        jac = -3.0 * G - dy;
      };
      double r, jac, dy;
      res_jac(Dgamma, r, jac, dy);
      unsigned it = 0;
      while (std::fabs(r) > tol)
      {
        if (it > 50)
        {
          fail = 1;
          break;
        }
        Dgamma -= r / jac;
        res_jac(Dgamma, r, jac, dy);
        ++it;
      }
      double flow[3];
      for (int a = 0; a < 3; ++a) flow[a] = dev_trial[a] / (std::sqrt(2.0 / 3.0) * abs_trial);
      for (int a = 0; a < 3; ++a)
      {
        dev[a] = (1.0 - 2.0 * G * Dgamma / (std::sqrt(2.0 / 3.0) * abs_trial)) * dev_trial[a];
        lam2_new[a] = lam2[a] / std::exp(2.0 * flow[a] * Dgamma);
      }
      const double f1 = -4.0 * G * G * Dgamma / (std::sqrt(2.0 / 3.0) * abs_trial);
      const double f2 = 4.0 * G * G * (std::sqrt(2.0 / 3.0) * Dgamma / abs_trial - 1.0 / (3.0 * G + dy));
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) D[a][b] += f1 * ((a == b ? 1.0 : 0.0) - 1.0 / 3.0) + f2 * flow[a] * flow[b];
    }
    // S = sum_a (tau'_a + p) N_a (x) N_a (:513-516)
    for (int v = 0; v < 6; ++v)
    {
      double sv = 0.0;
      for (int a = 0; a < 3; ++a) sv += (dev[a] + pressure) * N[a][VA[v]] * N[a][VB[v]];
      S[v] = sv;
    }
    // tangent (:520-571); cmat is a SymmetricTensor<3,3,3,3>: assume_symmetry keeps the minor-symmetric part
    for (int I = 0; I < 6; ++I)
      for (int J = 0; J < 6; ++J)
      {
        const int i = VA[I], j = VB[I], k = VA[J], l = VB[J];
        double c = 0.0;
        for (int a = 0; a < 3; ++a)
        {
          c += -2.0 * (dev[a] + pressure) * N[a][i] * N[a][j] * N[a][k] * N[a][l];
          for (int b = 0; b < 3; ++b)
          {
            c += D[a][b] * N[a][i] * N[a][j] * N[b][k] * N[b][l];
            if (a == b) continue;
            const double fac = (lam2[a] != lam2[b])
                                   ? ((dev[a] + pressure) * lam2[b] - (dev[b] + pressure) * lam2[a]) / (lam2[a] - lam2[b])
                                   : 0.5 * (D[b][b] - D[a][b]) - (dev[b] + pressure);
            // N_ab (x) N_ab + N_ab (x) N_ab^T, N_ab = N_a (x) N_b, symmetrised over (i,j) and (k,l)
            const double t1 = 0.25 * (N[a][i] * N[b][j] + N[a][j] * N[b][i]) * (N[a][k] * N[b][l] + N[a][l] * N[b][k]);
            const double t2 = 0.25 * (N[a][i] * N[b][j] + N[a][j] * N[b][i]) * (N[b][k] * N[a][l] + N[b][l] * N[a][k]);
            c += fac * (t1 + t2);
          }
        }
        Cm[I + 6 * J] = c;
      }
    if (hist)
    {
      // invplrcgcurr = F^-1 Be F^-T = sum_a lambda_a,n+1^2 N_a (x) N_a (:575-581)
      hist[7] = accpl_last + Dgamma;
      for (int v = 0; v < 6; ++v)
      {
        double sv = 0.0;
        for (int a = 0; a < 3; ++a) sv += lam2_new[a] * N[a][VA[v]] * N[a][VB[v]];
        hist[8 + v] = sv;
      }
    }
    return fail;
  }

  thread_local double* g_material_hist = nullptr;  // history of the point being evaluated (set by the element loop)
  // finite-strain history material in the nonlinear / EAS element loops: history block of the element being
  // evaluated (ORC_VISCO_HIST doubles per Gauss point) or null
  thread_local double* g_element_material_hist = nullptr;
  double* g_material_hist_base = nullptr;  // orc_set_material_history: [n_ele][ngp][ORC_VISCO_HIST]
  void material(int kind, const double* par, const double* F, const double* E, double* S,
      double* Cm)
  {
    if (kind == ORC_MAT_PLASTIC_NLNLOG)
    {
      // The response depends on F only through C = 2 E + I (objectivity: N_a = F^-1 n_a and the eigenvalues of
      // F Cp^-1 F^T are those of Cp^-1 C).  Callers whose strain is not F^T F (EAS: the reference hands in the enhanced
      // deformation gradient R U_enh, eas_helpers.hpp:244-262) are served by ANY F with F^T F = C: the Cholesky factor.
      double c[6], Fu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      bool same = F != nullptr;
      for (int v = 0; v < 6; ++v)
      {
        c[v] = 2.0 * E[v] + (v < 3 ? 1.0 : 0.0);
        if (F)
        {
          double sC = 0.0;
          for (int q = 0; q < 3; ++q) sC += F[q + 3 * VA[v]] * F[q + 3 * VB[v]];
          if (std::fabs(sC - c[v]) > 1e-13 * (1.0 + std::fabs(c[v]))) same = false;
        }
      }
      if (!same)
      {
        // C = L L^T, F := L^T (upper triangular): Fu(a, b) = L(b, a)
        const double l00 = std::sqrt(c[0]), l10 = c[3] / l00, l20 = c[5] / l00;
        const double l11 = std::sqrt(c[1] - l10 * l10), l21 = (c[4] - l20 * l10) / l11;
        const double l22 = std::sqrt(c[2] - l20 * l20 - l21 * l21);
        Fu[0] = l00, Fu[0 + 3 * 1] = l10, Fu[0 + 3 * 2] = l20, Fu[1 + 3 * 1] = l11, Fu[1 + 3 * 2] = l21, Fu[2 + 3 * 2] = l22;
      }
      plastic_nlnlog_neohooke(par, g_material_hist, same ? F : Fu, S, Cm);
      return;
    }
    if (kind == ORC_MAT_SVK)
      svk(par, E, S, Cm);
    else if (kind == ORC_MAT_DRUCKER_PRAGER)
      drucker_prager(par, g_material_hist, E, S, Cm);
    else if (kind == ORC_MAT_VISCONEOHOOKE)
      visco_neohooke(par, g_material_hist, E, S, Cm);
    else
      elasthyper(kind, par, E, S, Cm);
  }

  // ---- B-free stiffness and force at one point (src/solid_ele/4C_solid_ele_calc_lib.hpp:1107-1151)
  void add_force(int n, const JacMap& jm, const double* F, const double* S, double fac, double* fe)
  {
    double P[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int c = 0; c < 3; ++c) v += F[a + 3 * c] * S[VG[c][b]];
        P[a + 3 * b] = v;
      }
    for (int k = 0; k < n; ++k)
      for (int a = 0; a < 3; ++a)
      {
        double v = 0.0;
        for (int b = 0; b < 3; ++b) v += P[a + 3 * b] * jm.G[k][b];
        fe[3 * k + a] += v * fac;
      }
  }

  void add_stiffness_bfree(int n, const JacMap& jm, const double* F, const double* S,
      const double* Cm, double fac, double* Ke)
  {
    const int nd = 3 * n;
    // FCFTw(a,b,c,d) = sum_ij F(b,i) C(a,i,j,d) F(c,j) w   (einsum<"bi","aijd","cj">, free
    // indices sorted by letter: 4C_linalg_tensor_einstein.hpp:222)
    double A[3][3][3][3];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d)
          {
            double v = 0.0;
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 3; ++j)
                v += F[b + 3 * i] * Cm[VG[a][i] + 6 * VG[j][d]] * (F[c + 3 * j] * fac);
            A[a][b][c][d] = v;
          }
    for (int i = 0; i < n; ++i)
    {
      double Ai[3][3][3], gi[3];
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c)
          for (int d = 0; d < 3; ++d)
          {
            double v = 0.0;
            for (int a = 0; a < 3; ++a) v += jm.G[i][a] * A[a][b][c][d];
            Ai[b][c][d] = v;
          }
      for (int d = 0; d < 3; ++d)
      {
        double v = 0.0;
        for (int a = 0; a < 3; ++a) v += jm.G[i][a] * S[VG[a][d]];
        gi[d] = v * fac;
      }
      for (int j = 0; j < n; ++j)
      {
        for (int b = 0; b < 3; ++b)
          for (int c = 0; c < 3; ++c)
          {
            double v = 0.0;
            for (int d = 0; d < 3; ++d) v += Ai[b][c][d] * jm.G[j][d];
            Ke[(3 * i + b) + nd * (3 * j + c)] += v;
          }
        double sg = 0.0;
        for (int d = 0; d < 3; ++d) sg += gi[d] * jm.G[j][d];
        for (int b = 0; b < 3; ++b) Ke[(3 * i + b) + nd * (3 * j + b)] += sg;
      }
    }
  }

  // B-operator variant (evaluate_strain_gradient :777-826, add_elastic_stiffness_matrix
  // :1163-1176, add_geometric_stiffness_matrix :1188-1207)
  void add_stiffness_bop(int n, const JacMap& jm, const double* F, const double* S,
      const double* Cm, double fac, double* Ke)
  {
    const int nd = 3 * n;
    std::vector<double> B(6 * nd), CB(6 * nd);
    for (int i = 0; i < n; ++i)
    {
      const double* g = jm.G[i];
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) B[d + 6 * (3 * i + e)] = F[e + 3 * d] * g[d];
      for (int e = 0; e < 3; ++e)
      {
        B[3 + 6 * (3 * i + e)] = F[e + 3 * 0] * g[1] + F[e + 3 * 1] * g[0];
        B[4 + 6 * (3 * i + e)] = F[e + 3 * 1] * g[2] + F[e + 3 * 2] * g[1];
        B[5 + 6 * (3 * i + e)] = F[e + 3 * 2] * g[0] + F[e + 3 * 0] * g[2];
      }
    }
    for (int c = 0; c < nd; ++c)
      for (int I = 0; I < 6; ++I)
      {
        double v = 0.0;
        for (int J = 0; J < 6; ++J) v += Cm[I + 6 * J] * B[J + 6 * c];
        CB[I + 6 * c] = v;
      }
    for (int r = 0; r < nd; ++r)
      for (int c = 0; c < nd; ++c)
      {
        double v = 0.0;
        for (int I = 0; I < 6; ++I) v += B[I + 6 * r] * CB[I + 6 * c];
        Ke[r + nd * c] += fac * v;
      }
    for (int i = 0; i < n; ++i)
    {
      double gi[3];
      for (int d = 0; d < 3; ++d)
      {
        double v = 0.0;
        for (int a = 0; a < 3; ++a) v += jm.G[i][a] * S[VG[a][d]];
        gi[d] = v * fac;
      }
      for (int j = 0; j < n; ++j)
      {
        double sg = 0.0;
        for (int d = 0; d < 3; ++d) sg += gi[d] * jm.G[j][d];
        for (int b = 0; b < 3; ++b) Ke[(3 * i + b) + nd * (3 * j + b)] += sg;
      }
    }
  }

  // ---- SolidEleCalc<celltype, DisplacementBased>::evaluate_nonlinear_force_stiffness_mass
  // (src/solid_ele/4C_solid_ele_calc.cpp:137-281 with the formulation of
  //  src/solid_ele/4C_solid_ele_calc_displacement_based.hpp:44-63,214-236)
  int element(int ct, int matkind, const double* par, int action, const double* X, const double* u,
      double* Ke, double* fe)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);  // clear_element_storage
    std::memset(fe, 0, sizeof(double) * nd);

    // ensure_positive_jacobian_determinant_at_element_nodes (calc_lib.hpp:543-562)
    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;

    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    // for_each_gauss_point (calc_lib.hpp:1254-1286)
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      double F[9], C[6], E[6], S[6], Cm[36];
      deformation_gradient(n, u, jm, F);
      // evaluate_cauchy_green / evaluate_green_lagrange_strain (calc_lib.hpp:697-722)
      for (int v = 0; v < 6; ++v)
      {
        double s = 0.0;
        for (int c = 0; c < 3; ++c) s += F[c + 3 * VA[v]] * F[c + 3 * VB[v]];
        C[v] = s;
        E[v] = 0.5 * (s - (v < 3 ? 1.0 : 0.0));
      }
      g_material_hist = (g_element_material_hist && finite_strain_hist(matkind))
                            ? g_element_material_hist + finite_strain_hist(matkind) * gp
                            : nullptr;
      material(matkind, par, F, E, S, Cm);
      g_material_hist = nullptr;
      add_force(n, jm, F, S, fac, fe);
      if (action == ORC_ACTION_FORCE_STIFF && Ke) add_stiffness_bfree(n, jm, F, S, Cm, fac, Ke);
    }
    return 0;
  }

  // ---- SolidEleCalc<celltype, DisplacementBasedLinearKinematicsFormulation>
  //      (src/solid_ele/4C_solid_ele_calc_displacement_based_linear_kinematics.hpp:44-62,101-150)
  int element_linear(int ct, int matkind, const double* par, int action, const double* X,
      const double* u, double* Ke, double* fe, double* hist = nullptr)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);
    std::memset(fe, 0, sizeof(double) * nd);
    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;
    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    const double Id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      // evaluate_linear_gl_strain (calc_lib.hpp:752-761): H = u N_XYZ^T, E = 1/2 (H + H^T)
      double H[9], E[6], S[6], Cm[36];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int k = 0; k < n; ++k) v += u[3 * k + a] * jm.G[k][b];
          H[a + 3 * b] = v;
        }
      for (int v = 0; v < 6; ++v) E[v] = 0.5 * (H[VA[v] + 3 * VB[v]] + H[VB[v] + 3 * VA[v]]);
      g_material_hist = (hist && matkind == ORC_MAT_DRUCKER_PRAGER) ? hist + ORC_DP_HIST * gp : nullptr;
      material(matkind, par, Id, E, S, Cm);
      g_material_hist = nullptr;
      // add_internal_force_vector(jacobian_mapping, F = I, pk2) (:122-131)
      add_force(n, jm, Id, S, fac, fe);
      if (!(action == ORC_ACTION_FORCE_STIFF && Ke)) continue;
      // add_stiffness_matrix (:133-150): K_ij(b,c) += sum_ad N_XYZ_i(a) C(a,b,c,d) N_XYZ_j(d) w
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int b = 0; b < 3; ++b)
            for (int c = 0; c < 3; ++c)
            {
              double v = 0.0;
              for (int a = 0; a < 3; ++a)
                for (int d = 0; d < 3; ++d) v += jm.G[i][a] * Cm[VG[a][b] + 6 * VG[c][d]] * jm.G[j][d];
              Ke[(3 * i + b) + nd * (3 * j + c)] += fac * v;
            }
    }
    return 0;
  }

  // evaluate_strain_gradient (src/solid_ele/4C_solid_ele_calc_lib.hpp:777-826): B-operator, 6 x 3n
  // column-major, rows in the order [00,11,22,01,12,02] (strain-like: shear rows hold both terms)
  void strain_gradient(int n, const JacMap& jm, const double* F, double* B)
  {
    for (int i = 0; i < n; ++i)
    {
      const double* g = jm.G[i];
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) B[d + 6 * (3 * i + e)] = F[e + 3 * d] * g[d];
      for (int e = 0; e < 3; ++e)
      {
        B[3 + 6 * (3 * i + e)] = F[e + 3 * 0] * g[1] + F[e + 3 * 1] * g[0];
        B[4 + 6 * (3 * i + e)] = F[e + 3 * 1] * g[2] + F[e + 3 * 2] * g[1];
        B[5 + 6 * (3 * i + e)] = F[e + 3 * 2] * g[0] + F[e + 3 * 0] * g[2];
      }
    }
  }

  // add_geometric_stiffness_matrix (calc_lib.hpp:1188-1207)
  void add_geometric_stiffness(int n, const JacMap& jm, const double* S, double fac, double* Ke)
  {
    const int nd = 3 * n;
    for (int i = 0; i < n; ++i)
    {
      double gi[3];
      for (int d = 0; d < 3; ++d)
      {
        double v = 0.0;
        for (int a = 0; a < 3; ++a) v += jm.G[i][a] * S[VG[a][d]];
        gi[d] = v * fac;
      }
      for (int j = 0; j < n; ++j)
      {
        double sg = 0.0;
        for (int d = 0; d < 3; ++d) sg += gi[d] * jm.G[j][d];
        for (int b = 0; b < 3; ++b) Ke[(3 * i + b) + nd * (3 * j + b)] += sg;
      }
    }
  }

  // Green-Lagrange strain and right Cauchy-Green tensor of F (calc_lib.hpp:697-722)
  void cauchy_green(const double* F, double* C, double* E)
  {
    for (int v = 0; v < 6; ++v)
    {
      double s = 0.0;
      for (int c = 0; c < 3; ++c) s += F[c + 3 * VA[v]] * F[c + 3 * VB[v]];
      C[v] = s;
      if (E) E[v] = 0.5 * (s - (v < 3 ? 1.0 : 0.0));
    }
  }

  // evaluate_parameter_coordinate_centroid (src/solid_ele/4C_solid_ele_calc_lib.hpp:281-334): hex / nurbs 0, tet 1/4,
  // pyramid (0, 0, 1/4), wedge (1/3, 1/3, 0)
  void centroid_xi(int ct, double* xi)
  {
    xi[0] = xi[1] = xi[2] = 0.0;
    if (ct == ORC_TET4 || ct == ORC_TET10) xi[0] = xi[1] = xi[2] = 0.25;
    if (ct == ORC_PYRAMID5) xi[2] = 0.25;
    if (ct == ORC_WEDGE6) xi[0] = xi[1] = 1.0 / 3.0;
  }

  void mulf_kinematics(int n, const double* dN, const double* u, const JacMap& jm, double* h, double* Fdelta, double* F);

  // ---- SolidEleCalc<hex8 | pyramid5, FBarFormulation>  (src/solid_ele/4C_solid_ele_calc_fbar.hpp:58-180,
  //      src/solid_ele/4C_solid_ele_calc_lib_fbar.hpp:47-160) and, with hist != nullptr,
  //      SolidEleCalc<hex8 | pyramid5, MulfFBarFormulation> (src/solid_ele/4C_solid_ele_calc_mulf_fbar.hpp:103-243):
  //      hist = ngauss Gauss-point histories followed by the element's global (centroid) history, ORC_MULF_HIST each;
  //      the deformation gradients at the points and at the centroid are the MULF ones, everything else -- B-operator and
  //      H-operator on the reference gradients N_XYZ / N_XYZ_0, the three stiffness terms -- is the F-bar formulation.
  int element_fbar(int ct, int matkind, const double* par, int action, const double* X,
      const double* u, double* Ke, double* fe, double* hist = nullptr)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN];
    JacMap jm, jm0;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);
    std::memset(fe, 0, sizeof(double) * nd);

    // prepare(): Jacobian mapping and deformation gradient at the element centroid (:58-66; mulf_fbar.hpp:118-149)
    double xi0[3];
    centroid_xi(ct, xi0);
    shape(ct, xi0, N, dN);
    jacobian_mapping(n, dN, X, jm0);
    double F0[9], iF0[9], Fd0[9];
    if (hist)
      mulf_kinematics(n, dN, u, jm0, hist + ORC_MULF_HIST * ngauss(ct), Fd0, F0);
    else
      deformation_gradient(n, u, jm0, F0);
    const double detF0 = det3(F0);
    inv3(iF0, F0);

    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;

    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    std::vector<double> B(6 * nd), CB(6 * nd), Hop(nd), bops(nd), bopccg(nd);
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      double F[9], iF[9], C[6], Fbar[9], Cbar[6], Ebar[6], S[6], Cm[36], Fd[9];
      if (hist)
        mulf_kinematics(n, dN, u, jm, hist + ORC_MULF_HIST * gp, Fd, F);
      else
        deformation_gradient(n, u, jm, F);
      const double detF = det3(F);
      inv3(iF, F);
      // evaluate_fbar_factor (lib_fbar.hpp:47-51)
      const double fbar = std::pow(detF0 / detF, 1.0 / 3.0);
      strain_gradient(n, jm, F, B.data());
      // evaluate_fbar_h_operator (lib_fbar.hpp:71-104)
      for (int idof = 0; idof < nd; ++idof)
      {
        const int node = idof / 3, c = idof % 3;
        double h = 0.0;
        for (int d = 0; d < 3; ++d)
        {
          h += iF0[d + 3 * c] * jm0.G[node][d];
          h -= iF[d + 3 * c] * jm.G[node][d];
        }
        Hop[idof] = h;
      }
      cauchy_green(F, C, nullptr);
      // F_bar = (detF_0/detF)^1/3 F and its strains (fbar.hpp:103-112)
      for (int i = 0; i < 9; ++i) Fbar[i] = fbar * F[i];
      cauchy_green(Fbar, Cbar, Ebar);
      material(matkind, par, Fbar, Ebar, S, Cm);

      // add_internal_force_vector(Bop, stress, fac / fbar) (fbar.hpp:152-162)
      for (int r = 0; r < nd; ++r)
      {
        double v = 0.0;
        for (int I = 0; I < 6; ++I) v += B[I + 6 * r] * S[I];
        fe[r] += (fac / fbar) * v;
      }
      if (!(action == ORC_ACTION_FORCE_STIFF && Ke)) continue;
      // add_elastic_stiffness_matrix(Bop, stress, fac * fbar) (fbar.hpp:173-174)
      for (int c = 0; c < nd; ++c)
        for (int I = 0; I < 6; ++I)
        {
          double v = 0.0;
          for (int J = 0; J < 6; ++J) v += Cm[I + 6 * J] * B[J + 6 * c];
          CB[I + 6 * c] = v;
        }
      for (int r = 0; r < nd; ++r)
        for (int c = 0; c < nd; ++c)
        {
          double v = 0.0;
          for (int I = 0; I < 6; ++I) v += B[I + 6 * r] * CB[I + 6 * c];
          Ke[r + nd * c] += (fac * fbar) * v;
        }
      // add_geometric_stiffness_matrix(jacobian_mapping, pk2, fac / fbar) (fbar.hpp:175-176)
      add_geometric_stiffness(n, jm, S, fac / fbar, Ke);
      // add_fbar_stiffness_matrix (lib_fbar.hpp:121-160): ccg = cmat : C (C of the unscaled F)
      double ccg[6];
      for (int I = 0; I < 6; ++I)
      {
        double v = 0.0;
        for (int J = 0; J < 6; ++J) v += Cm[I + 6 * J] * C[J] * (J < 3 ? 1.0 : 2.0);
        ccg[I] = v;
      }
      for (int r = 0; r < nd; ++r)
      {
        double a = 0.0, b = 0.0;
        for (int I = 0; I < 6; ++I) a += B[I + 6 * r] * ccg[I], b += B[I + 6 * r] * S[I];
        bopccg[r] = (fac * fbar / 3.0) * a;
        bops[r] = (-fac / fbar / 3.0) * b;
      }
      for (int r = 0; r < nd; ++r)
        for (int c = 0; c < nd; ++c) Ke[r + nd * c] += Hop[c] * (bops[r] + bopccg[r]);
    }
    return 0;
  }

  // ---- SolidEleCalc<celltype, MulfFormulation>  (src/solid_ele/4C_solid_ele_calc_mulf.hpp:36-148,
  //      src/solid_ele/4C_solid_ele_calc_lib_mulf.hpp:26-92)
  // Gauss-point history (MulfHistoryData), ORC_MULF_HIST doubles per point:
  //   [0..8] inverse_jacobian (a + 3 b), [9..17] deformation_gradient (a + 3 b), [18] is_setup
  // A fresh history is {identity, identity, 0}; orc_mulf_init_history writes it.
  void mulf_kinematics(int n, const double* dN, const double* u, const JacMap& jm, double* h, double* Fdelta, double* F)
  {
    if (h[18] == 0.0)  // calc_mulf.hpp:56-60
    {
      for (int i = 0; i < 9; ++i) h[i] = jm.iJ[i];
      h[18] = 1.0;
    }
    // evaluate_mulf_deformation_gradient_update (lib_mulf.hpp:41-60)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Fdelta[a + 3 * b] = (a == b) ? 1.0 : 0.0;
    for (int k = 0; k < n; ++k)
    {
      double Nh[3];
      for (int a = 0; a < 3; ++a)
      {
        double v = 0.0;
        for (int b = 0; b < 3; ++b) v += h[a + 3 * b] * dN[b + 3 * k];
        Nh[a] = v;
      }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Fdelta[a + 3 * b] += u[3 * k + a] * Nh[b];
    }
    // F = F_delta * F_history (lib_mulf.hpp:76-77)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int c = 0; c < 3; ++c) v += Fdelta[a + 3 * c] * h[9 + c + 3 * b];
        F[a + 3 * b] = v;
      }
  }

  int element_mulf(int ct, int matkind, const double* par, int action, const double* X,
      const double* u, double* hist, double* Ke, double* fe)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);
    std::memset(fe, 0, sizeof(double) * nd);
    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;
    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      double Fd[9], F[9], C[6], E[6], S[6], Cm[36];
      mulf_kinematics(n, dN, u, jm, hist + ORC_MULF_HIST * gp, Fd, F);
      cauchy_green(F, C, E);
      material(matkind, par, F, E, S, Cm);
      // add_internal_force_vector(Bop, stress) / add_elastic_stiffness_matrix(Bop, stress) +
      // add_geometric_stiffness_matrix with the B-operator of the UNMODIFIED reference gradients and
      // the MULF deformation gradient (calc_mulf.hpp:74-79,84-110); the B-free forms are the same
      // operator (unit test 4C_solid_ele_calc_lib_test.cpp:243-343)
      add_force(n, jm, F, S, fac, fe);
      if (action == ORC_ACTION_FORCE_STIFF && Ke) add_stiffness_bop(n, jm, F, S, Cm, fac, Ke);
    }
    return 0;
  }

  // SolidEleCalc::update_prestress (4C_solid_ele_calc.cpp:478-519) with
  // MulfFormulation::update_prestress (calc_mulf.hpp:124-143)
  // with_centroid: MulfFBarFormulation::update_prestress (mulf_fbar.hpp:245-277) -- the same history update at every
  // Gauss point and once more for the element's global history at the centroid (slot ngauss)
  void element_update_prestress(int ct, const double* X, const double* u, double* hist, bool with_centroid = false)
  {
    const int n = nnodes(ct), ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN], gxi[3 * MAXN + 3], gw[MAXN];
    JacMap jm;
    gauss(ct, gxi, gw);
    centroid_xi(ct, gxi + 3 * ng);
    for (int gp = 0; gp < ng + (with_centroid ? 1 : 0); ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      double* h = hist + ORC_MULF_HIST * gp;
      double Fd[9], F[9], iFd[9], niJ[9];
      mulf_kinematics(n, dN, u, jm, h, Fd, F);  // evaluate(): deformation_gradient of this state
      inv3(iFd, Fd);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int c = 0; c < 3; ++c) v += iFd[c + 3 * a] * h[c + 3 * b];
          niJ[a + 3 * b] = v;
        }
      for (int i = 0; i < 9; ++i) h[9 + i] = F[i], h[i] = niJ[i];
    }
  }


  // ---- SolidEleCalc<hex8, EASFormulation<eastype_h8_9 | eastype_h8_21>>
  //      (src/solid_ele/4C_solid_ele_calc_eas.hpp:44-290, 4C_solid_ele_calc_eas_helpers.hpp,
  //       driver 4C_solid_ele_calc.cpp:137-314)
  // Element history (EASHistoryData / EasIterationData, in the order pack() writes it, eas.hpp:266-275),
  // orc_eas_hist_size(neas) doubles: alpha_inc[neas], alpha[neas], s[neas], invKaa[neas x neas],
  // Kda[24 x neas] (both column-major).  A fresh history is all zero.
  // EASFormulation<hex8, eastype_h8_9 | eastype_h8_21, nonlinearTotLag | linear>: the four EAS evaluators of the factory
  // (4C_solid_ele_factory.hpp:62-76)
  bool is_eas(int tech)
  {
    return tech == ORC_TECH_EAS_MILD || tech == ORC_TECH_EAS_FULL || tech == ORC_TECH_EAS_MILD_LINEAR ||
           tech == ORC_TECH_EAS_FULL_LINEAR || tech == ORC_TECH_SHELL_EAS || tech == ORC_TECH_SHELL_EAS_ANS;
  }
  bool is_eas_linear(int tech) { return tech == ORC_TECH_EAS_MILD_LINEAR || tech == ORC_TECH_EAS_FULL_LINEAR; }
  int eas_num(int tech)
  {
    if (tech == ORC_TECH_SHELL_EAS || tech == ORC_TECH_SHELL_EAS_ANS) return 7;  // eastype_sh8_7
    return (tech == ORC_TECH_EAS_MILD || tech == ORC_TECH_EAS_MILD_LINEAR) ? 9 : 21;
  }

  // dense inverse by Gauss-Jordan with partial pivoting (the reference calls LAPACK getrf/getri through
  // FixedSizeSerialDenseSolver::invert; same operator up to round-off)
  void dense_inverse(int n, double* A /*column-major, in place*/)
  {
    std::vector<double> M(n * 2 * n, 0.0);  // [A | I], row-major
    for (int r = 0; r < n; ++r)
    {
      for (int c = 0; c < n; ++c) M[r * 2 * n + c] = A[r + n * c];
      M[r * 2 * n + n + r] = 1.0;
    }
    for (int c = 0; c < n; ++c)
    {
      int piv = c;
      for (int r = c + 1; r < n; ++r)
        if (std::fabs(M[r * 2 * n + c]) > std::fabs(M[piv * 2 * n + c])) piv = r;
      if (piv != c)
        for (int k = 0; k < 2 * n; ++k) std::swap(M[c * 2 * n + k], M[piv * 2 * n + k]);
      const double d = 1.0 / M[c * 2 * n + c];
      for (int k = 0; k < 2 * n; ++k) M[c * 2 * n + k] *= d;
      for (int r = 0; r < n; ++r)
      {
        if (r == c) continue;
        const double f = M[r * 2 * n + c];
        if (f != 0.0)
          for (int k = 0; k < 2 * n; ++k) M[r * 2 * n + k] -= f * M[c * 2 * n + k];
      }
    }
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) A[r + n * c] = M[r * 2 * n + n + c];
  }

  // evaluate_voigt_transformation_matrix at the centroid (4C_solid_ele_calc_lib.hpp:1451-1518):
  // T0invT (6x6 column-major, strain-like Voigt), jacobian_(a,b) = J[a + 3 b]
  void eas_centroid_transformation(int ct, const double* X, double* T0invT, double& detJ0)
  {
    const int n = nnodes(ct);
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    const double xi0[3] = {0.0, 0.0, 0.0};
    shape(ct, xi0, N, dN);
    jacobian_mapping(n, dN, X, jm);
    detJ0 = jm.det;
    auto J = [&](int a, int b) { return jm.J[a + 3 * b]; };
    double* T = T0invT;
    auto at = [&](int r, int c) -> double& { return T[r + 6 * c]; };
    for (int c = 0; c < 3; ++c)
    {
      at(0, c) = J(0, c) * J(0, c);
      at(1, c) = J(1, c) * J(1, c);
      at(2, c) = J(2, c) * J(2, c);
      at(3, c) = 2 * J(0, c) * J(1, c);
      at(4, c) = 2 * J(1, c) * J(2, c);
      at(5, c) = 2 * J(0, c) * J(2, c);
    }
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 2};  // columns 3,4,5 pair the jacobian columns (0,1),(1,2),(0,2)
    for (int q = 0; q < 3; ++q)
    {
      const int a = pa[q], b = pb[q], c = 3 + q;
      at(0, c) = J(0, a) * J(0, b);
      at(1, c) = J(1, a) * J(1, b);
      at(2, c) = J(2, a) * J(2, b);
      at(3, c) = J(0, a) * J(1, b) + J(1, a) * J(0, b);
      at(4, c) = J(1, a) * J(2, b) + J(2, a) * J(1, b);
      at(5, c) = J(0, a) * J(2, b) + J(2, a) * J(0, b);
    }
    dense_inverse(6, T);
  }


  // evaluate_voigt_transformation_matrix (src/solid_ele/4C_solid_ele_calc_lib.hpp:1451-1517): T^-T of the Jacobian jm.J
  // for strain-like Voigt vectors [xx,yy,zz,xy,yz,xz], column-major 6 x 6 (the matrix eas_centroid_transformation builds
  // at the centroid)
  void voigt_transformation(const double* Jm, double* TinvT)
  {
    auto J = [&](int a, int b) { return Jm[a + 3 * b]; };
    auto at = [&](int r, int c) -> double& { return TinvT[r + 6 * c]; };
    for (int c = 0; c < 3; ++c)
    {
      at(0, c) = J(0, c) * J(0, c);
      at(1, c) = J(1, c) * J(1, c);
      at(2, c) = J(2, c) * J(2, c);
      at(3, c) = 2 * J(0, c) * J(1, c);
      at(4, c) = 2 * J(1, c) * J(2, c);
      at(5, c) = 2 * J(0, c) * J(2, c);
    }
    const int pa[3] = {0, 1, 0}, pb[3] = {1, 2, 2};
    for (int q = 0; q < 3; ++q)
    {
      const int a = pa[q], b = pb[q], c = 3 + q;
      at(0, c) = J(0, a) * J(0, b);
      at(1, c) = J(1, a) * J(1, b);
      at(2, c) = J(2, a) * J(2, b);
      at(3, c) = J(0, a) * J(1, b) + J(1, a) * J(0, b);
      at(4, c) = J(1, a) * J(2, b) + J(2, a) * J(1, b);
      at(5, c) = J(0, a) * J(2, b) + J(2, a) * J(0, b);
    }
    dense_inverse(6, TinvT);
  }

  // ---- SolidEleCalc<hex8 | wedge6, ShellANSFormulation>  (TECH shell_ans: src/solid_ele/4C_solid_ele_calc_shell_ans.hpp:52-196,
  //      4C_solid_ele_calc_shell_ans_lib.hpp).  The thickness direction is the t-direction of the parameter space.  In
  //      the parametric strain components the transverse normal strain tt and the transverse shear strains st, rt are
  //      not taken at the Gauss point but interpolated from sampling points (hex8: A..D mid-edges of the mid-surface for
  //      the shears, E..H its corners for tt; wedge6: five points), for the strains, the B-operator and the geometric
  //      stiffness alike; T^-T of the Jacobian at the Gauss point takes everything to the global frame.
  // ANS Green-Lagrange strain at xi in the global frame, stress-like Voigt (evaluate_local_glstrain,
  // shell_ans_lib.hpp:169-291 hex8 / :294-380 wedge6, then T^-T and Voigt::Strains::to_stress_like as in
  // ShellANSFormulation::evaluate, shell_ans.hpp:96-110): rr, ss, rs at the point itself, tt / st / rt interpolated from
  // the sampling points of the mid-surface.  Used by the element routine below and by the output actions.
  void ans_gl_strain(int ct, const double* X, const double* u, const double* xi, const JacMap& jm, double* E)
  {
    const int n = nnodes(ct);
    const bool hex = ct == ORC_HEX8;
    const double sp_hex[8][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
    const double sp_wedge[5][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    double N[MAXN], dN[3 * MAXN];
    auto jac = [&](const double* dNq, bool current, double* J)
    {
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 3; ++d)
        {
          double v = 0.0;
          for (int k = 0; k < n; ++k) v += dNq[a + 3 * k] * (X[3 * k + d] + (current ? u[3 * k + d] : 0.0));
          J[a + 3 * d] = v;
        }
    };
    auto row_dot = [](const double* J, int a, int b) { return J[a] * J[b] + J[a + 3] * J[b + 3] + J[a + 6] * J[b + 6]; };
    // difference of the metric components (a, b) between the current and the reference configuration at sampling point q
    auto sampled = [&](int q, int a, int b)
    {
      double Jr[9], Jc[9];
      shape(ct, hex ? sp_hex[q] : sp_wedge[q], N, dN);
      jac(dN, false, Jr), jac(dN, true, Jc);
      return row_dot(Jc, a, b) - row_dot(Jr, a, b);
    };
    double tt, st, rt;
    if (hex)
    {
      tt = 0.25 * (1 - xi[0]) * (1 - xi[1]) * sampled(4, 2, 2) + 0.25 * (1 + xi[0]) * (1 - xi[1]) * sampled(5, 2, 2) +
           0.25 * (1 + xi[0]) * (1 + xi[1]) * sampled(6, 2, 2) + 0.25 * (1 - xi[0]) * (1 + xi[1]) * sampled(7, 2, 2);
      st = 0.5 * (1 + xi[0]) * sampled(1, 1, 2) + 0.5 * (1 - xi[0]) * sampled(3, 1, 2);
      rt = 0.5 * (1 - xi[1]) * sampled(0, 0, 2) + 0.5 * (1 + xi[1]) * sampled(2, 0, 2);
    }
    else
    {
      tt = (1 - xi[0] - xi[1]) * sampled(2, 2, 2) + xi[0] * sampled(3, 2, 2) + xi[1] * sampled(4, 2, 2);
      st = xi[0] * sampled(1, 1, 2);
      rt = xi[1] * sampled(0, 0, 2);
    }
    double Jc[9], El[6], TinvT[36];
    shape(ct, xi, N, dN);
    jac(dN, true, Jc);
    El[0] = 0.5 * (row_dot(Jc, 0, 0) - row_dot(jm.J, 0, 0));
    El[1] = 0.5 * (row_dot(Jc, 1, 1) - row_dot(jm.J, 1, 1));
    El[3] = row_dot(Jc, 0, 1) - row_dot(jm.J, 0, 1);
    El[2] = 0.5 * tt, El[4] = st, El[5] = rt;
    voigt_transformation(jm.J, TinvT);
    for (int r = 0; r < 6; ++r)
    {
      double v = 0.0;
      for (int q = 0; q < 6; ++q) v += TinvT[r + 6 * q] * El[q];
      E[r] = r < 3 ? v : 0.5 * v;  // Voigt::Strains::to_stress_like
    }
  }

  void eas_shape_functions(int neas, const double* xi, double* M);
  // hist != nullptr: TECH shell_eas_ans = SolidEleCalc<hex8, ShellEASANSFormulation<hex8, eastype_sh8_7>>
  // (src/solid_ele/4C_solid_ele_calc_shell_eas_ans.hpp:70-343): the ANS strains / B-operator / geometric stiffness below, the
  // strain enhanced by Mtilde alpha (:126-131), and the EAS condensation with the ANS B-operator (integrate_eas with
  // linearization.b_op, :217-225; prepare_condensed_contribution :227-241; the two condensed contributions :300-320).
  // hist layout as element_eas: alpha_inc | alpha | s | invKaa | Kda.  wedge6 (eastype_sw6_1) is not restated.
  int element_shell_ans(int ct, int matkind, const double* par, int action, const double* X, const double* u, double* Ke,
      double* fe, double* hist = nullptr)
  {
    if (ct != ORC_HEX8 && ct != ORC_WEDGE6) return 2;
    if (hist && ct != ORC_HEX8) return 2;
    const int neas = 7;
    double *alpha = nullptr, *s_eas = nullptr, *invKaa = nullptr, *Kda = nullptr, T0invT[36], detJ0 = 0.0;
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    if (hist)
    {
      alpha = hist + neas, s_eas = hist + 2 * neas, invKaa = hist + 3 * neas, Kda = invKaa + neas * neas;
      eas_centroid_transformation(ct, X, T0invT, detJ0);
      // reset_condensed_variable_integration (:208-215)
      std::fill(invKaa, invKaa + neas * neas, 0.0);
      std::fill(Kda, Kda + nd * neas, 0.0);
      std::fill(s_eas, s_eas + neas, 0.0);
    }
    const bool hex = ct == ORC_HEX8;
    const int nsp = hex ? 8 : 5;
    // SamplingPoints<celltype>::value (shell_ans_lib.hpp:26-40)
    const double sp_hex[8][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
    const double sp_wedge[5][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);
    std::memset(fe, 0, sizeof(double) * nd);
    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;
    // evaluate_sampling_point_data (shell_ans_lib.hpp:515-571): J(a,d) = sum_k dN(a,k) x(k,d), reference and current
    struct SP
    {
      double dN[3 * MAXN], Jr[9], Jc[9];
      std::vector<double> bop;  // 3 x nd: rows tt, st, rt
    };
    std::vector<SP> sp(nsp);
    auto jac = [&](const double* dNq, bool current, double* J)
    {
      for (int a = 0; a < 3; ++a)
        for (int d = 0; d < 3; ++d)
        {
          double v = 0.0;
          for (int k = 0; k < n; ++k) v += dNq[a + 3 * k] * (X[3 * k + d] + (current ? u[3 * k + d] : 0.0));
          J[a + 3 * d] = v;
        }
    };
    for (int q = 0; q < nsp; ++q)
    {
      shape(ct, hex ? sp_hex[q] : sp_wedge[q], N, sp[q].dN);
      jac(sp[q].dN, false, sp[q].Jr), jac(sp[q].dN, true, sp[q].Jc);
      sp[q].bop.assign(3 * nd, 0.0);
      for (int k = 0; k < n; ++k)
        for (int d = 0; d < 3; ++d)
        {
          const double* D = sp[q].dN + 3 * k;
          const double* Jc = sp[q].Jc;
          sp[q].bop[0 + 3 * (3 * k + d)] = D[2] * Jc[2 + 3 * d];
          sp[q].bop[1 + 3 * (3 * k + d)] = D[1] * Jc[2 + 3 * d] + D[2] * Jc[1 + 3 * d];
          sp[q].bop[2 + 3 * (3 * k + d)] = D[0] * Jc[2 + 3 * d] + D[2] * Jc[0 + 3 * d];
        }
    }
    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    std::vector<double> Bl(6 * nd), B(6 * nd), CB(6 * nd);
    for (int gp = 0; gp < ng; ++gp)
    {
      const double* xi = gxi + 3 * gp;
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      double Jc[9];
      jac(dN, true, Jc);
      // interpolation of the sampled rows: (sampling point, weight) lists of tt, st, rt
      int q_tt[4], q_st[2], q_rt[2], n_tt, n_st, n_rt;
      double w_tt[4], w_st[2], w_rt[2];
      if (hex)
      {
        n_tt = 4, n_st = 2, n_rt = 2;
        q_tt[0] = 4, q_tt[1] = 5, q_tt[2] = 6, q_tt[3] = 7;
        w_tt[0] = 0.25 * (1 - xi[0]) * (1 - xi[1]), w_tt[1] = 0.25 * (1 + xi[0]) * (1 - xi[1]);
        w_tt[2] = 0.25 * (1 + xi[0]) * (1 + xi[1]), w_tt[3] = 0.25 * (1 - xi[0]) * (1 + xi[1]);
        q_st[0] = 1, q_st[1] = 3, w_st[0] = 0.5 * (1 + xi[0]), w_st[1] = 0.5 * (1 - xi[0]);
        q_rt[0] = 0, q_rt[1] = 2, w_rt[0] = 0.5 * (1 - xi[1]), w_rt[1] = 0.5 * (1 + xi[1]);
      }
      else
      {
        n_tt = 3, n_st = 1, n_rt = 1;
        q_tt[0] = 2, q_tt[1] = 3, q_tt[2] = 4, w_tt[0] = 1 - xi[0] - xi[1], w_tt[1] = xi[0], w_tt[2] = xi[1];
        q_st[0] = 1, w_st[0] = xi[0];
        q_rt[0] = 0, w_rt[0] = xi[1];
      }
      // evaluate_local_b_operator (:60-118 / :119-167)
      for (int k = 0; k < n; ++k)
        for (int d = 0; d < 3; ++d)
        {
          const int c = 3 * k + d;
          const double* D = dN + 3 * k;
          Bl[0 + 6 * c] = D[0] * Jc[0 + 3 * d];
          Bl[1 + 6 * c] = D[1] * Jc[1 + 3 * d];
          Bl[3 + 6 * c] = D[0] * Jc[1 + 3 * d] + D[1] * Jc[0 + 3 * d];
          double tt = 0.0, st = 0.0, rt = 0.0;
          for (int m = 0; m < n_tt; ++m) tt += w_tt[m] * sp[q_tt[m]].bop[0 + 3 * c];
          for (int m = 0; m < n_st; ++m) st += w_st[m] * sp[q_st[m]].bop[1 + 3 * c];
          for (int m = 0; m < n_rt; ++m) rt += w_rt[m] * sp[q_rt[m]].bop[2 + 3 * c];
          Bl[2 + 6 * c] = tt, Bl[4 + 6 * c] = st, Bl[5 + 6 * c] = rt;
        }
      double TinvT[36];
      voigt_transformation(jm.J, TinvT);
      for (int c = 0; c < nd; ++c)
        for (int r = 0; r < 6; ++r)
        {
          double v = 0.0;
          for (int q = 0; q < 6; ++q) v += TinvT[r + 6 * q] * Bl[q + 6 * c];
          B[r + 6 * c] = v;
        }
      // evaluate_local_glstrain (:169-291 / :294-380), to the global frame, stress-like: ans_gl_strain above
      double E[6], S[6], Cm[36], F[9];
      ans_gl_strain(ct, X, u, xi, jm, E);
      double Mt[6 * 7];
      if (hist)
      {
        // m_tilde = detJ0 / detJ T0^-T M(xi) (evaluate_eas_shape_functions_material_config, eas_helpers.hpp:438-453);
        // gl_strain_eas = gl_strain + stress-like(Mtilde alpha) (shell_eas_ans.hpp:129-131)
        double M[6 * 21];
        eas_shape_functions(neas, xi, M);
        for (int c = 0; c < neas; ++c)
          for (int r = 0; r < 6; ++r)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v += T0invT[r + 6 * k] * M[k + 6 * c];
            Mt[r + 6 * c] = (detJ0 / jm.det) * v;
          }
        for (int r = 0; r < 6; ++r)
        {
          double v = 0.0;
          for (int c = 0; c < neas; ++c) v += Mt[r + 6 * c] * alpha[c];
          E[r] += r < 3 ? v : 0.5 * v;
        }
      }
      // the materials of this path use the strain only (the reference also hands in a deformation gradient consistent
      // with it, compute_deformation_gradient_from_gl_strains)
      deformation_gradient(n, u, jm, F);
      material(matkind, par, F, E, S, Cm);
      // add_internal_force_vector(Bop, stress) (calc_lib.hpp)
      for (int c = 0; c < nd; ++c)
      {
        double v = 0.0;
        for (int r = 0; r < 6; ++r) v += B[r + 6 * c] * S[r];
        fe[c] += fac * v;
      }
      if (hist)
      {
        // integrate_eas (eas_helpers.hpp:352-371) with the ANS B-operator
        double cmatM[6 * 7];
        for (int c = 0; c < neas; ++c)
          for (int r = 0; r < 6; ++r)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v += Cm[r + 6 * k] * Mt[k + 6 * c];
            cmatM[r + 6 * c] = v;
          }
        for (int c = 0; c < neas; ++c)
        {
          for (int r = 0; r < neas; ++r)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v += Mt[k + 6 * r] * cmatM[k + 6 * c];
            invKaa[r + neas * c] += fac * v;
          }
          for (int r = 0; r < nd; ++r)
          {
            double v = 0.0;
            for (int k = 0; k < 6; ++k) v += B[k + 6 * r] * cmatM[k + 6 * c];
            Kda[r + nd * c] += fac * v;
          }
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Mt[k + 6 * c] * S[k];
          s_eas[c] += fac * v;
        }
      }
      if (!(action == ORC_ACTION_FORCE_STIFF && Ke)) continue;
      for (int c = 0; c < nd; ++c)
        for (int I = 0; I < 6; ++I)
        {
          double v = 0.0;
          for (int J = 0; J < 6; ++J) v += Cm[I + 6 * J] * B[J + 6 * c];
          CB[I + 6 * c] = v;
        }
      for (int r = 0; r < nd; ++r)
        for (int c = 0; c < nd; ++c)
        {
          double v = 0.0;
          for (int I = 0; I < 6; ++I) v += B[I + 6 * r] * CB[I + 6 * c];
          Ke[r + nd * c] += fac * v;
        }
      // add_ans_geometric_stiffness (:383-454 / :456-512)
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
        {
          const double* Di = dN + 3 * i;
          const double* Dj = dN + 3 * j;
          double G[6];
          G[0] = Di[0] * Dj[0];
          G[1] = Di[1] * Dj[1];
          G[3] = Di[0] * Dj[1] + Di[1] * Dj[0];
          double gtt = 0.0, gst = 0.0, grt = 0.0;
          for (int m = 0; m < n_tt; ++m) gtt += w_tt[m] * sp[q_tt[m]].dN[2 + 3 * i] * sp[q_tt[m]].dN[2 + 3 * j];
          for (int m = 0; m < n_st; ++m)
            gst += w_st[m] * (sp[q_st[m]].dN[1 + 3 * i] * sp[q_st[m]].dN[2 + 3 * j] + sp[q_st[m]].dN[2 + 3 * i] * sp[q_st[m]].dN[1 + 3 * j]);
          for (int m = 0; m < n_rt; ++m)
            grt += w_rt[m] * (sp[q_rt[m]].dN[0 + 3 * i] * sp[q_rt[m]].dN[2 + 3 * j] + sp[q_rt[m]].dN[2 + 3 * i] * sp[q_rt[m]].dN[0 + 3 * j]);
          G[2] = gtt, G[4] = gst, G[5] = grt;
          double Gij = 0.0;
          for (int r = 0; r < 6; ++r)
          {
            double v = 0.0;
            for (int q = 0; q < 6; ++q) v += TinvT[r + 6 * q] * G[q];
            Gij += S[r] * v;
          }
          for (int d = 0; d < 3; ++d) Ke[(3 * i + d) + nd * (3 * j + d)] += fac * Gij;
        }
    }
    if (hist)
    {
      dense_inverse(neas, invKaa);
      std::vector<double> mKdaiKaa(nd * neas);
      for (int r = 0; r < nd; ++r)
        for (int c = 0; c < neas; ++c)
        {
          double v = 0.0;
          for (int k = 0; k < neas; ++k) v += Kda[r + nd * k] * invKaa[k + neas * c];
          mKdaiKaa[r + nd * c] = -v;
        }
      for (int r = 0; r < nd; ++r)
      {
        double v = 0.0;
        for (int c = 0; c < neas; ++c) v += mKdaiKaa[r + nd * c] * s_eas[c];
        fe[r] += v;
      }
      if (action == ORC_ACTION_FORCE_STIFF && Ke)
        for (int r = 0; r < nd; ++r)
          for (int c = 0; c < nd; ++c)
          {
            double v = 0.0;
            for (int k = 0; k < neas; ++k) v += mKdaiKaa[r + nd * k] * Kda[c + nd * k];
            Ke[r + nd * c] += v;
          }
    }
    return 0;
  }

  // evaluate_eas_shape_functions_parameter_space (eas_helpers.hpp:205-275): M 6 x neas column-major
  void eas_shape_functions(int neas, const double* xi, double* M)
  {
    std::fill(M, M + 6 * neas, 0.0);
    auto at = [&](int r, int c) -> double& { return M[r + 6 * c]; };
    if (neas == 7)
    {
      // eastype_sh8_7 (:330-357): the solid-shell interpolation with t = thickness direction
      at(0, 0) = xi[0], at(1, 1) = xi[1], at(2, 2) = xi[2], at(2, 5) = xi[0] * xi[2], at(2, 6) = xi[1] * xi[2];
      at(3, 3) = xi[0], at(3, 4) = xi[1];
      return;
    }
    at(0, 0) = xi[0], at(1, 1) = xi[1], at(2, 2) = xi[2];
    at(3, 3) = xi[0], at(3, 4) = xi[1], at(4, 5) = xi[1], at(4, 6) = xi[2], at(5, 7) = xi[0], at(5, 8) = xi[2];
    if (neas == 21)
    {
      at(0, 15) = xi[0] * xi[1], at(0, 16) = xi[0] * xi[2];
      at(1, 17) = xi[0] * xi[1], at(1, 18) = xi[1] * xi[2];
      at(2, 19) = xi[0] * xi[2], at(2, 20) = xi[1] * xi[2];
      at(3, 9) = xi[0] * xi[2], at(3, 10) = xi[1] * xi[2];
      at(4, 11) = xi[0] * xi[1], at(4, 12) = xi[0] * xi[2];
      at(5, 13) = xi[0] * xi[1], at(5, 14) = xi[1] * xi[2];
    }
  }

  // Gauss-point kinematics of the EAS formulation (evaluate_eas_kinematics, eas_helpers.hpp:405-448):
  // Mtilde = detJ0/detJ T0invT M, E_enh = E(F) + stress-like(Mtilde alpha)
  void eas_point(int neas, const double* T0invT, double detJ0, double detJ, const double* xi,
      const double* alpha, const double* F, double* Mtilde, double* E_enh)
  {
    double M[6 * 21];
    eas_shape_functions(neas, xi, M);
    for (int c = 0; c < neas; ++c)
      for (int r = 0; r < 6; ++r)
      {
        double v = 0.0;
        for (int k = 0; k < 6; ++k) v += T0invT[r + 6 * k] * M[k + 6 * c];
        Mtilde[r + 6 * c] = (detJ0 / detJ) * v;
      }
    double C[6], E[6];
    cauchy_green(F, C, E);
    for (int r = 0; r < 6; ++r)
    {
      double v = 0.0;
      for (int c = 0; c < neas; ++c) v += Mtilde[r + 6 * c] * alpha[c];
      E_enh[r] = E[r] + (r < 3 ? v : 0.5 * v);  // Voigt::Strains::to_stress_like
    }
  }

  int element_eas(int ct, int tech, int matkind, const double* par, int action, const double* X,
      const double* u, double* hist, double* Ke, double* fe)
  {
    if (ct != ORC_HEX8) return 2;
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct), neas = eas_num(tech);
    double* alpha = hist + neas;
    double* s = hist + 2 * neas;
    double* invKaa = hist + 3 * neas;
    double* Kda = invKaa + neas * neas;
    double N[MAXN], dN[3 * MAXN];
    JacMap jm;
    if (Ke) std::memset(Ke, 0, sizeof(double) * nd * nd);
    std::memset(fe, 0, sizeof(double) * nd);
    // prepare(): centroid transformation (eas.hpp:66-70)
    double T0invT[36], detJ0;
    eas_centroid_transformation(ct, X, T0invT, detJ0);
    // reset_condensed_variable_integration (eas.hpp:168-175)
    std::fill(invKaa, invKaa + neas * neas, 0.0);
    std::fill(Kda, Kda + nd * neas, 0.0);
    std::fill(s, s + neas, 0.0);
    int err = 0;
    for (int k = 0; k < n; ++k)
    {
      double xi[3];
      node_ref(ct, k, xi);
      shape(ct, xi, N, dN);
      jacobian_mapping(n, dN, X, jm);
      if (!(jm.det > 0)) err = 1;
    }
    if (err) return 1;
    double gxi[3 * MAXN], gw[MAXN];
    gauss(ct, gxi, gw);
    std::vector<double> B(6 * nd), cmatM(6 * neas), Mt(6 * neas);
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      double F[9], E[6], S[6], Cm[36];
      deformation_gradient(n, u, jm, F);
      double Elin[6] = {0, 0, 0, 0, 0, 0};
      if (is_eas_linear(tech))
      {
        // evaluate_eas_kinematics<..., KinemType::linear> (eas_helpers.hpp:648-666): b_op = linear strain gradient,
        // enhanced strain = B_lin u + Mtilde alpha, deformation gradient = identity.  add_stiffness_matrix below keeps
        // its geometric term, as the reference's EASFormulation does for both kinematic types (eas.hpp:142-157)
        for (int v = 0; v < 6; ++v)
          Elin[v] = 0.5 * ((F[VA[v] + 3 * VB[v]] - (v < 3 ? 1.0 : 0.0)) + (F[VB[v] + 3 * VA[v]] - (v < 3 ? 1.0 : 0.0)));
        for (int i = 0; i < 9; ++i) F[i] = (i % 4 == 0) ? 1.0 : 0.0;
      }
      strain_gradient(n, jm, F, B.data());
      eas_point(neas, T0invT, detJ0, jm.det, gxi + 3 * gp, alpha, F, Mt.data(), E);
      for (int v = 0; v < 6; ++v) E[v] += Elin[v];
      // the materials on this path use the GL strain only; the enhanced deformation gradient
      // (compute_deformation_gradient_from_gl_strains) is not needed by them
      g_material_hist = (g_element_material_hist && finite_strain_hist(matkind))
                            ? g_element_material_hist + finite_strain_hist(matkind) * gp
                            : nullptr;
      material(matkind, par, F, E, S, Cm);
      g_material_hist = nullptr;
      // integrate_eas (eas_helpers.hpp:352-371)
      for (int c = 0; c < neas; ++c)
        for (int r = 0; r < 6; ++r)
        {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Cm[r + 6 * k] * Mt[k + 6 * c];
          cmatM[r + 6 * c] = v;
        }
      for (int c = 0; c < neas; ++c)
      {
        for (int r = 0; r < neas; ++r)
        {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Mt[k + 6 * r] * cmatM[k + 6 * c];
          invKaa[r + neas * c] += fac * v;
        }
        for (int r = 0; r < nd; ++r)
        {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += B[k + 6 * r] * cmatM[k + 6 * c];
          Kda[r + nd * c] += fac * v;
        }
        double v = 0.0;
        for (int k = 0; k < 6; ++k) v += Mt[k + 6 * c] * S[k];
        s[c] += fac * v;
      }
      add_force(n, jm, F, S, fac, fe);
      if (action == ORC_ACTION_FORCE_STIFF && Ke) add_stiffness_bop(n, jm, F, S, Cm, fac, Ke);
    }
    // prepare_condensed_contribution (eas.hpp:188-203) and the two condensed contributions (:245-263)
    dense_inverse(neas, invKaa);
    std::vector<double> mKdaiKaa(nd * neas);
    for (int r = 0; r < nd; ++r)
      for (int c = 0; c < neas; ++c)
      {
        double v = 0.0;
        for (int k = 0; k < neas; ++k) v += Kda[r + nd * k] * invKaa[k + neas * c];
        mKdaiKaa[r + nd * c] = -v;
      }
    for (int r = 0; r < nd; ++r)
    {
      double v = 0.0;
      for (int c = 0; c < neas; ++c) v += mKdaiKaa[r + nd * c] * s[c];
      fe[r] += v;
    }
    if (action == ORC_ACTION_FORCE_STIFF && Ke)
      for (int r = 0; r < nd; ++r)
        for (int c = 0; c < nd; ++c)
        {
          double v = 0.0;
          for (int k = 0; k < neas; ++k) v += mKdaiKaa[r + nd * k] * Kda[c + nd * k];
          Ke[r + nd * c] += v;
        }
    return 0;
  }

  // SolidEleCalc::recover, default step (4C_solid_ele_calc.cpp:284-314; update_alpha,
  // eas_helpers.hpp:166-190): alpha_inc = -invKaa (s + Kda^T du), alpha += step_length alpha_inc
  void element_eas_recover(int tech, const double* du, double step_length, double* hist)
  {
    const int neas = eas_num(tech), nd = 24;
    double* alpha_inc = hist;
    double* alpha = hist + neas;
    const double* s = hist + 2 * neas;
    const double* invKaa = hist + 3 * neas;
    const double* Kda = invKaa + neas * neas;
    std::vector<double> tmp(s, s + neas);
    for (int c = 0; c < neas; ++c)
      for (int r = 0; r < nd; ++r) tmp[c] += Kda[r + nd * c] * du[r];
    for (int r = 0; r < neas; ++r)
    {
      double v = 0.0;
      for (int c = 0; c < neas; ++c) v += invKaa[r + neas * c] * tmp[c];
      alpha_inc[r] = -v;
    }
    for (int r = 0; r < neas; ++r) alpha[r] += step_length * alpha_inc[r];
  }

  int element_tech(int ct, int tech, int matkind, const double* par, int action, const double* X,
      const double* u, double* Ke, double* fe, double* hist = nullptr)
  {
    if (tech == ORC_TECH_SHELL_EAS_ANS) return element_shell_ans(ct, matkind, par, action, X, u, Ke, fe, hist);
    if (is_eas(tech))
      return element_eas(ct, tech, matkind, par, action, X, u, hist, Ke, fe);
    if (tech == ORC_TECH_LINEAR) return element_linear(ct, matkind, par, action, X, u, Ke, fe, hist);
    if (tech == ORC_TECH_MULF) return element_mulf(ct, matkind, par, action, X, u, hist, Ke, fe);
    if (tech == ORC_TECH_FBAR) return element_fbar(ct, matkind, par, action, X, u, Ke, fe);
    if (tech == ORC_TECH_MULF_FBAR) return element_fbar(ct, matkind, par, action, X, u, Ke, fe, hist);
    if (tech == ORC_TECH_SHELL_ANS) return element_shell_ans(ct, matkind, par, action, X, u, Ke, fe);
    return element(ct, matkind, par, action, X, u, Ke, fe);
  }
}  // namespace

extern "C"
{
  int orc_num_nodes(int ct) { return nnodes(ct); }
  int orc_num_gp(int ct) { return ngauss(ct); }
  void orc_shape(int ct, const double* xi, double* N, double* dN) { shape(ct, xi, N, dN); }
  void orc_gauss(int ct, double* xi, double* w) { gauss(ct, xi, w); }
  void orc_node_param_coords(int ct, double* xi)
  {
    for (int k = 0; k < nnodes(ct); ++k) node_ref(ct, k, xi + 3 * k);
  }
  // test hook: add_coefficients_stretches_modified of one summand (pins ELAST_IsoOgden against
  // unittests/mat/elast/4C_elast_isoogden_test.cpp:44-68)
  void orc_test_stretch_coefficients_modified(int kind, const double* p, const double* modstr, double* modgamma,
      double* moddelta)
  {
    for (int i = 0; i < 3; ++i) modgamma[i] = 0.0;
    for (int i = 0; i < 6; ++i) moddelta[i] = 0.0;
    stretch_coefficients_modified(kind, p, modstr, modgamma, moddelta);
  }

  // test hook: CoupAnisoExpoBase for a given structural tensor / scalar product
  // (pins the summand against unittests/mat/elast/4C_elast_coupanisoexpobase_test.cpp:146-230)
  void orc_test_aniso_expo(const double* k4, const double* A6, double scalar_product, const double* C6, double* psi,
      double* dPI, double* ddPII)
  {
    aniso_expo_point(k4, A6, scalar_product, C6, psi, dPI, ddPII);
  }
  void orc_material_hist(int kind, const double* par, const double* E, double* hist, double* S, double* C)
  {
    const double Id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    g_material_hist = hist;
    material(kind, par, Id, E, S, C);
    g_material_hist = nullptr;
  }
  void orc_material_update(int kind, int64_t n_points, double* hist)
  {
    if (kind == ORC_MAT_VISCONEOHOOKE)  // Mat::ViscoNeoHooke::update (:197-216): last <- current, current <- 0
      for (int64_t k = 0; k < n_points; ++k)
        for (int i = 0; i < 12; ++i)
        {
          hist[ORC_VISCO_HIST * k + i] = hist[ORC_VISCO_HIST * k + 12 + i];
          hist[ORC_VISCO_HIST * k + 12 + i] = 0.0;
        }
    if (kind == ORC_MAT_PLASTIC_NLNLOG)  // Mat::PlasticNlnLogNeoHooke::update (:359-378): last <- current, current <- {0, I}
      for (int64_t k = 0; k < n_points; ++k)
        for (int i = 0; i < 7; ++i)
        {
          hist[ORC_PLASTIC_HIST * k + i] = hist[ORC_PLASTIC_HIST * k + 7 + i];
          hist[ORC_PLASTIC_HIST * k + 7 + i] = (i >= 1 && i <= 3) ? 1.0 : 0.0;
        }
    if (kind != ORC_MAT_DRUCKER_PRAGER) return;
    for (int64_t k = 0; k < n_points; ++k)
      for (int i = 0; i < 7; ++i) hist[ORC_DP_HIST * k + i] = hist[ORC_DP_HIST * k + 7 + i];
  }
  void orc_set_material_history(double* mat_hist) { g_material_hist_base = mat_hist; }
  void orc_set_element_material_history(double* mat_hist) { g_element_material_hist = mat_hist; }
  void orc_material(
      int kind, const double* par, const double* F, const double* E, double* S, double* C)
  {
    material(kind, par, F, E, S, C);
  }
  void orc_calculate_gamma_delta(
      const double* prinv, const double* dPI, const double* ddPII, double* gamma, double* delta)
  {
    gamma_delta(prinv, dPI, ddPII, gamma, delta);
  }
  void orc_invariant_derivatives(
      int kind, const double* par, const double* prinv, double* dPI, double* ddPII)
  {
    invariant_derivatives(kind, par, prinv, dPI, ddPII);
  }
  // Mat::StVenantKirchhoff::strain_energy = 1/2 E:C:E (src/mat/4C_mat_stvenantkirchhoff.cpp:119-135)
  double orc_svk_strain_energy(const double* par, const double* E)
  {
    double S[6], Cm[36];
    svk(par, E, S, Cm);
    double psi = 0.0;
    for (int i = 0; i < 6; ++i) psi += 0.5 * S[i] * E[i] * (i < 3 ? 1.0 : 2.0);
    return psi;
  }
  int orc_element(int ct, int matkind, const double* par, int action, const double* X,
      const double* u, double* Ke, double* fe)
  {
    return element(ct, matkind, par, action, X, u, Ke, fe);
  }
  int orc_element_tech(int ct, int tech, int matkind, const double* par, int action,
      const double* X, const double* u, double* Ke, double* fe)
  {
    return element_tech(ct, tech, matkind, par, action, X, u, Ke, fe);
  }

  static void point_setup(
      int ct, const double* X, const double* u, const double* xi, JacMap& jm, double* F)
  {
    double N[MAXN], dN[3 * MAXN];
    shape(ct, xi, N, dN);
    jacobian_mapping(nnodes(ct), dN, X, jm);
    deformation_gradient(nnodes(ct), u, jm, F);
  }
  void orc_stiffness_bfree(int ct, const double* X, const double* u, const double* xi,
      const double* S, const double* C, double fac, double* Ke)
  {
    JacMap jm;
    double F[9];
    point_setup(ct, X, u, xi, jm, F);
    add_stiffness_bfree(nnodes(ct), jm, F, S, C, fac, Ke);
  }
  void orc_stiffness_bop(int ct, const double* X, const double* u, const double* xi,
      const double* S, const double* C, double fac, double* Ke)
  {
    JacMap jm;
    double F[9];
    point_setup(ct, X, u, xi, jm, F);
    add_stiffness_bop(nnodes(ct), jm, F, S, C, fac, Ke);
  }


  // ---- further element actions -------------------------------------------------
  namespace
  {
    // Mat::So3Material::strain_energy of the oracle materials
    double strain_energy(int kind, const double* par, const double* E)
    {
      if (kind == ORC_MAT_SVK) return orc_svk_strain_energy(par, E);
      double C[6], I[3];
      for (int i = 0; i < 6; ++i) C[i] = 2.0 * E[i] + (i < 3 ? 1.0 : 0.0);
      I[0] = C[0] + C[1] + C[2];
      I[1] = 0.5 * (I[0] * I[0] - C[0] * C[0] - C[1] * C[1] - C[2] * C[2]) - C[3] * C[3] - C[4] * C[4] -
             C[5] * C[5];
      I[2] = C[0] * C[1] * C[2] + 2 * (C[3] * C[4] * C[5]) - C[1] * C[5] * C[5] - C[2] * C[3] * C[3] -
             C[0] * C[4] * C[4];
      if (kind == ORC_MAT_ELASTHYPER)
      {
        double psi = hyper_generic(par, I, nullptr, nullptr);
        aniso_expo_add(par, C, nullptr, nullptr, &psi);
        return psi;
      }
      double blk[13];
      if (kind == ORC_MAT_COUP_NEOHOOKE)
      {
        const double b[7] = {1, ORC_SUM_COUP_NEOHOOKE, 0, par[0], par[1], 0, 0};
        std::copy(b, b + 7, blk);
      }
      else
      {
        const double b[13] = {2, ORC_SUM_ISO_NEOHOOKE, 0, par[0], 0, 0, 0, ORC_SUM_VOL_SUSSMANBATHE, 0,
            par[1], 0, 0, 0};
        std::copy(b, b + 13, blk);
      }
      return hyper_generic(blk, I, nullptr, nullptr);
    }

    // pure stretch tensor U = C^(1/2) (compute_pure_stretch_tensor, 4C_solid_ele_calc_lib.hpp:1577-1586)
    void pure_stretch(const double* Cv, double* U)
    {
      double Cf[9], w[3], V[9];
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) Cf[a + 3 * b] = Cv[VG[a][b]];
      sym_eig3(Cf, w, V);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int k = 0; k < 3; ++k) v += V[a + 3 * k] * std::sqrt(w[k]) * V[b + 3 * k];
          U[a + 3 * b] = v;
        }
    }
    // compute_deformation_gradient_from_gl_strains (calc_lib.hpp:1566-1595): F_enh = R U_enh, R = F U^-1
    void enhanced_deformation_gradient(const double* F, const double* E_enh, double* Fenh)
    {
      double C[6], Ce[6], U[9], iU[9], Ue[9], R[9];
      cauchy_green(F, C, nullptr);
      for (int v = 0; v < 6; ++v) Ce[v] = 2.0 * E_enh[v] + (v < 3 ? 1.0 : 0.0);
      pure_stretch(C, U);
      pure_stretch(Ce, Ue);
      inv3(iU, U);
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int k = 0; k < 3; ++k) v += F[a + 3 * k] * iU[k + 3 * b];
          R[a + 3 * b] = v;
        }
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
        {
          double v = 0.0;
          for (int k = 0; k < 3; ++k) v += R[a + 3 * k] * Ue[k + 3 * b];
          Fenh[a + 3 * b] = v;
        }
    }

    // element history of the EAS formulations for the output actions (set by orc_set_output_history)
    const double* g_eas_hist = nullptr;
    const double* g_xi = nullptr;  // parameter coordinates of the Gauss point being evaluated

    // kinematics of the formulation at one Gauss point: F and E as handed to the material
    // (DisplacementBasedFormulation::evaluate, 4C_solid_ele_calc_displacement_based.hpp:44-63;
    //  FBarFormulation::evaluate, 4C_solid_ele_calc_fbar.hpp:83-118)
    void formulation_kinematics(int ct, int tech, const double* X, const double* u, const JacMap& jm,
        double* F, double* E)
    {
      const int n = nnodes(ct);
      deformation_gradient(n, u, jm, F);
      if (tech == ORC_TECH_SHELL_ANS || tech == ORC_TECH_SHELL_EAS_ANS)
      {
        // ShellANSFormulation::evaluate (shell_ans.hpp:96-117) / ShellEASANSFormulation::evaluate (shell_eas_ans.hpp:
        // 116-139): the ANS strain, with EAS enhanced by Mtilde alpha; deformation gradient consistent with it
        double Fc[9];
        ans_gl_strain(ct, X, u, g_xi, jm, E);
        if (tech == ORC_TECH_SHELL_EAS_ANS)
        {
          const int neas = eas_num(tech);
          double T0invT[36], detJ0, M[6 * 21];
          eas_centroid_transformation(ct, X, T0invT, detJ0);
          eas_shape_functions(neas, g_xi, M);
          const double* alpha = g_eas_hist + neas;
          for (int r = 0; r < 6; ++r)
          {
            double v = 0.0;
            for (int c = 0; c < neas; ++c)
              for (int k = 0; k < 6; ++k) v += (detJ0 / jm.det) * T0invT[r + 6 * k] * M[k + 6 * c] * alpha[c];
            E[r] += r < 3 ? v : 0.5 * v;
          }
        }
        enhanced_deformation_gradient(F, E, Fc);
        for (int i = 0; i < 9; ++i) F[i] = Fc[i];
        return;
      }
      if (is_eas(tech))
      {
        // EASFormulation::evaluate (4C_solid_ele_calc_eas.hpp:72-88): enhanced GL strain and deformation gradient
        const int neas = eas_num(tech);
        double T0invT[36], detJ0, Mt[6 * 21], Fenh[9];
        eas_centroid_transformation(ct, X, T0invT, detJ0);
        if (is_eas_linear(tech))
        {
          // KINEM linear (eas_helpers.hpp:648-666): linear strain + Mtilde alpha, deformation gradient = identity
          double Elin[6];
          for (int v = 0; v < 6; ++v)
            Elin[v] = 0.5 * ((F[VA[v] + 3 * VB[v]] - (v < 3 ? 1.0 : 0.0)) + (F[VB[v] + 3 * VA[v]] - (v < 3 ? 1.0 : 0.0)));
          for (int i = 0; i < 9; ++i) F[i] = (i % 4 == 0) ? 1.0 : 0.0;
          eas_point(neas, T0invT, detJ0, jm.det, g_xi, g_eas_hist + neas, F, Mt, E);
          for (int v = 0; v < 6; ++v) E[v] += Elin[v];
          return;
        }
        eas_point(neas, T0invT, detJ0, jm.det, g_xi, g_eas_hist + neas, F, Mt, E);
        enhanced_deformation_gradient(F, E, Fenh);
        for (int i = 0; i < 9; ++i) F[i] = Fenh[i];
        return;
      }
      if (tech == ORC_TECH_LINEAR)
      {
        for (int v = 0; v < 6; ++v)
          E[v] = 0.5 * ((F[VA[v] + 3 * VB[v]] - (v < 3 ? 1.0 : 0.0)) + (F[VB[v] + 3 * VA[v]] - (v < 3 ? 1.0 : 0.0)));
        for (int i = 0; i < 9; ++i) F[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
      }
      if (tech == ORC_TECH_FBAR)
      {
        double N[MAXN], dN[3 * MAXN], F0[9];
        JacMap jm0;
        double xi0[3];
        centroid_xi(ct, xi0);
        shape(ct, xi0, N, dN);
        jacobian_mapping(n, dN, X, jm0);
        deformation_gradient(n, u, jm0, F0);
        const double fbar = std::pow(det3(F0) / det3(F), 1.0 / 3.0);
        for (int i = 0; i < 9; ++i) F[i] *= fbar;
      }
      double C[6];
      cauchy_green(F, C, E);
    }
  }  // namespace

  // history of all elements for the output actions of history-carrying formulations (EAS): set before
  // orc_calc_stress / orc_calc_energy, NULL afterwards
  static const double* g_output_hist = nullptr;
  void orc_set_output_history(const double* hist) { g_output_hist = hist; }

  // Solid::Utils::pk2_to_cauchy (src/solid_ele/4C_solid_ele_utils.hpp:41-54): sigma = F S F^T / det F
  void stress_pk2_to_cauchy(const double* F, const double* S, double* out)
  {
    double FS[9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {
        double v = 0.0;
        for (int c = 0; c < 3; ++c) v += F[a + 3 * c] * S[VG[c][b]];
        FS[a + 3 * b] = v;
      }
    const double det = det3(F);
    for (int v = 0; v < 6; ++v)
    {
      double w = 0.0;
      for (int c = 0; c < 3; ++c) w += FS[VA[v] + 3 * c] * F[VB[v] + 3 * c];
      out[v] = w / det;
    }
  }
  // Solid::Utils::green_lagrange_to_euler_almansi (:64-79): e = F^-T E F^-1
  void strain_gl_to_euler_almansi(const double* F, const double* E, double* out)
  {
    double iF[9], T[9];
    inv3(iF, F);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
      {  // T = iF^T E
        double v = 0.0;
        for (int c = 0; c < 3; ++c) v += iF[c + 3 * a] * E[VG[c][b]];
        T[a + 3 * b] = v;
      }
    for (int v = 0; v < 6; ++v)
    {
      double w = 0.0;
      for (int c = 0; c < 3; ++c) w += T[VA[v] + 3 * c] * iF[c + 3 * VB[v]];
      out[v] = w;
    }
  }
  // test hooks for unittests/solid_ele/4C_solid_ele_utils_test.cpp
  void orc_test_pk2_to_cauchy(const double* F, const double* S, double* out) { stress_pk2_to_cauchy(F, S, out); }
  void orc_test_gl_to_euler_almansi(const double* F, const double* E, double* out) { strain_gl_to_euler_almansi(F, E, out); }

  // SolidEleCalc::calculate_stress (src/solid_ele/4C_solid_ele_calc.cpp:416-476) for one element:
  // stress / strain [ngp][6] in the row layout of assemble_symmetric_tensor_to_matrix_row
  // (4C_solid_ele_calc_lib_io.hpp:41-70).  stress_type 1 = 2PK, 2 = Cauchy (pk2_to_cauchy,
  // 4C_solid_ele_utils.hpp:41-54); strain_type 1 = GL, 2 = EA (green_lagrange_to_euler_almansi :64-79)
  void orc_element_stress(int ct, int tech, int matkind, const double* par, const double* X,
      const double* u, int stress_type, int strain_type, double* stress, double* strain)
  {
    const int n = nnodes(ct), ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN], gxi[3 * MAXN], gw[MAXN];
    JacMap jm;
    gauss(ct, gxi, gw);
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      double F[9], E[6], S[6], Cm[36];
      g_xi = gxi + 3 * gp;
      formulation_kinematics(ct, tech, X, u, jm, F, E);
      material(matkind, par, F, E, S, Cm);
      if (stress && stress_type == 1)
        for (int v = 0; v < 6; ++v) stress[6 * gp + v] = S[v];
      if (stress && stress_type == 2)
      {
        stress_pk2_to_cauchy(F, S, stress + 6 * gp);
      }
      if (strain && strain_type == 1)
        for (int v = 0; v < 6; ++v) strain[6 * gp + v] = E[v];
      if (strain && strain_type == 2)
      {
        strain_gl_to_euler_almansi(F, E, strain + 6 * gp);
      }
    }
  }

  // SolidEleCalc::calculate_internal_energy (4C_solid_ele_calc.cpp:365-414)
  double orc_element_energy(int ct, int tech, int matkind, const double* par, const double* X,
      const double* u)
  {
    const int n = nnodes(ct), ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN], gxi[3 * MAXN], gw[MAXN];
    JacMap jm;
    gauss(ct, gxi, gw);
    double intenergy = 0.0;
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      double F[9], E[6];
      g_xi = gxi + 3 * gp;
      formulation_kinematics(ct, tech, X, u, jm, F, E);
      intenergy += strain_energy(matkind, par, E) * (jm.det * gw[gp]);
    }
    return intenergy;
  }
  double orc_strain_energy(int matkind, const double* par, const double* E)
  {
    return strain_energy(matkind, par, E);
  }

  // mass rule: DisTypeToOptGaussRule (4C_fem_general_element_integration_select.hpp:37-67):
  // hex_8point / hex_27point (= stiffness rule) and tet_11point
  // (4C_fem_general_utils_integration.cpp:447-495)
  static int mass_rule(int ct, double* xi, double* w)
  {
    if (ct == ORC_TET4)
    {
      // DisTypeToOptGaussRule<tet4> = tet_4point (:331-352): the points of tet10's stiffness rule
      gauss(ORC_TET10, xi, w);
      return 4;
    }
    if (ct != ORC_TET10)
    {
      gauss(ct, xi, w);
      return ngauss(ct);
    }
    const double Q16 = 1.0 / 6.0;
    const double q[11][3] = {{0.2500000000000000, 0.2500000000000000, 0.2500000000000000},
        {0.7857142857142857, 0.0714285714285714, 0.0714285714285714},
        {0.0714285714285714, 0.0714285714285714, 0.0714285714285714},
        {0.0714285714285714, 0.0714285714285714, 0.7857142857142857},
        {0.0714285714285714, 0.7857142857142857, 0.0714285714285714},
        {0.1005964238332008, 0.3994035761667992, 0.3994035761667992},
        {0.3994035761667992, 0.1005964238332008, 0.3994035761667992},
        {0.3994035761667992, 0.3994035761667992, 0.1005964238332008},
        {0.3994035761667992, 0.1005964238332008, 0.1005964238332008},
        {0.1005964238332008, 0.3994035761667992, 0.1005964238332008},
        {0.1005964238332008, 0.1005964238332008, 0.3994035761667992}};
    for (int i = 0; i < 11; ++i)
      for (int d = 0; d < 3; ++d) xi[3 * i + d] = q[i][d];
    w[0] = -0.0789333333333333 * Q16;
    for (int i = 1; i < 5; ++i) w[i] = 0.0457333333333333 * Q16;
    for (int i = 5; i < 11; ++i) w[i] = 0.1493333333333333 * Q16;
    return 11;
  }

  // Element mass matrix as evaluate_nonlinear_force_stiffness_mass integrates it
  // (4C_solid_ele_calc.cpp:151-153,226-236,265-280; add_mass_matrix calc_lib.hpp:1221-1238),
  // optional lump_matrix (calc_lib.hpp:73-89), inertia force M_e * acc
  // (4C_solid_ele_evaluate.cpp:36-59).  Me column-major 3n x 3n; finert may be NULL.
  void orc_element_mass(int ct, double density, const double* X, int lumped, const double* acc,
      double* Me, double* finert)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    double N[MAXN], dN[3 * MAXN], gxi[3 * MAXN], gw[MAXN], mxi[3 * MAXN], mw[MAXN];
    JacMap jm;
    std::memset(Me, 0, sizeof(double) * nd * nd);
    gauss(ct, gxi, gw);
    const int nm = mass_rule(ct, mxi, mw);
    const bool equal = (nm == ng);  // compare_gauss_integration: number of points
    auto add_mass = [&](double fac, double rho)
    {
      for (int inod = 0; inod < n; ++inod)
      {
        const double ifactor = N[inod] * fac * rho;
        for (int jnod = 0; jnod < n; ++jnod)
        {
          const double massfactor = N[jnod] * ifactor;
          for (int d = 0; d < 3; ++d) Me[(3 * inod + d) + nd * (3 * jnod + d)] += massfactor;
        }
      }
    };
    double element_mass = 0.0, element_volume = 0.0;
    for (int gp = 0; gp < ng; ++gp)
    {
      shape(ct, gxi + 3 * gp, N, dN);
      jacobian_mapping(n, dN, X, jm);
      const double fac = jm.det * gw[gp];
      if (equal)
        add_mass(fac, density);
      else
        element_mass += density * fac, element_volume += fac;
    }
    if (!equal)
      for (int gp = 0; gp < nm; ++gp)
      {
        shape(ct, mxi + 3 * gp, N, dN);
        jacobian_mapping(n, dN, X, jm);
        add_mass(jm.det * mw[gp], element_mass / element_volume);
      }
    if (lumped)
      for (int c = 0; c < nd; ++c)
      {
        double d = 0.0;
        for (int r = 0; r < nd; ++r)
        {
          d += Me[r + nd * c];
          Me[r + nd * c] = 0.0;
        }
        Me[c + nd * c] = d;
      }
    if (finert && acc)
      for (int r = 0; r < nd; ++r)
      {
        double v = 0.0;
        for (int c = 0; c < nd; ++c) v += Me[r + nd * c] * acc[c];
        finert[r] = v;
      }
  }

  // Whole-mesh versions: Gauss-point stresses of every column element, internal energy of the row
  // elements, mass matrix + inertia force assembled owner-computes into the stiffness graph.
  void orc_calc_stress(int ct, int tech, int matkind, const double* matparams, const double* mue_per_ele,
      int64_t n_ele, const int32_t* conn, const double* X, const double* u_col, int stress_type,
      int strain_type, double* stress, double* strain)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    std::vector<double> Xe(nd), ue(nd);
    std::vector<double> par(matparams, matparams + (matkind == ORC_MAT_ELASTHYPER ? 1 + 6 * (int)matparams[0] : 2));
    for (int64_t e = 0; e < n_ele; ++e)
    {
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
          Xe[3 * i + d] = X[3 * (int64_t)conn[e * n + i] + d];
          ue[3 * i + d] = u_col[3 * (int64_t)conn[e * n + i] + d];
        }
      if (mue_per_ele) par[0] = mue_per_ele[e];
      if (g_output_hist && is_eas(tech))
        g_eas_hist = g_output_hist + orc_eas_hist_size(tech) * e;
      orc_element_stress(ct, tech, matkind, par.data(), Xe.data(), ue.data(), stress_type, strain_type,
          stress ? stress + e * ng * 6 : nullptr, strain ? strain + e * ng * 6 : nullptr);
    }
  }

  double orc_calc_energy(int ct, int tech, int matkind, const double* matparams,
      const double* mue_per_ele, int64_t n_ele, const int32_t* conn, const double* X,
      const double* u_col, const uint8_t* ele_is_row)
  {
    const int n = nnodes(ct), nd = 3 * n;
    std::vector<double> Xe(nd), ue(nd);
    std::vector<double> par(matparams, matparams + (matkind == ORC_MAT_ELASTHYPER ? 1 + 6 * (int)matparams[0] : 2));
    double sum = 0.0;
    for (int64_t e = 0; e < n_ele; ++e)
    {
      if (ele_is_row && !ele_is_row[e]) continue;  // row elements only (model_evaluator_structure.cpp:1714)
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
          Xe[3 * i + d] = X[3 * (int64_t)conn[e * n + i] + d];
          ue[3 * i + d] = u_col[3 * (int64_t)conn[e * n + i] + d];
        }
      if (mue_per_ele) par[0] = mue_per_ele[e];
      if (g_output_hist && is_eas(tech))
        g_eas_hist = g_output_hist + orc_eas_hist_size(tech) * e;
      sum += orc_element_energy(ct, tech, matkind, par.data(), Xe.data(), ue.data());
    }
    return sum;
  }

  void orc_assemble_mass(int ct, double density, int lumped, int64_t n_ele, const int32_t* conn,
      const double* X, const double* acc_col, const int32_t* node_row_lid, const int32_t* node_col_lid,
      int64_t n_row_nodes, const int64_t* rowptr, const int32_t* colind, double* mass_values,
      double* f_inertia)
  {
    const int n = nnodes(ct), nd = 3 * n;
    std::vector<double> Me(nd * nd), fi(nd), Xe(nd), ae(nd);
    if (mass_values) std::fill(mass_values, mass_values + rowptr[3 * n_row_nodes], 0.0);
    if (f_inertia) std::fill(f_inertia, f_inertia + 3 * n_row_nodes, 0.0);
    for (int64_t e = 0; e < n_ele; ++e)
    {
      const int32_t* ce = conn + e * n;
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
          Xe[3 * i + d] = X[3 * (int64_t)ce[i] + d];
          ae[3 * i + d] = acc_col ? acc_col[3 * (int64_t)ce[i] + d] : 0.0;
        }
      orc_element_mass(ct, density, Xe.data(), lumped, acc_col ? ae.data() : nullptr, Me.data(), fi.data());
      for (int i = 0; i < n; ++i)
      {
        const int32_t r = node_row_lid[ce[i]];
        if (r < 0) continue;
        for (int b = 0; b < 3; ++b)
        {
          const int64_t rlid = 3 * (int64_t)r + b;
          if (f_inertia && acc_col) f_inertia[rlid] += fi[3 * i + b];
          if (!mass_values) continue;
          const int32_t* idx = colind + rowptr[rlid];
          const int len = (int)(rowptr[rlid + 1] - rowptr[rlid]);
          for (int j = 0; j < n; ++j)
            for (int c = 0; c < 3; ++c)
            {
              const int32_t col = 3 * node_col_lid[ce[j]] + c;
              const int32_t* loc = std::lower_bound(idx, idx + len, col);
              mass_values[rowptr[rlid] + (loc - idx)] += Me[(3 * i + b) + nd * (3 * j + c)];
            }
        }
      }
    }
  }

  // First assembly inserts every (owned row, element column) pair in global ids, zeros included
  // (src/core/linalg/src/sparse/4C_linalg_sparsematrix.cpp:469-494), complete() then sorts each
  // row by local column id (:700-721).  All three dofs of a node share the node's pattern.
  int orc_build_graph(int ct, int64_t n_ele, const int32_t* conn, int64_t n_col_nodes,
      const int32_t* node_row_lid, const int32_t* node_col_lid, int64_t n_row_nodes,
      int64_t* rowptr, int32_t* colind)
  {
    const int n = nnodes(ct);
    (void)n_col_nodes;
    std::vector<std::vector<int32_t>> adj(n_row_nodes);
    for (int64_t e = 0; e < n_ele; ++e)
      for (int i = 0; i < n; ++i)
      {
        const int32_t r = node_row_lid[conn[e * n + i]];
        if (r < 0) continue;
        for (int j = 0; j < n; ++j) adj[r].push_back(node_col_lid[conn[e * n + j]]);
      }
    int64_t nnz = 0;
    rowptr[0] = 0;
    for (int64_t r = 0; r < n_row_nodes; ++r)
    {
      auto& a = adj[r];
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
      for (int b = 0; b < 3; ++b)
      {
        if (colind)
          for (size_t q = 0; q < a.size(); ++q)
            for (int c = 0; c < 3; ++c) colind[nnz + 3 * q + c] = 3 * a[q] + c;
        nnz += 3 * (int64_t)a.size();
        rowptr[3 * r + b + 1] = nnz;
      }
    }
    return 0;
  }

  int64_t orc_evaluate_assemble(int ct, int matkind, const double* matparams,
      const double* mue_per_ele, int action, int64_t n_ele, const int32_t* conn,
      int64_t n_col_nodes, const double* X, const double* u_col, const int32_t* node_row_lid,
      const int32_t* node_col_lid, int64_t n_row_nodes, const int64_t* rowptr,
      const int32_t* colind, double* values, double* f, int nthreads)
  {
    return orc_evaluate_assemble_tech(ct, ORC_TECH_NONE, matkind, matparams, mue_per_ele, action,
        n_ele, conn, n_col_nodes, X, u_col, node_row_lid, node_col_lid, n_row_nodes, rowptr, colind,
        values, f, nthreads);
  }

  int64_t orc_evaluate_assemble_tech(int ct, int tech, int matkind, const double* matparams,
      const double* mue_per_ele, int action, int64_t n_ele, const int32_t* conn,
      int64_t n_col_nodes, const double* X, const double* u_col, const int32_t* node_row_lid,
      const int32_t* node_col_lid, int64_t n_row_nodes, const int64_t* rowptr,
      const int32_t* colind, double* values, double* f, int nthreads)
  {
    return orc_evaluate_assemble_hist(ct, tech, matkind, matparams, mue_per_ele, action, n_ele, conn,
        n_col_nodes, X, u_col, node_row_lid, node_col_lid, n_row_nodes, rowptr, colind, values, f,
        nthreads, nullptr);
  }

  int64_t orc_eas_hist_size(int tech)
  {
    const int neas = eas_num(tech);
    return 3 * neas + neas * neas + 24 * neas;
  }
  // struct_calc_recover over the column elements with the Newton increment on the column map
  // (Solid::ModelEvaluator::Structure::run_recover, model_evaluator_structure.cpp:1482-1509)
  void orc_eas_recover(int tech, int64_t n_ele, const int32_t* conn, const double* du_col,
      double step_length, double* hist)
  {
    const int64_t per = orc_eas_hist_size(tech);
    double due[24];
    for (int64_t e = 0; e < n_ele; ++e)
    {
      for (int i = 0; i < 8; ++i)
        for (int d = 0; d < 3; ++d) due[3 * i + d] = du_col[3 * (int64_t)conn[e * 8 + i] + d];
      element_eas_recover(tech, due, step_length, hist + per * e);
    }
  }

  void orc_mulf_init_history(int64_t n_points, double* hist)
  {
    for (int64_t p = 0; p < n_points; ++p)
    {
      double* h = hist + ORC_MULF_HIST * p;
      for (int i = 0; i < 9; ++i) h[i] = h[9 + i] = (i % 4 == 0) ? 1.0 : 0.0;
      h[18] = 0.0;
    }
  }

  // struct_update_prestress over the column elements (model_evaluator_structure.cpp:1608-1642)
  void orc_update_prestress_fbar(int ct, int64_t n_ele, const int32_t* conn, const double* X, const double* u_col,
      double* hist)
  {
    // MulfFBarFormulation: ngauss + 1 histories per element (the last one is the global / centroid history)
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    std::vector<double> Xe(nd), ue(nd);
    for (int64_t e = 0; e < n_ele; ++e)
    {
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
          Xe[3 * i + d] = X[3 * (int64_t)conn[e * n + i] + d];
          ue[3 * i + d] = u_col[3 * (int64_t)conn[e * n + i] + d];
        }
      element_update_prestress(ct, Xe.data(), ue.data(), hist + (int64_t)ORC_MULF_HIST * (ng + 1) * e, true);
    }
  }
  void orc_update_prestress(int ct, int64_t n_ele, const int32_t* conn, const double* X,
      const double* u_col, double* hist)
  {
    const int n = nnodes(ct), nd = 3 * n, ng = ngauss(ct);
    std::vector<double> Xe(nd), ue(nd);
    for (int64_t e = 0; e < n_ele; ++e)
    {
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d)
        {
          Xe[3 * i + d] = X[3 * (int64_t)conn[e * n + i] + d];
          ue[3 * i + d] = u_col[3 * (int64_t)conn[e * n + i] + d];
        }
      element_update_prestress(ct, Xe.data(), ue.data(), hist + (int64_t)ORC_MULF_HIST * ng * e);
    }
  }

  // hist: Gauss-point history of the column elements (n_ele * ngp * ORC_MULF_HIST), required for
  // ORC_TECH_MULF (use nthreads = 1: the first evaluate initialises it)
  int64_t orc_evaluate_assemble_hist(int ct, int tech, int matkind, const double* matparams,
      const double* mue_per_ele, int action, int64_t n_ele, const int32_t* conn,
      int64_t n_col_nodes, const double* X, const double* u_col, const int32_t* node_row_lid,
      const int32_t* node_col_lid, int64_t n_row_nodes, const int64_t* rowptr,
      const int32_t* colind, double* values, double* f, int nthreads, double* hist)
  {
    (void)n_col_nodes;
    const bool mat_hist = matkind == ORC_MAT_DRUCKER_PRAGER;  // history material: small strains only
    if (mat_hist && tech != ORC_TECH_LINEAR) return -1;
    const bool needs_hist = tech == ORC_TECH_MULF || tech == ORC_TECH_MULF_FBAR || is_eas(tech) || mat_hist;
    if (needs_hist && hist == nullptr) return -1;
    const int64_t hist_per_ele = mat_hist ? (int64_t)ORC_DP_HIST * ngauss(ct)
                                 : tech == ORC_TECH_MULF_FBAR ? (int64_t)ORC_MULF_HIST * (ngauss(ct) + 1)
                                 : tech == ORC_TECH_MULF ? (int64_t)ORC_MULF_HIST * ngauss(ct)
                                                         : (needs_hist ? orc_eas_hist_size(tech) : 0);
    const int n = nnodes(ct), nd = 3 * n;
    if (nthreads < 1) nthreads = 1;
    std::vector<int64_t> bad(nthreads, 0);
    auto worker = [&](int tid)
    {
      // "rank" tid owns row nodes [r0,r1)
      const int64_t r0 = n_row_nodes * tid / nthreads, r1 = n_row_nodes * (tid + 1) / nthreads;
      // SparseMatrix::zero / Vector::put_scalar(0) on the owned rows
      if (values && action == ORC_ACTION_FORCE_STIFF)
        std::fill(values + rowptr[3 * r0], values + rowptr[3 * r1], 0.0);
      std::fill(f + 3 * r0, f + 3 * r1, 0.0);
      std::vector<double> Ke(nd * nd), fe(nd), Xe(nd), ue(nd);
      std::vector<int32_t> localcol(nd);
      // generic MAT_ElastHyper: packed summand block (1 + 6 n doubles), used in place
      std::vector<double> par(matparams, matparams + (matkind == ORC_MAT_ELASTHYPER ? 1 + 6 * (int)matparams[0]
                                                      : (matkind == ORC_MAT_DRUCKER_PRAGER
                                                                ? 10
                                                                : (matkind == ORC_MAT_VISCONEOHOOKE
                                                                          ? 7
                                                                          : (matkind == ORC_MAT_PLASTIC_NLNLOG ? 11 : 2)))));
      par.resize(par.size() + 2, 0.0);
      // Discretization::evaluate loop over column elements
      // (src/core/fem/src/discretization/4C_fem_discretization_evaluate.cpp:85-117)
      for (int64_t e = 0; e < n_ele; ++e)
      {
        const int32_t* ce = conn + e * n;
        bool mine = false;
        for (int i = 0; i < n; ++i)
        {
          const int32_t r = node_row_lid[ce[i]];
          if (r >= r0 && r < r1) mine = true;
        }
        if (!mine) continue;
        for (int i = 0; i < n; ++i)
          for (int d = 0; d < 3; ++d)
          {
            Xe[3 * i + d] = X[3 * (int64_t)ce[i] + d];
            ue[3 * i + d] = u_col[3 * (int64_t)ce[i] + d];  // extract_values via lm
          }
        if (mue_per_ele && matkind == ORC_MAT_ELASTHYPER)
        {
          // element fibres (FIBER1 of the element line): 3 doubles per element into the ORC_SUM_FIBER entry
          for (int k = 0; k < (int)par[0]; ++k)
            if ((int)par[1 + 6 * k] == ORC_SUM_FIBER)
              for (int d = 0; d < 3; ++d) par[1 + 6 * k + 2 + d] = mue_per_ele[3 * e + d];
        }
        else if (mue_per_ele)
          par[0] = mue_per_ele[e];  // InputField MUE looked up by element id
        const bool stiff = (action == ORC_ACTION_FORCE_STIFF) && values;
        g_element_material_hist =
            g_material_hist_base ? g_material_hist_base + (int64_t)finite_strain_hist(matkind) * ngauss(ct) * e : nullptr;
        const int ele_err = element_tech(ct, tech, matkind, par.data(), action, Xe.data(), ue.data(),
            stiff ? Ke.data() : nullptr, fe.data(), hist ? hist + hist_per_ele * e : nullptr);
        g_element_material_hist = nullptr;
        if (ele_err)
        {
          bad[tid]++;
          continue;
        }
        // SparseMatrix::assemble, filled() branch (4C_linalg_sparsematrix.cpp:356-467)
        for (int j = 0; j < n; ++j)
          for (int c = 0; c < 3; ++c) localcol[3 * j + c] = 3 * node_col_lid[ce[j]] + c;
        for (int i = 0; i < n; ++i)
        {
          const int32_t r = node_row_lid[ce[i]];
          if (r < r0 || r >= r1) continue;  // lmrowowner != myrank
          for (int b = 0; b < 3; ++b)
          {
            const int lrow = 3 * i + b;
            const int64_t rlid = 3 * (int64_t)r + b;
            // LinAlg::assemble for the vector (4C_linalg_utils_sparse_algebra_assemble.cpp:60-82)
            f[rlid] += fe[lrow];
            if (!stiff) continue;
            const int length = (int)(rowptr[rlid + 1] - rowptr[rlid]);
            double* valview = values + rowptr[rlid];
            const int32_t* indices = colind + rowptr[rlid];
            int dofcount = 0, pos = 0;
            for (int node = 0; node < n; ++node)
            {
              if (pos >= length || indices[pos] != localcol[dofcount])
              {
                const int32_t* loc = std::lower_bound(indices, indices + length, localcol[dofcount]);
                pos = (int)(loc - indices);
              }
              const int stride = 3;
              bool continuous = true;
              if (stride + pos > length)
                continuous = false;
              else
                for (int j = 1; j < stride; ++j)
                  if (indices[pos + j] != localcol[dofcount + j])
                  {
                    continuous = false;
                    break;
                  }
              if (continuous)
                for (int j = 0; j < stride; ++j) valview[pos++] += Ke[lrow + nd * (dofcount++)];
              else
                for (int j = 0; j < stride; ++j)
                {  // sum_into_my_values fallback
                  const int32_t* loc =
                      std::lower_bound(indices, indices + length, localcol[dofcount]);
                  valview[loc - indices] += Ke[lrow + nd * dofcount];
                  dofcount++;
                }
            }
          }
        }
      }
    };
    if (nthreads == 1)
      worker(0);
    else
    {
      std::vector<std::thread> th;
      for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
      for (auto& t : th) t.join();
    }
    int64_t nb = 0;
    for (auto b : bad) nb += b;
    return nb;
  }
}

// shape_tables.hpp -- host-side generation of the constant per-cell-type tables the kernels use:
// parametric shape-function derivatives at the Gauss points and at the element nodes, and the
// Gauss weights.  For Lagrange cells these are compile-time constants of the reference
// (recomputed there per element and Gauss point); here they are built once
// per context and kept in constant/shared memory.
//
// Reference definitions:
//   shape functions  src/core/fem/src/general/utils/4C_fem_general_utils_fem_shapefunctions.hpp
//                    hex8 :409-450, hex27 :706-830, tet10 :831-875
//   Gauss rules      src/core/fem/src/general/utils/4C_fem_general_utils_integration.cpp
//                    hex_8point :74-104, hex_27point :130-245, tet_4point :331-352
//                    (selection src/solid_ele/4C_solid_ele_calc_lib_integration.hpp:43-64)
//   node parameter coordinates  4C_fem_general_utils_local_connectivity_matrices.hpp:291,314
#pragma once

#include <cmath>
#include <vector>

namespace b200s
{
  struct CellTables
  {
    int nn = 0;   // nodes per element
    int ngp = 0;  // Gauss points (force/stiffness rule)
    std::vector<double> dN_gp;    // [ngp][nn][3]  dN_k/dxi_a at Gauss point g
    std::vector<double> dN_node;  // [nn][nn][3]   same at the element nodes (Jacobian check)
    std::vector<double> w;        // [ngp]
    // mass-matrix rule (DisTypeToOptGaussRule, 4C_fem_general_element_integration_select.hpp:37-67:
    // hex_8point, hex_27point, tet_11point) -- equals the stiffness rule except for tet10
    int ngp_mass = 0;
    std::vector<double> N_mass;   // [ngp_mass][nn]     shape function values
    std::vector<double> dN_mass;  // [ngp_mass][nn][3]
    std::vector<double> w_mass;   // [ngp_mass]
    std::vector<double> N_gp;     // [ngp][nn] values at the stiffness rule's points
    std::vector<double> dN_cen;   // [nn][3] derivatives at the element centroid (F-bar, EAS)
    std::vector<double> dN_ans;   // hex8 only: [8][nn][3] derivatives at the ANS sampling points A..H (TECH shell_ans)
    std::vector<double> xi_gp;    // [ngp][3] parameter coordinates of the stiffness rule's points (EAS)
  };

  namespace detail
  {
    // hex27 node order; the first 8 are the hex8 corners
    inline const double (*hex_nodes())[3]
    {
      static const double c[27][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
          {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {0, -1, -1}, {1, 0, -1}, {0, 1, -1},
          {-1, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {1, 0, 1},
          {0, 1, 1}, {-1, 0, 1}, {0, 0, -1}, {0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0},
          {0, 0, 1}, {0, 0, 0}};
      return c;
    }
    // hex18 node order (4C_fem_general_utils_local_connectivity_matrices.hpp:174-179): the quad9 of the plane (corners,
    // mid-sides, centre) at t = -1, then at t = +1
    inline const double (*hex18_nodes())[3]
    {
      static const double c[18][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {0, -1, -1}, {1, 0, -1},
          {0, 1, -1}, {-1, 0, -1}, {0, 0, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}, {0, -1, 1}, {1, 0, 1},
          {0, 1, 1}, {-1, 0, 1}, {0, 0, 1}};
      return c;
    }
    inline const double (*tet_nodes())[3]
    {
      static const double c[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.5, 0, 0},
          {0.5, 0.5, 0}, {0, 0.5, 0}, {0, 0, 0.5}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};
      return c;
    }

    // dN[k][a] at xi for celltype 0 hex8, 1 hex27, 2 tet10
    inline void deriv1(int celltype, const double* xi, double* dN)
    {
      const double r = xi[0], s = xi[1], t = xi[2];
      if (celltype == 0)
      {
        const double(*c)[3] = hex_nodes();
        for (int k = 0; k < 8; ++k)
        {
          const double fr = 1.0 + c[k][0] * r, fs = 1.0 + c[k][1] * s, ft = 1.0 + c[k][2] * t;
          dN[3 * k + 0] = c[k][0] * 0.125 * fs * ft;
          dN[3 * k + 1] = c[k][1] * 0.125 * ft * fr;
          dN[3 * k + 2] = c[k][2] * 0.125 * fr * fs;
        }
      }
      else if (celltype == 1)
      {
        const double(*c)[3] = hex_nodes();
        double L[3][3], dL[3][3];
        for (int d = 0; d < 3; ++d)
        {
          const double x = xi[d];
          L[d][0] = 0.5 * x * (x - 1.0), L[d][1] = 1.0 - x * x, L[d][2] = 0.5 * x * (x + 1.0);
          dL[d][0] = x - 0.5, dL[d][1] = -2.0 * x, dL[d][2] = x + 0.5;
        }
        for (int k = 0; k < 27; ++k)
        {
          const int a = (int)c[k][0] + 1, b = (int)c[k][1] + 1, e = (int)c[k][2] + 1;
          dN[3 * k + 0] = L[1][b] * L[2][e] * dL[0][a];
          dN[3 * k + 1] = L[0][a] * L[2][e] * dL[1][b];
          dN[3 * k + 2] = L[0][a] * L[1][b] * dL[2][e];
        }
      }
      else if (celltype == 7)
      {
        // hex18 (:531-600): quadratic Lagrange functions in r and s times (1 -+ t) / 2
        const double(*c)[3] = hex18_nodes();
        double L[2][3], dL[2][3];
        for (int d = 0; d < 2; ++d)
        {
          const double x = xi[d];
          L[d][0] = 0.5 * x * (x - 1.0), L[d][1] = 1.0 - x * x, L[d][2] = 0.5 * x * (x + 1.0);
          dL[d][0] = x - 0.5, dL[d][1] = -2.0 * x, dL[d][2] = x + 0.5;
        }
        for (int k = 0; k < 18; ++k)
        {
          const int a = (int)c[k][0] + 1, b = (int)c[k][1] + 1;
          const double lin = 0.5 * (1.0 + c[k][2] * t);
          dN[3 * k + 0] = dL[0][a] * L[1][b] * lin;
          dN[3 * k + 1] = L[0][a] * dL[1][b] * lin;
          dN[3 * k + 2] = L[0][a] * L[1][b] * 0.5 * c[k][2];
        }
      }
      else if (celltype == 5)
      {
        // wedge6 (:877-903): triangle coordinates r, s, 1 - r - s times (1 -+ t) / 2
        const double tri[3] = {r, s, 1.0 - r - s}, dr[3] = {1.0, 0.0, -1.0}, ds[3] = {0.0, 1.0, -1.0};
        for (int k = 0; k < 6; ++k)
        {
          const double sg = k < 3 ? -1.0 : 1.0, lin = 0.5 * (1.0 + sg * t);
          dN[3 * k + 0] = dr[k % 3] * lin, dN[3 * k + 1] = ds[k % 3] * lin, dN[3 * k + 2] = 0.5 * sg * tri[k % 3];
        }
      }
      else if (celltype == 6)
      {
        // pyramid5 (:976-1018): bilinear base minus t plus / minus the rational term r s t / (1 - t); at the apex the
        // reference replaces the three derivatives of that term by 1
        const bool apex = !(std::fabs(t - 1.0) > 1e-14);
        const double qr = apex ? 1.0 : s * t / (1.0 - t), qs = apex ? 1.0 : r * t / (1.0 - t),
                     qt = apex ? 1.0 : r * s / ((1.0 - t) * (1.0 - t));
        const double c[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int k = 0; k < 4; ++k)
        {
          const double sg = c[k][0] * c[k][1];
          dN[3 * k + 0] = 0.25 * (c[k][0] * (1.0 + c[k][1] * s) + sg * qr);
          dN[3 * k + 1] = 0.25 * (c[k][1] * (1.0 + c[k][0] * r) + sg * qs);
          dN[3 * k + 2] = 0.25 * (sg * qt - 1.0);
        }
        dN[12] = 0.0, dN[13] = 0.0, dN[14] = 1.0;
      }
      else if (celltype == 3)
      {
        // tet4 (4C_fem_general_utils_fem_shapefunctions.hpp:813-830)
        const double D[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int k = 0; k < 4; ++k)
          for (int a = 0; a < 3; ++a) dN[3 * k + a] = D[k][a];
      }
      else if (celltype == 4)
      {
        // hex20 serendipity (:607-705): corners 1/8 c_a f_b f_e (2 f_a + f_b + f_e - 5), f_d = 1 + c_d x_d;
        // mid-edge node on an edge along a: -1/2 x_a f_b f_e, 1/4 (1 - x_a^2) c_b f_e, 1/4 (1 - x_a^2) f_b c_e
        const double(*c)[3] = hex_nodes();
        for (int k = 0; k < 20; ++k)
        {
          if (k < 8)
          {
            const double f[3] = {1.0 + c[k][0] * r, 1.0 + c[k][1] * s, 1.0 + c[k][2] * t};
            for (int a = 0; a < 3; ++a)
            {
              const int b = (a + 1) % 3, e = (a + 2) % 3;
              dN[3 * k + a] = 0.125 * c[k][a] * f[b] * f[e] * (2.0 * f[a] + f[b] + f[e] - 5.0);
            }
          }
          else
          {
            int a = 0;
            for (int d = 0; d < 3; ++d)
              if (c[k][d] == 0.0) a = d;
            const int b = (a + 1) % 3, e = (a + 2) % 3;
            const double q = 1.0 - xi[a] * xi[a], fb = 1.0 + c[k][b] * xi[b], fe = 1.0 + c[k][e] * xi[e];
            dN[3 * k + a] = -0.5 * xi[a] * fb * fe;
            dN[3 * k + b] = 0.25 * q * c[k][b] * fe;
            dN[3 * k + e] = 0.25 * q * fb * c[k][e];
          }
        }
      }
      else
      {
        const double u = 1.0 - r - s - t;
        const double d0 = -4.0 * u + 1.0;
        const double D[10][3] = {{d0, d0, d0}, {4.0 * r - 1.0, 0, 0}, {0, 4.0 * s - 1.0, 0},
            {0, 0, 4.0 * t - 1.0}, {4.0 * (u - r), -4.0 * r, -4.0 * r}, {4.0 * s, 4.0 * r, 0},
            {-4.0 * s, 4.0 * (u - s), -4.0 * s}, {-4.0 * t, -4.0 * t, 4.0 * (u - t)},
            {4.0 * t, 0, 4.0 * r}, {0, 4.0 * t, 4.0 * s}};
        for (int k = 0; k < 10; ++k)
          for (int a = 0; a < 3; ++a) dN[3 * k + a] = D[k][a];
      }
    }

    // N[k] at xi (same celltype ids); 4C_fem_general_utils_fem_shapefunctions.hpp hex8 :61-79,
    // hex27 :198-238, tet10 :252-267
    inline void shape0(int celltype, const double* xi, double* N)
    {
      const double r = xi[0], s = xi[1], t = xi[2];
      if (celltype == 0)
      {
        const double(*c)[3] = hex_nodes();
        for (int k = 0; k < 8; ++k)
          N[k] = 0.125 * (1.0 + c[k][0] * r) * (1.0 + c[k][1] * s) * (1.0 + c[k][2] * t);
      }
      else if (celltype == 1)
      {
        const double(*c)[3] = hex_nodes();
        double L[3][3];
        for (int d = 0; d < 3; ++d)
        {
          const double x = xi[d];
          L[d][0] = 0.5 * x * (x - 1.0), L[d][1] = 1.0 - x * x, L[d][2] = 0.5 * x * (x + 1.0);
        }
        for (int k = 0; k < 27; ++k)
          N[k] = L[0][(int)c[k][0] + 1] * L[1][(int)c[k][1] + 1] * L[2][(int)c[k][2] + 1];
      }
      else if (celltype == 7)
      {
        const double(*c)[3] = hex18_nodes();  // hex18 (:112-145)
        double L[2][3];
        for (int d = 0; d < 2; ++d)
        {
          const double x = xi[d];
          L[d][0] = 0.5 * x * (x - 1.0), L[d][1] = 1.0 - x * x, L[d][2] = 0.5 * x * (x + 1.0);
        }
        for (int k = 0; k < 18; ++k) N[k] = L[0][(int)c[k][0] + 1] * L[1][(int)c[k][1] + 1] * 0.5 * (1.0 + c[k][2] * t);
      }
      else if (celltype == 3)
      {
        N[0] = 1.0 - r - s - t, N[1] = r, N[2] = s, N[3] = t;  // tet4 (:239-251)
      }
      else if (celltype == 5)
      {
        const double tri[3] = {r, s, 1.0 - r - s};  // wedge6 (:268-279)
        for (int k = 0; k < 6; ++k) N[k] = tri[k % 3] * 0.5 * (k < 3 ? 1.0 - t : 1.0 + t);
      }
      else if (celltype == 6)
      {
        const double q = !(std::fabs(t - 1.0) > 1e-14) ? 0.0 : r * s * t / (1.0 - t);  // pyramid5 (:320-348)
        const double c[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int k = 0; k < 4; ++k) N[k] = 0.25 * ((1.0 + c[k][0] * r) * (1.0 + c[k][1] * s) - t + c[k][0] * c[k][1] * q);
        N[4] = t;
      }
      else if (celltype == 4)
      {
        const double(*c)[3] = hex_nodes();  // hex20 (:146-197)
        for (int k = 0; k < 20; ++k)
        {
          if (k < 8)
          {
            const double f[3] = {1.0 + c[k][0] * r, 1.0 + c[k][1] * s, 1.0 + c[k][2] * t};
            N[k] = 0.125 * f[0] * f[1] * f[2] * (f[0] + f[1] + f[2] - 5.0);
          }
          else
          {
            int a = 0;
            for (int d = 0; d < 3; ++d)
              if (c[k][d] == 0.0) a = d;
            const int b = (a + 1) % 3, e = (a + 2) % 3;
            N[k] = 0.25 * (1.0 - xi[a] * xi[a]) * (1.0 + c[k][b] * xi[b]) * (1.0 + c[k][e] * xi[e]);
          }
        }
      }
      else
      {
        const double u = 1.0 - r - s - t;
        N[0] = u * (2.0 * u - 1.0), N[1] = r * (2.0 * r - 1.0), N[2] = s * (2.0 * s - 1.0);
        N[3] = t * (2.0 * t - 1.0), N[4] = 4.0 * r * u, N[5] = 4.0 * r * s, N[6] = 4.0 * s * u;
        N[7] = 4.0 * t * u, N[8] = 4.0 * r * t, N[9] = 4.0 * s * t;
      }
    }
  }  // namespace detail

  // Shape-derivative tables of a hexahedral cell with <= 27 nodes and <= 27 Gauss points in the layout of the hex27 kernel
  // (dN[g * 81 + 3 k + a], w[g]): zero derivatives for the padding nodes, and padding Gauss points that repeat point 0 with
  // weight zero -- they contribute exactly nothing (hex20: 7 padding nodes; hex18: 9 padding nodes, 9 padding points)
  inline void pad_tables_to_hex27(const CellTables& T, double* dN /*27 * 81*/, double* w /*27*/)
  {
    for (int g = 0; g < 27; ++g)
    {
      const int gs = g < T.ngp ? g : 0;
      for (int k = 0; k < 27; ++k)
        for (int a = 0; a < 3; ++a) dN[g * 81 + 3 * k + a] = k < T.nn ? T.dN_gp[((size_t)gs * T.nn + k) * 3 + a] : 0.0;
      w[g] = g < T.ngp ? T.w[g] : 0.0;
    }
  }

  inline CellTables make_cell_tables(int celltype)
  {
    CellTables T;
    // celltype 0 hex8, 1 hex27, 2 tet10, 3 tet4, 4 hex20; stiffness rules hex_8point, hex_27point (hex27 and hex20),
    // tet_4point (tet10), tet_1point (tet4) (4C_solid_ele_calc_lib_integration.hpp:43-64)
    // 5 wedge6 (wedge_6point), 6 pyramid5 (pyramid_8point): DisTypeToOptGaussRule for stiffness and mass
    // (4C_fem_general_element_integration_select.hpp:68-82)
    // 7 hex18 (hex_18point = quad_9point x line_2point, 4C_fem_general_utils_integration.cpp:112-129)
    T.nn = celltype == 0 ? 8 : (celltype == 1 ? 27 : (celltype == 3 ? 4 : (celltype == 4 ? 20 : (celltype == 5 ? 6 : (celltype == 6 ? 5 : (celltype == 7 ? 18 : 10))))));
    T.ngp = celltype == 0 ? 8 : ((celltype == 1 || celltype == 4) ? 27 : (celltype == 3 ? 1 : (celltype == 5 ? 6 : (celltype == 6 ? 8 : (celltype == 7 ? 18 : 4)))));
    std::vector<double> gxi(3 * T.ngp);
    T.w.resize(T.ngp);
    const double(*hc)[3] = detail::hex_nodes();
    if (celltype == 0)
    {
      const double g = 1.0 / std::sqrt(3.0);
      for (int q = 0; q < 8; ++q)
      {
        for (int d = 0; d < 3; ++d) gxi[3 * q + d] = hc[q][d] * g;
        T.w[q] = 1.0;
      }
    }
    else if (celltype == 7)
    {
      // quad_9point (:6185-6233: corners, mid-sides, centre; truncated decimals as tabulated) at the two points of
      // line_2point (:6946-6953)
      const double g = 0.7745966692415, w1 = 0.5555555555556, w2 = 0.8888888888889, d = 0.5773502691896;
      const double(*c)[3] = detail::hex18_nodes();
      for (int q = 0; q < 18; ++q)
      {
        gxi[3 * q] = c[q % 9][0] * g, gxi[3 * q + 1] = c[q % 9][1] * g, gxi[3 * q + 2] = q < 9 ? -d : d;
        T.w[q] = (c[q % 9][0] == 0.0 ? w2 : w1) * (c[q % 9][1] == 0.0 ? w2 : w1);
      }
    }
    else if (celltype == 3)
    {
      gxi[0] = gxi[1] = gxi[2] = 0.25;  // tet_1point (4C_fem_general_utils_integration.cpp:322-330)
      T.w[0] = 1.0 / 6.0;
    }
    else if (celltype == 5)
    {
      // wedge_6point (:5992-6026): the 3-point triangle rule at t = +-1/sqrt(3)
      const double g = 1.0 / std::sqrt(3.0), a = 2.0 / 3.0, b = 1.0 / 6.0;
      const double tri[3][2] = {{a, b}, {b, a}, {b, b}};
      for (int q = 0; q < 6; ++q)
      {
        gxi[3 * q] = tri[q % 3][0], gxi[3 * q + 1] = tri[q % 3][1], gxi[3 * q + 2] = q < 3 ? g : -g;
        T.w[q] = b;
      }
    }
    else if (celltype == 6)
    {
      // pyramid_8point (:6092-6132), decimals as the reference has them
      const double a1 = 0.26318405556971, a2 = 0.50661630334979, t1 = 0.54415184401122, t2 = 0.12251482265544;
      for (int q = 0; q < 8; ++q)
      {
        const double a = (q & 1) ? a2 : a1;
        gxi[3 * q] = q < 4 ? -a : a, gxi[3 * q + 1] = (q & 2) ? a : -a, gxi[3 * q + 2] = (q & 1) ? t2 : t1;
        T.w[q] = (q & 1) ? 0.23254745125351 : 0.10078588207983;
      }
    }
    else if (celltype == 1 || celltype == 4)
    {
      // truncated decimals exactly as the reference has them
      const double xi3 = 0.7745966692415, w1 = 0.5555555555556, w2 = 0.8888888888889;
      for (int q = 0; q < 27; ++q)
      {
        for (int d = 0; d < 3; ++d) gxi[3 * q + d] = hc[q][d] * xi3;
        T.w[q] = (hc[q][0] == 0.0 ? w2 : w1) * (hc[q][1] == 0.0 ? w2 : w1) *
                 (hc[q][2] == 0.0 ? w2 : w1);
      }
    }
    else
    {
      const double pa = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0, pb = (5.0 - std::sqrt(5.0)) / 20.0;
      const double P[4][3] = {{pb, pb, pb}, {pa, pb, pb}, {pb, pa, pb}, {pb, pb, pa}};
      for (int q = 0; q < 4; ++q)
      {
        for (int d = 0; d < 3; ++d) gxi[3 * q + d] = P[q][d];
        T.w[q] = 1.0 / 6.0 / 4.0;
      }
    }
    T.dN_gp.resize((size_t)T.ngp * T.nn * 3);
    for (int q = 0; q < T.ngp; ++q) detail::deriv1(celltype, &gxi[3 * q], &T.dN_gp[(size_t)q * T.nn * 3]);
    T.dN_node.resize((size_t)T.nn * T.nn * 3);
    // nodal parameter coordinates as the reference tabulates them (4C_fem_general_utils_local_connectivity_matrices.hpp:
    // wedge6 reads the first rows of the wedge18 table :321-325, pyramid5 :318-319)
    static const double wedge_nodes[6][3] = {{0, 0, -1}, {1, 0, -1}, {0, 1, -1}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}};
    static const double pyramid_nodes[5][3] = {{-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}, {0, 0, 1}};
    const double(*nc)[3] = (celltype == 2 || celltype == 3) ? detail::tet_nodes()
                           : (celltype == 5 ? wedge_nodes : (celltype == 6 ? pyramid_nodes : (celltype == 7 ? detail::hex18_nodes() : hc)));
    for (int k = 0; k < T.nn; ++k) detail::deriv1(celltype, nc[k], &T.dN_node[(size_t)k * T.nn * 3]);
    T.xi_gp = gxi;
    T.N_gp.resize((size_t)T.ngp * T.nn);
    for (int q = 0; q < T.ngp; ++q) detail::shape0(celltype, &gxi[3 * q], &T.N_gp[(size_t)q * T.nn]);
    // mass rule
    std::vector<double> mxi = gxi;
    T.w_mass = T.w, T.ngp_mass = T.ngp;
    if (celltype == 3)
    {
      // DisTypeToOptGaussRule<tet4> = tet_4point (:331-352)
      const double pa = (5.0 + 3.0 * std::sqrt(5.0)) / 20.0, pb = (5.0 - std::sqrt(5.0)) / 20.0;
      const double P[4][3] = {{pb, pb, pb}, {pa, pb, pb}, {pb, pa, pb}, {pb, pb, pa}};
      T.ngp_mass = 4;
      mxi.assign(&P[0][0], &P[0][0] + 12);
      T.w_mass.assign(4, 1.0 / 24.0);
    }
    if (celltype == 2)
    {
      // tet_11point (4C_fem_general_utils_integration.cpp:447-495), decimals as the reference has them
      const double a = 0.0714285714285714, b = 0.7857142857142857, c = 0.1005964238332008,
                   d = 0.3994035761667992;
      const double P[11][3] = {{0.25, 0.25, 0.25}, {b, a, a}, {a, a, a}, {a, a, b}, {a, b, a},
          {c, d, d}, {d, c, d}, {d, d, c}, {d, c, c}, {c, d, c}, {c, c, d}};
      const double Q16 = 1.0 / 6.0;
      T.ngp_mass = 11;
      mxi.assign(&P[0][0], &P[0][0] + 33);
      T.w_mass.assign(11, 0.1493333333333333 * Q16);
      T.w_mass[0] = -0.0789333333333333 * Q16;
      for (int q = 1; q < 5; ++q) T.w_mass[q] = 0.0457333333333333 * Q16;
    }
    T.N_mass.resize((size_t)T.ngp_mass * T.nn);
    T.dN_mass.resize((size_t)T.ngp_mass * T.nn * 3);
    for (int q = 0; q < T.ngp_mass; ++q)
    {
      detail::shape0(celltype, &mxi[3 * q], &T.N_mass[(size_t)q * T.nn]);
      detail::deriv1(celltype, &mxi[3 * q], &T.dN_mass[(size_t)q * T.nn * 3]);
    }
    const double cen_hex[3] = {0.0, 0.0, 0.0}, cen_tet[3] = {0.25, 0.25, 0.25}, cen_wedge[3] = {1.0 / 3.0, 1.0 / 3.0, 0.0},
                 cen_pyramid[3] = {0.0, 0.0, 0.25};
    if (celltype == 0)
    {
      // SamplingPoints<hex8> (4C_solid_ele_calc_shell_ans_lib.hpp:26-33): A..D for the transverse shears, E..H for tt
      const double sp[8][3] = {{0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
      T.dN_ans.resize((size_t)8 * T.nn * 3);
      for (int q = 0; q < 8; ++q) detail::deriv1(celltype, sp[q], &T.dN_ans[(size_t)q * T.nn * 3]);
    }
    T.dN_cen.resize((size_t)T.nn * 3);
    detail::deriv1(celltype, (celltype == 2 || celltype == 3) ? cen_tet : (celltype == 5 ? cen_wedge : (celltype == 6 ? cen_pyramid : cen_hex)),
        T.dN_cen.data());
    return T;
  }
  // Surface cells of the three solid cell types (hex8 -> quad4, hex27 -> quad9, tet10 -> tri6) at the
  // Gauss rule SolidSurface uses (get_optimal_gauss_rule, src/solid_ele/4C_solid_ele_surface.cpp:58-76):
  // shape functions 4C_fem_general_utils_fem_shapefunctions.hpp quad4 :2025, quad9 :2087, tri6 :2120 and
  // their derivatives :2184,:2269,:2324; rules 4C_fem_general_utils_integration.cpp quad_4point :6151,
  // quad_9point :6185, tri_6point :6374 (decimals as the reference has them).
  struct FaceTables
  {
    int nn = 0, ngp = 0;
    std::vector<double> N, d0, d1, w;  // [ngp][nn] x3, [ngp]
  };

  inline FaceTables make_face_tables(int celltype)
  {
    FaceTables T;
    if (celltype == 5 || celltype == 6 || celltype == 7) return T;  // wedge6 / pyramid5 / hex18 have two surface shapes: no single face table
    std::vector<double> xi;
    if (celltype == 0)
    {
      const double g = 0.5773502691896;
      T.nn = 4, T.ngp = 4;
      xi = {-g, -g, g, -g, g, g, -g, g};
      T.w.assign(4, 1.0);
    }
    else if (celltype == 1)
    {
      const double g = 0.7745966692415, w1 = 0.5555555555556, w2 = 0.8888888888889;
      T.nn = 9, T.ngp = 9;
      xi = {-g, -g, g, -g, g, g, -g, g, 0, -g, g, 0, 0, g, -g, 0, 0, 0};
      T.w = {w1 * w1, w1 * w1, w1 * w1, w1 * w1, w2 * w1, w1 * w2, w2 * w1, w1 * w2, w2 * w2};
    }
    else if (celltype == 3)
    {
      // tet4 -> tri3 at tri_3point (:6341-6355)
      T.nn = 3, T.ngp = 3;
      xi = {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0};
      T.w.assign(3, 1.0 / 6.0);
    }
    else if (celltype == 4)
    {
      // hex20 -> quad8 at quad_9point
      const double g = 0.7745966692415, w1 = 0.5555555555556, w2 = 0.8888888888889;
      T.nn = 8, T.ngp = 9;
      xi = {-g, -g, g, -g, g, g, -g, g, 0, -g, g, 0, 0, g, -g, 0, 0, 0};
      T.w = {w1 * w1, w1 * w1, w1 * w1, w1 * w1, w2 * w1, w1 * w2, w2 * w1, w1 * w2, w2 * w2};
    }
    else
    {
      const double a = 0.0915762135098, b = 0.8168475729805, c = 0.4459484909160, d = 0.1081030181681;
      const double wa = 0.0549758718277, wb = 0.1116907948390;
      T.nn = 6, T.ngp = 6;
      xi = {a, a, b, a, a, b, c, d, c, c, d, c};
      T.w = {wa, wa, wa, wb, wb, wb};
    }
    T.N.resize((size_t)T.ngp * T.nn), T.d0.resize(T.N.size()), T.d1.resize(T.N.size());
    for (int q = 0; q < T.ngp; ++q)
    {
      const double r = xi[2 * q], s = xi[2 * q + 1];
      double* N = &T.N[(size_t)q * T.nn];
      double* d0 = &T.d0[(size_t)q * T.nn];
      double* d1 = &T.d1[(size_t)q * T.nn];
      if (celltype == 0)
      {
        const double c[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int k = 0; k < 4; ++k)
        {
          N[k] = 0.25 * (1.0 + c[k][0] * r) * (1.0 + c[k][1] * s);
          d0[k] = 0.25 * c[k][0] * (1.0 + c[k][1] * s);
          d1[k] = 0.25 * c[k][1] * (1.0 + c[k][0] * r);
        }
      }
      else if (celltype == 1)
      {
        const int c[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
        const double Lr[3] = {0.5 * r * (r - 1.0), 1.0 - r * r, 0.5 * r * (r + 1.0)};
        const double Ls[3] = {0.5 * s * (s - 1.0), 1.0 - s * s, 0.5 * s * (s + 1.0)};
        const double dLr[3] = {r - 0.5, -2.0 * r, r + 0.5};
        const double dLs[3] = {s - 0.5, -2.0 * s, s + 0.5};
        for (int k = 0; k < 9; ++k)
        {
          const int i = c[k][0] + 1, j = c[k][1] + 1;
          N[k] = Lr[i] * Ls[j], d0[k] = dLr[i] * Ls[j], d1[k] = Lr[i] * dLs[j];
        }
      }
      else if (celltype == 3)
      {
        N[0] = 1.0 - r - s, N[1] = r, N[2] = s;  // tri3 (:2110-2119, 2312-2323)
        d0[0] = -1.0, d0[1] = 1.0, d0[2] = 0.0;
        d1[0] = -1.0, d1[1] = 0.0, d1[2] = 1.0;
      }
      else if (celltype == 4)
      {
        // quad8 serendipity (:2049-2086): corner k 1/4 [f_r f_s - (1 - r^2) f_s - (1 - s^2) f_r], mid-side nodes
        // 1/2 (1 - r^2) f_s (on s = -+1) and 1/2 (1 - s^2) f_r (on r = +-1)
        const double c[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}}, m[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
        const double r2 = 1.0 - r * r, s2 = 1.0 - s * s;
        for (int k = 0; k < 4; ++k)
        {
          const double fr = 1.0 + c[k][0] * r, fs = 1.0 + c[k][1] * s;
          N[k] = 0.25 * (fr * fs - (r2 * fs + s2 * fr));
          d0[k] = 0.25 * (c[k][0] * fs - (-2.0 * r * fs + s2 * c[k][0]));
          d1[k] = 0.25 * (fr * c[k][1] - (r2 * c[k][1] - 2.0 * s * fr));
        }
        for (int k = 0; k < 4; ++k)
        {
          if (m[k][0] == 0.0)
          {
            const double fs = 1.0 + m[k][1] * s;
            N[4 + k] = 0.5 * r2 * fs, d0[4 + k] = -r * fs, d1[4 + k] = 0.5 * r2 * m[k][1];
          }
          else
          {
            const double fr = 1.0 + m[k][0] * r;
            N[4 + k] = 0.5 * s2 * fr, d0[4 + k] = 0.5 * s2 * m[k][0], d1[4 + k] = -s * fr;
          }
        }
      }
      else
      {
        const double t1 = 1.0 - r - s, t2 = r, t3 = s;
        const double n[6] = {t1 * (2.0 * t1 - 1.0), t2 * (2.0 * t2 - 1.0), t3 * (2.0 * t3 - 1.0),
            4.0 * t2 * t1, 4.0 * t2 * t3, 4.0 * t3 * t1};
        const double a0[6] = {-(4.0 * t1 - 1.0), 4.0 * t2 - 1.0, 0.0, 4.0 * (t1 - t2), 4.0 * t3, -4.0 * t3};
        const double a1[6] = {-(4.0 * t1 - 1.0), 0.0, 4.0 * t3 - 1.0, -4.0 * t2, 4.0 * t2, 4.0 * (t1 - t3)};
        for (int k = 0; k < 6; ++k) N[k] = n[k], d0[k] = a0[k], d1[k] = a1[k];
      }
    }
    return T;
  }
}  // namespace b200s

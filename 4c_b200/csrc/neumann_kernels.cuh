// neumann_kernels.cuh -- external loads of the solid element block on the device:
// surface Neumann conditions on faces of the block's elements and the volume (body) load.
//
// Reference:
//   Discretization::evaluate_neumann   src/core/fem/src/discretization/4C_fem_discretization_evaluate.cpp:264-297
//       (loop over the condition geometry incl. ghosted faces, owner-computes assemble of the force
//        vector and -- with LOADLIN -- of the load stiffness into the system matrix)
//   SolidSurface::evaluate_neumann     src/solid_ele/4C_solid_ele_surface_evaluate.cpp:43-385
//       TYPE Live (dead load, material configuration :279-317), orthopressure / pseudo_orthopressure
//       (follower pressure on the spatial / last converged configuration :319-378, its linearisation
//       analytical_d_surface_integration :502-547)
//   evaluate_neumann<celltype> (volume) src/solid_ele/4C_solid_ele_neumann_evaluator.cpp:43-113
//
// Functions of space and time stay with the host's function manager: the caller hands in one factor
// per condition dof (the time curve) and optionally one per Gauss point and dof, evaluated at the
// positions b200s_neumann_gp_coords returns.
//
// The per-thread functions are __host__ __device__ (tests/cpp/emulate_neumann.cpp runs them on the
// CPU against the test-side restatement).
#pragma once

#include <stdint.h>

#include "hex8_kernel.cuh"
#include "solid_math.cuh"

namespace b200s
{
  // shape functions of the surface cell at its Gauss rule (shape_tables.hpp make_face_tables)
  struct FaceTab
  {
    int nn, ngp;
    const double* N;   // [ngp][nn]
    const double* d0;  // [ngp][nn]  dN/dr
    const double* d1;  // [ngp][nn]  dN/ds
    const double* w;   // [ngp]
  };
  constexpr int FACE_MAX_NODES = 9;

  enum NeumannKind
  {
    NEUMANN_LIVE = 0,
    NEUMANN_PSEUDO_ORTHOPRESSURE = 1,
    NEUMANN_ORTHOPRESSURE = 2
  };

  struct NeumannSurfArgs
  {
    FaceTab T;
    const double* X;        // 3*n_col_nodes
    const double* u;        // 3*n_col_nodes or nullptr (material configuration)
    const int32_t* faces;   // [n_faces][nn] column-node lids, node order of the parent's surface
    int64_t n_faces;
    int kind;
    int onoff[3];
    double val[3];          // VAL * factor of the time function, per dof
    const double* gp_fac;   // [n_faces][ngp][3] or nullptr
    const NodeRow* node_row;
    const int32_t* node_colbase;
    const int32_t* colind;
    double* fext;           // n_rows
    double* K;              // nnz or nullptr (no load linearisation)
  };

  // nodes of face f in the configuration the load acts on
  B2_HD void face_coordinates(const NeumannSurfArgs& A, int64_t f, double* x)
  {
    const int32_t* fn = A.faces + f * A.T.nn;
    for (int b = 0; b < A.T.nn; ++b)
    {
      const int64_t n = fn[b];
#pragma unroll
      for (int d = 0; d < 3; ++d) x[3 * b + d] = A.X[3 * n + d] + (A.u ? A.u[3 * n + d] : 0.0);
    }
  }

  // covariant basis vectors dx/dr, dx/ds at Gauss point g
  B2_HD void face_basis(const FaceTab& T, int g, const double* x, double* g1, double* g2)
  {
#pragma unroll
    for (int d = 0; d < 3; ++d) g1[d] = 0.0, g2[d] = 0.0;
    for (int b = 0; b < T.nn; ++b)
    {
      const double a0 = T.d0[g * T.nn + b], a1 = T.d1[g * T.nn + b];
#pragma unroll
      for (int d = 0; d < 3; ++d) g1[d] += a0 * x[3 * b + d], g2[d] += a1 * x[3 * b + d];
    }
  }

  // load vector entries of face node a (SolidSurface::evaluate_neumann :279-354)
  B2_HD void neumann_face_force(const NeumannSurfArgs& A, int64_t f, const double* x, int a, double* out)
  {
    out[0] = out[1] = out[2] = 0.0;
    for (int g = 0; g < A.T.ngp; ++g)
    {
      double g1[3], g2[3];
      face_basis(A.T, g, x, g1, g2);
      const double Na = A.T.N[g * A.T.nn + a];
      const double* gf = A.gp_fac ? A.gp_fac + (f * A.T.ngp + g) * 3 : nullptr;
      if (A.kind == NEUMANN_LIVE)
      {
        const double m00 = g1[0] * g1[0] + g1[1] * g1[1] + g1[2] * g1[2];
        const double m11 = g2[0] * g2[0] + g2[1] * g2[1] + g2[2] * g2[2];
        const double m01 = g1[0] * g2[0] + g1[1] * g2[1] + g1[2] * g2[2];
        const double detA = sqrt(m00 * m11 - m01 * m01);
#pragma unroll
        for (int d = 0; d < 3; ++d)
          if (A.onoff[d]) out[d] += Na * (A.T.w[g] * detA * A.val[d] * (gf ? gf[d] : 1.0));
      }
      else
      {
        // normal of length dA (surface_integration :380-393)
        const double n0 = g1[1] * g2[2] - g1[2] * g2[1];
        const double n1 = g1[2] * g2[0] - g1[0] * g2[2];
        const double n2 = g1[0] * g2[1] - g1[1] * g2[0];
        const double fac = A.T.w[g] * (gf ? gf[0] : 1.0) * A.val[0];
        out[0] += Na * n0 * fac, out[1] += Na * n1 * fac, out[2] += Na * n2 * fac;
      }
    }
  }

  // load stiffness block (row node a, column node b) of the follower pressure, blk[dim + 3*k]:
  // -N_a * d normal_dim / d x_{b,k} * fac  (:356-377, analytical_d_surface_integration :502-547)
  B2_HD void neumann_face_lin_block(const NeumannSurfArgs& A, int64_t f, const double* x, int a, int b,
      double* blk)
  {
#pragma unroll
    for (int t = 0; t < 9; ++t) blk[t] = 0.0;
    for (int g = 0; g < A.T.ngp; ++g)
    {
      double g1[3], g2[3];
      face_basis(A.T, g, x, g1, g2);
      const double* gf = A.gp_fac ? A.gp_fac + (f * A.T.ngp + g) * 3 : nullptr;
      const double s = -A.T.N[g * A.T.nn + a] * A.T.w[g] * (gf ? gf[0] : 1.0) * A.val[0];
      const double d0 = A.T.d0[g * A.T.nn + b], d1 = A.T.d1[g * A.T.nn + b];
      // d(g1 x g2)/dx_{b,k} = d0 (e_k x g2) + d1 (g1 x e_k)
      blk[0 + 3 * 1] += s * (d0 * g2[2] - g1[2] * d1);
      blk[0 + 3 * 2] += s * (g1[1] * d1 - d0 * g2[1]);
      blk[1 + 3 * 2] += s * (d0 * g2[0] - g1[0] * d1);
      blk[1 + 3 * 0] += s * (g1[2] * d1 - d0 * g2[2]);
      blk[2 + 3 * 0] += s * (d0 * g2[1] - g1[1] * d1);
      blk[2 + 3 * 1] += s * (g1[0] * d1 - d0 * g2[0]);
    }
  }

  // CSR offset of the three columns of node b inside each row of node a (lower_bound as in
  // SparseMatrix::assemble, 4C_linalg_sparsematrix.cpp:404-441); -1 if missing
  B2_HD int neumann_find_slot(const NodeRow& nr, const int32_t* colind, int32_t col)
  {
    const int32_t* idx = colind + nr.rowstart;
    int lo = 0, hi = nr.rowlen;
    while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (idx[mid] < col)
        lo = mid + 1;
      else
        hi = mid;
    }
    return (lo + 3 <= nr.rowlen && idx[lo] == col && idx[lo + 2] == col + 2) ? lo : -1;
  }

  struct NeumannVolArgs
  {
    const double* X;
    const int32_t* conn;      // device element order
    const int32_t* ele_id;    // device position -> caller's element id (nullptr = identity)
    const int32_t* ele_slot;  // caller's element id -> index inside the condition, -1 = not loaded; nullptr = all
    const double* N_gp;       // [ngp][nn]
    const double* dN_gp;      // [ngp][nn][3]
    const double* w;          // [ngp]
    int64_t n_ele;
    int nn, ngp;
    int onoff[3];
    double val[3];
    const double* gp_fac;     // [n_loaded][ngp][3] or nullptr
    const NodeRow* node_row;
    double* fext;
  };

  // body load entries of element node a (4C_solid_ele_neumann_evaluator.cpp:71-111)
  B2_HD bool neumann_volume_force(const NeumannVolArgs& A, int64_t e, int a, double* out)
  {
    out[0] = out[1] = out[2] = 0.0;
    const int64_t eid = A.ele_id ? A.ele_id[e] : e;
    const int64_t slot = A.ele_slot ? A.ele_slot[eid] : eid;
    if (slot < 0) return false;
    const int32_t* ce = A.conn + e * A.nn;
    for (int g = 0; g < A.ngp; ++g)
    {
      const double* dN = A.dN_gp + (size_t)g * A.nn * 3;
      double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < A.nn; ++k)
      {
        const double* xk = A.X + 3 * (int64_t)ce[k];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) J[p + 3 * q] += dN[3 * k + p] * xk[q];
      }
      const double ifac = det3(J) * A.w[g];
      const double Na = A.N_gp[g * A.nn + a];
      const double* gf = A.gp_fac ? A.gp_fac + (slot * A.ngp + g) * 3 : nullptr;
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (A.onoff[d]) out[d] += Na * (A.val[d] * (gf ? gf[d] : 1.0) * ifac);
    }
    return true;
  }

  // position of Gauss point g of face f / element e (for the host's FunctionOfSpaceTime)
  B2_HD void neumann_face_gp_position(const NeumannSurfArgs& A, const double* x, int g, double* p)
  {
    p[0] = p[1] = p[2] = 0.0;
    for (int b = 0; b < A.T.nn; ++b)
    {
      const double Nb = A.T.N[g * A.T.nn + b];
#pragma unroll
      for (int d = 0; d < 3; ++d) p[d] += Nb * x[3 * b + d];
    }
  }

#if defined(__CUDACC__)
  // one thread per (face, node a): force vector
  __global__ void k_neumann_surface_force(NeumannSurfArgs A)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n_faces * A.T.nn) return;
    const int64_t f = t / A.T.nn;
    const int a = (int)(t % A.T.nn);
    const NodeRow nr = A.node_row[A.faces[f * A.T.nn + a]];
    if (nr.frow < 0) return;  // owner computes
    double x[3 * FACE_MAX_NODES], out[3];
    face_coordinates(A, f, x);
    neumann_face_force(A, f, x, a, out);
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (out[d] != 0.0) B2_RED_ADD(A.fext + nr.frow + d, out[d]);
  }

  // one thread per (face, node a, node b): load stiffness of the follower pressure
  __global__ void k_neumann_surface_lin(NeumannSurfArgs A, int* flag)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = A.T.nn;
    if (t >= A.n_faces * nn * nn) return;
    const int64_t f = t / (nn * nn);
    const int a = (int)((t / nn) % nn), b = (int)(t % nn);
    const NodeRow nr = A.node_row[A.faces[f * nn + a]];
    if (nr.rowstart < 0) return;
    const int lo = neumann_find_slot(nr, A.colind, A.node_colbase[A.faces[f * nn + b]]);
    if (lo < 0)
    {
      atomicOr(flag, 1);
      return;
    }
    double x[3 * FACE_MAX_NODES], blk[9];
    face_coordinates(A, f, x);
    neumann_face_lin_block(A, f, x, a, b, blk);
#pragma unroll
    for (int dim = 0; dim < 3; ++dim)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (dim != k) B2_RED_ADD(A.K + nr.rowstart + (int64_t)dim * nr.rowlen + lo + k, blk[dim + 3 * k]);
  }

  __global__ void k_neumann_surface_gp_coords(NeumannSurfArgs A, double* out)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n_faces * A.T.ngp) return;
    const int64_t f = t / A.T.ngp;
    double x[3 * FACE_MAX_NODES];
    face_coordinates(A, f, x);
    neumann_face_gp_position(A, x, (int)(t % A.T.ngp), out + 3 * t);
  }

  // one thread per (element, node a): body load
  __global__ void k_neumann_volume_force(NeumannVolArgs A)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n_ele * A.nn) return;
    const int64_t e = t / A.nn;
    const int a = (int)(t % A.nn);
    const NodeRow nr = A.node_row[A.conn[e * A.nn + a]];
    if (nr.frow < 0) return;
    double out[3];
    if (!neumann_volume_force(A, e, a, out)) return;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (out[d] != 0.0) B2_RED_ADD(A.fext + nr.frow + d, out[d]);
  }

  // reference position of the Gauss points of the loaded elements, [slot][gp][3]
  __global__ void k_neumann_volume_gp_coords(NeumannVolArgs A, double* out)
  {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n_ele * A.ngp) return;
    const int64_t e = t / A.ngp;
    const int g = (int)(t % A.ngp);
    const int64_t eid = A.ele_id ? A.ele_id[e] : e;
    const int64_t slot = A.ele_slot ? A.ele_slot[eid] : eid;
    if (slot < 0) return;
    double p[3] = {0, 0, 0};
    for (int k = 0; k < A.nn; ++k)
    {
      const double Nk = A.N_gp[g * A.nn + k];
      const double* xk = A.X + 3 * (int64_t)A.conn[e * A.nn + k];
      for (int d = 0; d < 3; ++d) p[d] += Nk * xk[d];
    }
    for (int d = 0; d < 3; ++d) out[(slot * A.ngp + g) * 3 + d] = p[d];
  }
#endif
}  // namespace b200s

/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * oracle/_ref/libref_solid.so: the element / material arithmetic of the hot path evaluated with
 * the REFERENCE'S OWN header-only numerics, included where they lie under /root/reference
 * (nothing is copied into this repo):
 *   Core::LinAlg::Tensor / SymmetricTensor / einsum / einsum_sym / dyadic / inv / det,
 *   Core::LinAlg::Matrix (multiply_nt, multiply), Core::LinAlg::Voigt::Stresses::invariants_principal,
 *   Core::FE::shape_function<celltype> / shape_function_deriv1<celltype>.
 * What cannot be included (needs Teuchos / Discretization / Global::Problem) is
 * glue and is restated here with the same expressions on the reference's types:
 *   src/solid_ele/4C_solid_ele_calc_lib.hpp:499-516,543-562,613-722,1107-1151,1254-1286
 *   src/mat/4C_mat_elasthyper_service.cpp:23-206,417-435, src/mat/elast/*, 4C_mat_stvenantkirchhoff.hpp
 * Built only where /root/reference exists (oracle/Makefile target `ref`); the prebuilt .so travels
 * to the GPU box.  Used to validate oracle/solid_oracle.cpp and as the "reference" CPU baseline.
 */
#include "4C_config.hpp"

#include "4C_fem_general_utils_fem_shapefunctions.hpp"
#include "4C_fem_general_utils_local_connectivity_matrices.hpp"
#include "4C_linalg_fixedsizematrix.hpp"
#include "4C_linalg_fixedsizematrix_voigt_notation.hpp"
#include "4C_linalg_symmetric_tensor.hpp"
#include "4C_linalg_tensor.hpp"
#include "4C_linalg_tensor_conversion.hpp"
#include "4C_linalg_tensor_einstein.hpp"
#include "4C_linalg_tensor_generators.hpp"
#include "4C_linalg_tensor_symmetric_einstein.hpp"

#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <source_location>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

// The real one lives in 4C_utils_exceptions.cpp (needs boost + MPI).
namespace FourC::Core::Internal
{
  [[noreturn]] void throw_error(const std::source_location&, const std::string& msg)
  {
    throw std::runtime_error(msg);
  }
}  // namespace FourC::Core::Internal

using namespace FourC;
namespace LA = FourC::Core::LinAlg;
using CT = Core::FE::CellType;

namespace
{
  template <CT celltype>
  constexpr int nn = Core::FE::num_nodes(celltype);

  template <CT celltype>
  struct ElementNodes
  {
    LA::Matrix<3, nn<celltype>> reference_coordinates;
    LA::Matrix<3, nn<celltype>> displacements;
  };

  template <CT celltype>
  struct ShapeFunctionsAndDerivatives
  {
    LA::Matrix<nn<celltype>, 1> shapefunctions_;
    LA::Matrix<3, nn<celltype>> derivatives_;
  };

  template <CT celltype>
  struct JacobianMapping
  {
    double determinant_;
    LA::Tensor<double, 3, 3> jacobian_;
    LA::Tensor<double, 3, 3> inverse_jacobian_;
    std::array<LA::Tensor<double, 3>, nn<celltype>> N_XYZ;
  };

  template <CT celltype>
  ShapeFunctionsAndDerivatives<celltype> evaluate_shape_functions_and_derivs(
      const LA::Tensor<double, 3>& xi)
  {
    ShapeFunctionsAndDerivatives<celltype> shapefcns;
    Core::FE::shape_function<celltype>(xi, shapefcns.shapefunctions_);
    Core::FE::shape_function_deriv1<celltype>(xi, shapefcns.derivatives_);
    return shapefcns;
  }

  template <CT celltype>
  JacobianMapping<celltype> evaluate_jacobian_mapping(
      const ShapeFunctionsAndDerivatives<celltype>& shapefcns, const ElementNodes<celltype>& nodes)
  {
    JacobianMapping<celltype> jacobian;
    LA::make_matrix_view(jacobian.jacobian_)
        .multiply_nt(shapefcns.derivatives_, nodes.reference_coordinates);
    jacobian.inverse_jacobian_ = LA::inv(jacobian.jacobian_);
    jacobian.determinant_ = LA::det(jacobian.jacobian_);
    LA::make_matrix_view(jacobian.N_XYZ)
        .multiply(LA::make_matrix_view(jacobian.inverse_jacobian_), shapefcns.derivatives_);
    return jacobian;
  }

  template <CT celltype>
  LA::Tensor<double, 3, 3> evaluate_deformation_gradient(
      const JacobianMapping<celltype>& jacobian_mapping, const ElementNodes<celltype>& nodes)
  {
    LA::Tensor<double, 3, 3> defgrd = LA::get_full(LA::TensorGenerators::identity<double, 3, 3>);
    LA::make_matrix_view(defgrd).multiply_nt(
        1.0, nodes.displacements, LA::make_matrix_view(jacobian_mapping.N_XYZ), 1.0);
    return defgrd;
  }

  // ---- materials on the reference's tensor types ------------------------------------------------
  void calculate_gamma_delta(LA::Matrix<3, 1>& gamma, LA::Matrix<8, 1>& delta,
      const LA::Matrix<3, 1>& prinv, const LA::Matrix<3, 1>& dPI, const LA::Matrix<6, 1>& ddPII)
  {
    gamma(0) = 2. * (dPI(0) + prinv(0) * dPI(1));
    gamma(1) = -2. * dPI(1);
    gamma(2) = 2. * prinv(2) * dPI(2);
    delta(0) = 4. * (ddPII(0) + 2. * prinv(0) * ddPII(5) + dPI(1) + prinv(0) * prinv(0) * ddPII(1));
    delta(1) = -4. * (ddPII(5) + prinv(0) * ddPII(1));
    delta(2) = 4. * (prinv(2) * ddPII(4) + prinv(0) * prinv(2) * ddPII(3));
    delta(3) = 4. * ddPII(1);
    delta(4) = -4. * prinv(2) * ddPII(3);
    delta(5) = 4. * (prinv(2) * dPI(2) + prinv(2) * prinv(2) * ddPII(2));
    delta(6) = -4. * prinv(2) * dPI(2);
    delta(7) = -4. * dPI(1);
  }

  void invariant_derivatives(int kind, const double* par, const LA::Matrix<3, 1>& prinv,
      LA::Matrix<3, 1>& dPI, LA::Matrix<6, 1>& ddPII)
  {
    dPI.clear();
    ddPII.clear();
    if (kind == 0)
    {
      const double c = par[0] / (4.0 * (1.0 + par[1]));
      const double beta = par[1] / (1.0 - 2.0 * par[1]);
      dPI(0) += c;
      if (prinv(2) > 0)
      {
        const double prinv2_to_beta_m1 = std::exp(std::log(prinv(2)) * (-beta - 1.));
        dPI(2) -= c * prinv2_to_beta_m1;
        ddPII(2) += c * (beta + 1.) * prinv2_to_beta_m1 / prinv(2);
      }
      else
        dPI(2) = ddPII(2) = std::numeric_limits<double>::quiet_NaN();
    }
    else
    {
      const double mue = par[0], kappa = par[1];
      LA::Matrix<3, 1> modinv(LA::Initialization::zero), dPmodI(LA::Initialization::zero);
      LA::Matrix<6, 1> ddPmodII(LA::Initialization::zero);
      modinv(0) = prinv(0) * std::pow(prinv(2), -1. / 3.);
      modinv(1) = prinv(1) * std::pow(prinv(2), -2. / 3.);
      modinv(2) = std::pow(prinv(2), 1. / 2.);
      dPmodI(0) += 0.5 * mue;
      dPmodI(2) += kappa * (modinv(2) - 1.);
      ddPmodII(2) += kappa;
      dPI(0) += std::pow(prinv(2), -1. / 3.) * dPmodI(0);
      dPI(1) += std::pow(prinv(2), -2. / 3.) * dPmodI(1);
      dPI(2) += 0.5 * std::pow(prinv(2), -0.5) * dPmodI(2) -
                1. / 3. * prinv(0) * std::pow(prinv(2), -4. / 3.) * dPmodI(0) -
                2. / 3. * prinv(1) * std::pow(prinv(2), -5. / 3.) * dPmodI(1);
      ddPII(0) += std::pow(prinv(2), -2. / 3.) * ddPmodII(0);
      ddPII(1) += std::pow(prinv(2), -4. / 3.) * ddPmodII(1);
      ddPII(2) += (1. / 9.) * std::pow(prinv(2), -8. / 3.) * prinv(0) * prinv(0) * ddPmodII(0) +
                  (4. / 9.) * prinv(0) * prinv(1) * std::pow(prinv(2), -3.) * ddPmodII(5) -
                  (1. / 3.) * std::pow(prinv(2), -11. / 6.) * prinv(0) * ddPmodII(4) +
                  (4. / 9.) * std::pow(prinv(2), -7. / 3.) * prinv(0) * dPmodI(0) +
                  (4. / 9.) * std::pow(prinv(2), -10. / 3.) * prinv(1) * prinv(1) * ddPmodII(1) -
                  (2. / 3.) * std::pow(prinv(2), -13. / 6.) * prinv(1) * ddPmodII(3) +
                  (10. / 9.) * std::pow(prinv(2), -8. / 3.) * prinv(1) * dPmodI(1) +
                  0.25 * std::pow(prinv(2), -1.) * ddPmodII(2) -
                  0.25 * std::pow(prinv(2), -1.5) * dPmodI(2);
      ddPII(3) += -(1. / 3.) * std::pow(prinv(2), -2.) * prinv(0) * ddPmodII(5) -
                  (2. / 3.) * std::pow(prinv(2), -7. / 3.) * prinv(1) * ddPmodII(1) +
                  0.5 * std::pow(prinv(2), -7. / 6.) * ddPmodII(3) -
                  (2. / 3.) * std::pow(prinv(2), -5. / 3.) * dPmodI(1);
      ddPII(4) += -(1. / 3.) * std::pow(prinv(2), -5. / 3.) * prinv(0) * ddPmodII(0) -
                  (2. / 3.) * std::pow(prinv(2), -2.) * prinv(1) * ddPmodII(5) +
                  0.5 * std::pow(prinv(2), -5. / 6.) * ddPmodII(4) -
                  (1. / 3.) * std::pow(prinv(2), -4. / 3.) * dPmodI(0);
      ddPII(5) += std::pow(prinv(2), -1.) * ddPmodII(5);
    }
  }

  void material_evaluate(int kind, const double* par, const LA::SymmetricTensor<double, 3, 3>& E,
      LA::SymmetricTensor<double, 3, 3>& stress, LA::SymmetricTensor<double, 3, 3, 3, 3>& cmat)
  {
    if (kind == 2)
    {
      const double Ey = par[0], nu = par[1];
      cmat = {};
      LA::Matrix<6, 6> cmat_view = LA::make_stress_like_voigt_view(cmat);
      const double mfac = Ey / ((1.0 + nu) * (1.0 - 2.0 * nu));
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cmat_view(i, j) = (i == j) ? mfac * (1.0 - nu) : mfac * nu;
      for (int i = 3; i < 6; ++i) cmat_view(i, i) = mfac * 0.5 * (1.0 - 2.0 * nu);
      const double mue = Ey / (2 * (1 + nu));
      const double lambda = Ey * nu / ((1 + nu) * (1 - 2 * nu));
      stress = lambda * LA::trace(E) * LA::TensorGenerators::identity<double, 3, 3> + 2 * mue * E;
      return;
    }
    LA::Matrix<3, 1> prinv, dPI;
    LA::Matrix<6, 1> ddPII;
    LA::SymmetricTensor<double, 3, 3> C_strain;
    C_strain = 2 * E + LA::TensorGenerators::identity<double, 3, 3>;
    LA::SymmetricTensor<double, 3, 3> iC = LA::inv(C_strain);
    LA::Voigt::Stresses::invariants_principal(prinv, LA::make_stress_like_voigt_view(C_strain));
    invariant_derivatives(kind, par, prinv, dPI, ddPII);
    stress = {};
    cmat = {};
    LA::Matrix<3, 1> gamma(LA::Initialization::zero);
    LA::Matrix<8, 1> delta(LA::Initialization::zero);
    calculate_gamma_delta(gamma, delta, prinv, dPI, ddPII);
    constexpr auto Id = LA::TensorGenerators::identity<double, 3, 3>;
    stress += gamma(0) * Id;
    stress += gamma(1) * C_strain;
    stress += gamma(2) * iC;
    cmat += delta(0) * LA::dyadic(Id, Id);
    cmat += delta(1) * (LA::dyadic(Id, C_strain) + LA::dyadic(C_strain, Id));
    cmat += delta(2) * (LA::dyadic(Id, iC) + LA::dyadic(iC, Id));
    cmat += delta(3) * LA::dyadic(C_strain, C_strain);
    cmat += delta(4) * (LA::dyadic(C_strain, iC) + LA::dyadic(iC, C_strain));
    cmat += delta(5) * LA::dyadic(iC, iC);
    cmat += delta(6) *
            ((0.5) * (LA::einsum_sym<"ik", "jl">(iC, iC) + LA::einsum_sym<"il", "jk">(iC, iC)));
    cmat += LA::TensorGenerators::symmetric_identity<double, 3, 3, 3, 3> * (delta(7));
  }

  template <CT celltype>
  void add_stiffness_matrix(const JacobianMapping<celltype>& jacobian_mapping,
      const LA::Tensor<double, 3, 3>& F, const LA::SymmetricTensor<double, 3, 3>& pk2,
      const LA::SymmetricTensor<double, 3, 3, 3, 3>& cmat, const double integration_fac,
      LA::Matrix<3 * nn<celltype>, 3 * nn<celltype>>& K)
  {
    const LA::Tensor<double, 3, 3, 3, 3> FCFTw =
        LA::einsum<"bi", "aijd", "cj">(F, cmat, F * integration_fac);
    for (int i = 0; i < nn<celltype>; ++i)
    {
      const auto dNidX_FCFTw = jacobian_mapping.N_XYZ[i] * FCFTw;
      const auto dNidX_PK2w = jacobian_mapping.N_XYZ[i] * pk2 * integration_fac;
      for (int j = 0; j < nn<celltype>; ++j)
      {
        const auto m = dNidX_FCFTw * jacobian_mapping.N_XYZ[j];
        for (int a = 0; a < 3; ++a)
          for (int b = 0; b < 3; ++b) K(3 * i + a, 3 * j + b) += m(a, b);
        const double g = dNidX_PK2w * jacobian_mapping.N_XYZ[j];
        for (int a = 0; a < 3; ++a) K(3 * i + a, 3 * j + a) += g;
      }
    }
  }

  template <CT celltype>
  ElementNodes<celltype> make_nodes(const double* X, const double* u)
  {
    ElementNodes<celltype> nodes;
    for (int i = 0; i < nn<celltype>; ++i)
      for (int d = 0; d < 3; ++d)
      {
        nodes.reference_coordinates(d, i) = X[3 * i + d];
        nodes.displacements(d, i) = u[3 * i + d];
      }
    return nodes;
  }

  template <CT celltype>
  int element(int matkind, const double* par, int action, const double* X, const double* u,
      const double* gxi, const double* gw, int ngp, double* Ke, double* fe)
  {
    constexpr int n = nn<celltype>;
    const ElementNodes<celltype> nodes = make_nodes<celltype>(X, u);
    LA::Matrix<3 * n, 3 * n> stiff(LA::Initialization::zero);
    LA::Matrix<3 * n, 1> force(LA::Initialization::zero);

    for (const auto& xi : Core::FE::get_element_nodes_in_parameter_space<celltype>())
    {
      LA::Tensor<double, 3> xi_t = LA::make_tensor<3>(xi);
      const auto jm = evaluate_jacobian_mapping(evaluate_shape_functions_and_derivs<celltype>(xi_t), nodes);
      if (!(jm.determinant_ > 0)) return 1;
    }
    for (int gp = 0; gp < ngp; ++gp)
    {
      const LA::Tensor<double, 3> xi{{gxi[3 * gp], gxi[3 * gp + 1], gxi[3 * gp + 2]}};
      const auto shape_functions = evaluate_shape_functions_and_derivs<celltype>(xi);
      const auto jm = evaluate_jacobian_mapping(shape_functions, nodes);
      const double integration_factor = jm.determinant_ * gw[gp];
      const LA::Tensor<double, 3, 3> F = evaluate_deformation_gradient(jm, nodes);
      const LA::SymmetricTensor<double, 3, 3> cauchygreen =
          LA::assume_symmetry(LA::transpose(F) * F);
      const LA::SymmetricTensor<double, 3, 3> gl_strain =
          0.5 * (cauchygreen - LA::TensorGenerators::identity<double, 3, 3>);
      LA::SymmetricTensor<double, 3, 3> pk2{};
      LA::SymmetricTensor<double, 3, 3, 3, 3> cmat{};
      material_evaluate(matkind, par, gl_strain, pk2, cmat);
      const auto PK1 = F * pk2;
      for (int i = 0; i < n; ++i)
      {
        const auto c = PK1 * jm.N_XYZ[i] * integration_factor;
        for (int j = 0; j < 3; ++j) force(3 * i + j) += c(j);
      }
      if (action == 1) add_stiffness_matrix<celltype>(jm, F, pk2, cmat, integration_factor, stiff);
    }
    std::memcpy(fe, force.data(), sizeof(double) * 3 * n);
    if (Ke && action == 1) std::memcpy(Ke, stiff.data(), sizeof(double) * 9 * n * n);
    return 0;
  }
}  // namespace

extern "C"
{
  void ref_shape(int ct, const double* xi_, double* N, double* dN)
  {
    const LA::Tensor<double, 3> xi{{xi_[0], xi_[1], xi_[2]}};
    auto run = [&]<CT celltype>()
    {
      const auto s = evaluate_shape_functions_and_derivs<celltype>(xi);
      std::memcpy(N, s.shapefunctions_.data(), sizeof(double) * nn<celltype>);
      std::memcpy(dN, s.derivatives_.data(), sizeof(double) * 3 * nn<celltype>);
    };
    if (ct == 0)
      run.template operator()<CT::hex8>();
    else if (ct == 1)
      run.template operator()<CT::hex27>();
    else if (ct == 3)
      run.template operator()<CT::tet4>();
    else if (ct == 4)
      run.template operator()<CT::hex20>();
    else if (ct == 5)
      run.template operator()<CT::wedge6>();
    else if (ct == 6)
      run.template operator()<CT::pyramid5>();
    else if (ct == 7)
      run.template operator()<CT::hex18>();
    else
      run.template operator()<CT::tet10>();
  }

  void ref_node_param_coords(int ct, double* out)
  {
    auto run = [&]<CT celltype>()
    {
      int k = 0;
      for (const auto& xi : Core::FE::get_element_nodes_in_parameter_space<celltype>())
      {
        for (int d = 0; d < 3; ++d) out[3 * k + d] = xi[d];
        ++k;
      }
    };
    if (ct == 0)
      run.template operator()<CT::hex8>();
    else if (ct == 1)
      run.template operator()<CT::hex27>();
    else if (ct == 3)
      run.template operator()<CT::tet4>();
    else if (ct == 4)
      run.template operator()<CT::hex20>();
    else if (ct == 5)
      run.template operator()<CT::wedge6>();
    else if (ct == 6)
      run.template operator()<CT::pyramid5>();
    else if (ct == 7)
      run.template operator()<CT::hex18>();
    else
      run.template operator()<CT::tet10>();
  }

  void ref_material(int kind, const double* par, const double* E_, double* S, double* C)
  {
    LA::SymmetricTensor<double, 3, 3> E{}, pk2{};
    LA::SymmetricTensor<double, 3, 3, 3, 3> cmat{};
    std::memcpy(E.data(), E_, 6 * sizeof(double));
    material_evaluate(kind, par, E, pk2, cmat);
    std::memcpy(S, pk2.data(), 6 * sizeof(double));
    std::memcpy(C, cmat.data(), 36 * sizeof(double));
  }

  int ref_element(int ct, int matkind, const double* par, int action, const double* X,
      const double* u, const double* gxi, const double* gw, int ngp, double* Ke, double* fe)
  {
    if (ct == 0) return element<CT::hex8>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 1) return element<CT::hex27>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 3) return element<CT::tet4>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 4) return element<CT::hex20>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 5) return element<CT::wedge6>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 6) return element<CT::pyramid5>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    if (ct == 7) return element<CT::hex18>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
    return element<CT::tet10>(matkind, par, action, X, u, gxi, gw, ngp, Ke, fe);
  }

  // Whole-mesh evaluate + owner-computes assembly with the element arithmetic done by the
  // reference's headers (same contract as orc_evaluate_assemble in solid_oracle.h; the loop and the
  // assembly restate 4C_fem_discretization_evaluate.cpp:85-117 and 4C_linalg_sparsematrix.cpp:356-467).
  // This is the "reference" CPU arm of bench.py: one thread per emulated MPI rank.
  int64_t ref_evaluate_assemble(int ct, int matkind, const double* matparams,
      const double* mue_per_ele, int action, int64_t n_ele, const int32_t* conn,
      const double* X, const double* u_col, const int32_t* node_row_lid, const int32_t* node_col_lid,
      int64_t n_row_nodes, const int64_t* rowptr, const int32_t* colind, const double* gxi,
      const double* gw, int ngp, double* values, double* f, int nthreads)
  {
    const int n = ct == 0 ? 8 : (ct == 1 ? 27 : 10), nd = 3 * n;
    if (nthreads < 1) nthreads = 1;
    std::vector<int64_t> bad(nthreads, 0);
    auto worker = [&](int tid)
    {
      const int64_t r0 = n_row_nodes * tid / nthreads, r1 = n_row_nodes * (tid + 1) / nthreads;
      const bool stiff = action == 1 && values;
      if (stiff) std::fill(values + rowptr[3 * r0], values + rowptr[3 * r1], 0.0);
      std::fill(f + 3 * r0, f + 3 * r1, 0.0);
      std::vector<double> Ke(nd * nd), fe(nd), Xe(nd), ue(nd);
      std::vector<int32_t> localcol(nd);
      double par[2] = {matparams[0], matparams[1]};
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
            ue[3 * i + d] = u_col[3 * (int64_t)ce[i] + d];
          }
        if (mue_per_ele) par[0] = mue_per_ele[e];
        if (ref_element(ct, matkind, par, action, Xe.data(), ue.data(), gxi, gw, ngp,
                stiff ? Ke.data() : nullptr, fe.data()))
        {
          bad[tid]++;
          continue;
        }
        for (int j = 0; j < n; ++j)
          for (int c = 0; c < 3; ++c) localcol[3 * j + c] = 3 * node_col_lid[ce[j]] + c;
        for (int i = 0; i < n; ++i)
        {
          const int32_t r = node_row_lid[ce[i]];
          if (r < r0 || r >= r1) continue;
          for (int b = 0; b < 3; ++b)
          {
            const int lrow = 3 * i + b;
            const int64_t rlid = 3 * (int64_t)r + b;
            f[rlid] += fe[lrow];
            if (!stiff) continue;
            const int length = (int)(rowptr[rlid + 1] - rowptr[rlid]);
            double* valview = values + rowptr[rlid];
            const int32_t* indices = colind + rowptr[rlid];
            int dofcount = 0, pos = 0;
            for (int node = 0; node < n; ++node)
            {
              if (pos >= length || indices[pos] != localcol[dofcount])
                pos = (int)(std::lower_bound(indices, indices + length, localcol[dofcount]) - indices);
              for (int j = 0; j < 3; ++j) valview[pos++] += Ke[lrow + nd * (dofcount++)];
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

  void ref_stiffness_bfree(int ct, const double* X, const double* u, const double* xi_,
      const double* S_, const double* C_, double fac, double* Ke)
  {
    const LA::Tensor<double, 3> xi{{xi_[0], xi_[1], xi_[2]}};
    LA::SymmetricTensor<double, 3, 3> pk2{};
    LA::SymmetricTensor<double, 3, 3, 3, 3> cmat{};
    std::memcpy(pk2.data(), S_, 6 * sizeof(double));
    std::memcpy(cmat.data(), C_, 36 * sizeof(double));
    auto run = [&]<CT celltype>()
    {
      constexpr int n = nn<celltype>;
      const auto nodes = make_nodes<celltype>(X, u);
      const auto jm =
          evaluate_jacobian_mapping(evaluate_shape_functions_and_derivs<celltype>(xi), nodes);
      const auto F = evaluate_deformation_gradient(jm, nodes);
      LA::Matrix<3 * n, 3 * n> K(LA::Initialization::zero);
      add_stiffness_matrix<celltype>(jm, F, pk2, cmat, fac, K);
      for (int i = 0; i < 9 * n * n; ++i) Ke[i] += K.data()[i];
    };
    if (ct == 0)
      run.template operator()<CT::hex8>();
    else if (ct == 1)
      run.template operator()<CT::hex27>();
    else if (ct == 3)
      run.template operator()<CT::tet4>();
    else if (ct == 4)
      run.template operator()<CT::hex20>();
    else if (ct == 5)
      run.template operator()<CT::wedge6>();
    else if (ct == 6)
      run.template operator()<CT::pyramid5>();
    else if (ct == 7)
      run.template operator()<CT::hex18>();
    else
      run.template operator()<CT::tet10>();
  }

  // cmat used by the reference's BFreeElementStiffnessMatrix unit test
  // (unittests/solid_ele/4C_solid_ele_calc_lib_test.cpp:282-294)
  void ref_test_cmat_from_pk2(const double* S_, double* C)
  {
    LA::SymmetricTensor<double, 3, 3> pk2{};
    std::memcpy(pk2.data(), S_, 6 * sizeof(double));
    LA::SymmetricTensor<double, 3, 3, 3, 3> cmat =
        LA::einsum_sym<"AC", "BD">(pk2, pk2) + LA::einsum_sym<"AD", "BC">(pk2, pk2);
    std::memcpy(C, cmat.data(), 36 * sizeof(double));
  }
}

// host_check.cpp -- links the C++ adapter against lib4c_b200_solid.so and runs a one-element hex8
// evaluate when a GPU is present (tests/test_host_adapter.py).  Without a GPU it verifies that the
// adapter fails loudly (no CPU fallback) and exits 0.
#include <cmath>
#include <cstdio>
#include <vector>

#include "b200_solid_adapter.hpp"

int main()
{
  using namespace B200Solid;
  try
  {
    DeviceElementBlock blk(0, 0, 1);
    // unit cube, node order of tests/input_files/solid_ele_hex8_Standard_eas_none_stressout.4C.yaml
    const double X[24] = {0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1};
    const std::int32_t conn[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    std::vector<std::int32_t> lid(24);
    for (int i = 0; i < 24; ++i) lid[i] = i;
    ElementBlockView v;
    v.celltype = B200S_HEX8, v.n_col_nodes = 8, v.X = X, v.n_col_eles = 1, v.conn = conn;
    v.dof_row_lid = lid.data(), v.dof_col_lid = lid.data();
    blk.set_block(v);
    blk.set_material(B200S_MAT_COUP_NEOHOOKE, {10.0, 0.25});
    // dense 24x24 graph = what FillComplete gives for one element
    std::vector<int> off(25), ind(576);
    for (int r = 0; r <= 24; ++r) off[r] = 24 * r;
    for (int r = 0; r < 24; ++r)
      for (int c = 0; c < 24; ++c) ind[24 * r + c] = c;
    std::vector<double> K(576, -1.0), f(24, -1.0), u(24, 0.0);
    CrsView crs;
    crs.n_rows = 24, crs.index_offset = off.data(), crs.indices = ind.data(), crs.values = K.data();
    blk.set_graph(v, crs);
    blk.set_state(u.data());
    blk.evaluate(Action::struct_calc_nlnstiff, K.data(), f.data());
    // undeformed: zero internal force, symmetric K with zero row sums per direction
    double fmax = 0, asym = 0, rowsum = 0, kmax = 0;
    for (double x : f) fmax = std::fmax(fmax, std::fabs(x));
    for (int r = 0; r < 24; ++r)
    {
      double s[3] = {0, 0, 0};
      for (int c = 0; c < 24; ++c)
      {
        asym = std::fmax(asym, std::fabs(K[24 * r + c] - K[24 * c + r]));
        kmax = std::fmax(kmax, std::fabs(K[24 * r + c]));
        s[c % 3] += K[24 * r + c];
      }
      for (double x : s) rowsum = std::fmax(rowsum, std::fabs(x));
    }
    std::printf("HOST_ADAPTER_OK fmax=%.3g asym=%.3g rowsum=%.3g kmax=%.3g\n", fmax, asym, rowsum, kmax);
    const bool ok = fmax < 1e-13 && asym < 1e-13 * kmax && rowsum < 1e-12 * kmax && kmax > 0.1;
    // two element blocks (two MAT ids in the reference's terms) assembling into one matrix give the same K as one
    // block holding both elements (same material in both blocks here)
    bool blocks_ok = false;
    {
      const double X2[36] = {0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 1, 1, 1, 1, 0, 1, 1,
          0, 0, 2.1, 1, 0, 2, 1.1, 1, 2, 0, 1, 2.2};
      const std::int32_t conn2[16] = {0, 1, 2, 3, 4, 5, 6, 7, 4, 5, 6, 7, 8, 9, 10, 11};
      std::vector<std::int32_t> lid2(36);
      for (int i = 0; i < 36; ++i) lid2[i] = i;
      // graph of the two-element mesh: nodes 0-3 see 0-7, nodes 4-7 see all, nodes 8-11 see 4-11
      std::vector<int> off2(37, 0), ind2;
      for (int r = 0; r < 36; ++r)
      {
        const int node = r / 3, lo = node < 4 ? 0 : (node < 8 ? 0 : 12), hi = node < 4 ? 24 : 36;
        for (int c = lo; c < hi; ++c) ind2.push_back(c);
        off2[r + 1] = (int)ind2.size();
      }
      std::vector<double> u2(36), Ka(ind2.size()), Kb(ind2.size()), fa(36), fb(36);
      for (int i = 0; i < 36; ++i) u2[i] = 0.01 * std::sin(1.7 * i);
      CrsView crs2;
      crs2.n_rows = 36, crs2.index_offset = off2.data(), crs2.indices = ind2.data();
      ElementBlockView all, lower, upper;
      all.celltype = B200S_HEX8, all.n_col_nodes = 12, all.X = X2, all.n_col_eles = 2, all.conn = conn2;
      all.dof_row_lid = lid2.data(), all.dof_col_lid = lid2.data();
      lower = all, upper = all;
      lower.n_col_eles = 1, upper.n_col_eles = 1, upper.conn = conn2 + 8;
      DeviceElementBlock one(0, 0, 1), pa(0, 0, 1), pb(0, 0, 1);
      one.set_block(all), one.set_material(B200S_MAT_COUP_NEOHOOKE, {10.0, 0.25}), one.set_graph(all, crs2);
      one.set_state(u2.data());
      one.evaluate(Action::struct_calc_nlnstiff, Ka.data(), fa.data());
      pa.set_block(lower), pa.set_material(B200S_MAT_COUP_NEOHOOKE, {10.0, 0.25}), pa.set_graph(all, crs2);
      pb.set_block(upper), pb.set_material(B200S_MAT_COUP_NEOHOOKE, {10.0, 0.25}), pb.attach_to(pa);
      pa.set_state(u2.data());
      pa.evaluate(Action::struct_calc_nlnstiff, nullptr, nullptr);
      pb.evaluate(Action::struct_calc_nlnstiff, Kb.data(), fb.data());
      double dk = 0, df = 0, km = 0;
      for (size_t i = 0; i < Ka.size(); ++i) dk = std::fmax(dk, std::fabs(Ka[i] - Kb[i])), km = std::fmax(km, std::fabs(Ka[i]));
      for (int i = 0; i < 36; ++i) df = std::fmax(df, std::fabs(fa[i] - fb[i]));
      blocks_ok = dk < 1e-13 * km && df < 1e-13 * km && km > 0.1;
      std::printf("HOST_ADAPTER_BLOCKS dk=%.3g df=%.3g kmax=%.3g\n", dk, df, km);
      if (!DeviceElementBlock::supported(B200S_HEX8, B200S_KINEM_LINEAR, B200S_TECH_NONE, B200S_PRESTRESS_NONE,
              B200S_MAT_DRUCKER_PRAGER) ||
          DeviceElementBlock::supported(B200S_TET10, B200S_KINEM_NONLINEAR_TOTLAG, B200S_TECH_FBAR, B200S_PRESTRESS_NONE,
              B200S_MAT_SVK))
        blocks_ok = false;
    }
    // inverted element: the reference throws
    double Xbad[24];
    for (int i = 0; i < 24; ++i) Xbad[i] = X[i];
    Xbad[18] = -1.0, Xbad[19] = -1.0, Xbad[20] = -1.0;
    v.X = Xbad;
    blk.set_block(v);
    blk.set_graph(v, crs);
    bool thrown = false;
    try
    {
      blk.evaluate(Action::struct_calc_nlnstiff, K.data(), f.data());
    }
    catch (const Exception& e)
    {
      thrown = e.code() == B200S_ERR_NEG_JACOBIAN;
      std::printf("expected throw: %s\n", e.what());
    }
    return ok && thrown && blocks_ok ? 0 : 1;
  }
  catch (const Exception& e)
  {
    if (e.code() == B200S_ERR_CUDA)
    {
      std::printf("HOST_ADAPTER_NO_GPU %s\n", e.what());
      return 0;
    }
    std::printf("unexpected: %s\n", e.what());
    return 1;
  }
}

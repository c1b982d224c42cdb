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
    return ok && thrown ? 0 : 1;
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

// Shim that shadows the Teuchos-based SerialDenseMatrix header so that the reference's
// header-only shape-function file can be included without Trilinos. Test infrastructure only.
#ifndef FOUR_C_LINALG_SERIALDENSEMATRIX_HPP
#define FOUR_C_LINALG_SERIALDENSEMATRIX_HPP
#include "4C_config.hpp"
#include <iostream>
#include <vector>
FOUR_C_NAMESPACE_OPEN
namespace Core::LinAlg
{
  class SerialDenseMatrix
  {
  };
  class SerialDenseVector
  {
  };
}  // namespace Core::LinAlg
FOUR_C_NAMESPACE_CLOSE
#endif

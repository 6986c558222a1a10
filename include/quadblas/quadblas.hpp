// quadblas/quadblas.hpp — umbrella header of the B200 build of QuadBLAS.
//
// Same include path and public names as the reference's umbrella
// (/root/reference/include/quadblas/quadblas.hpp:27-59), so `#include <quadblas/quadblas.hpp>`
// (README:64,104) keeps compiling; link with -lqblas_b200.  Every routine runs on the GPU through
// the C ABI in include/qblas_b200.h — there is no host implementation behind these headers.
#ifndef QUADBLAS_HPP
#define QUADBLAS_HPP

#include "core/platform.hpp"
#include "core/constants.hpp"
#include "core/types.hpp"
#include "memory/allocation.hpp"
#include "simd/quad_vector.hpp"
#include "threading/openmp_utils.hpp"
#include "detail/blocking.hpp"
#include "algorithms/level1.hpp"
#include "algorithms/level2.hpp"
#include "algorithms/level3.hpp"
#include "interface/c_interface.hpp"
#include "interface/cpp_classes.hpp"

namespace QuadBLAS
{
  constexpr const char *VERSION = "1.0.0";
  constexpr int VERSION_MAJOR = 1;
  constexpr int VERSION_MINOR = 0;
  constexpr int VERSION_PATCH = 0;
}

#endif // QUADBLAS_HPP

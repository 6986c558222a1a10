// quadblas/threading/openmp_utils.hpp (B200 build) — get/set_num_threads
// (/root/reference/include/quadblas/threading/openmp_utils.hpp:10-24).  There is no OpenMP in the
// CUDA library; the value is kept because it is part of the numerical contract: in reference-order
// mode dot/nrm2 cut [0,n) into that many chunks (level1.hpp:46-65).
#ifndef QUADBLAS_THREADING_OPENMP_UTILS_HPP
#define QUADBLAS_THREADING_OPENMP_UTILS_HPP
#include "../core/platform.hpp"
namespace QuadBLAS
{
  inline int get_num_threads() { return quadblas_get_num_threads(); }
  inline void set_num_threads(int num_threads) { quadblas_set_num_threads(num_threads); }
} // namespace QuadBLAS
#endif // QUADBLAS_THREADING_OPENMP_UTILS_HPP

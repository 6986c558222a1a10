// quadblas/core/constants.hpp (B200 build).  Names and values of the reference's public constants
// (/root/reference/include/quadblas/core/constants.hpp:10-18) are kept because user code prints and
// tests them (benchmarks/benchmark.cpp:411-412); on the GPU only two keep a meaning.
#ifndef QUADBLAS_CORE_CONSTANTS_HPP
#define QUADBLAS_CORE_CONSTANTS_HPP
#include <cstddef>
namespace QuadBLAS
{
  constexpr size_t VECTOR_SIZE = 2;          // lanes of the reference's dot kernel: part of reference-order mode (two chains)
  constexpr size_t PARALLEL_THRESHOLD = 500; // n >= 500: the reference chunks dot by thread count (level1.hpp:40-46); the device kernel follows it
  constexpr size_t ALIGNMENT = 32;           // quadblas_is_aligned / aligned_alloc contract
  constexpr size_t CACHE_LINE_SIZE = 64;     // host cache geometry below: unused by the CUDA library,
  constexpr size_t L1_CACHE_SIZE = 32768;    //   kept so BlockingParams reproduces the reference's kc = 126
  constexpr size_t L2_CACHE_SIZE = 262144;
  constexpr size_t GEMM_BLOCK_SIZE = 64;
} // namespace QuadBLAS
#endif // QUADBLAS_CORE_CONSTANTS_HPP

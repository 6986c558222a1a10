// quadblas/algorithms/level3.hpp (B200 build) — QuadBLAS::gemm
// (/root/reference/include/quadblas/algorithms/level3.hpp:215-336).  C <- alpha A B + beta C with
// A m x k, B k x n, C m x n in the given layout; no transposes at this level (the reference's C
// entry point drops transa/transb, c_interface.hpp:109-112).  Forwards to qb_gemm: the
// reference-order integer-limb kernel (csrc/qb_level3.cu) or, in fast mode, the tensor-core path
// (csrc/qb_ozaki.cu).
#ifndef QUADBLAS_ALGORITHMS_LEVEL3_HPP
#define QUADBLAS_ALGORITHMS_LEVEL3_HPP
#include "level1.hpp"
#include "../core/types.hpp"
#include "../detail/blocking.hpp"
namespace QuadBLAS
{
  constexpr size_t GEMM_MR = 4; // the reference's register tile (level3.hpp:17-18); no effect on results
  constexpr size_t GEMM_NR = 4;

  inline void gemm(Layout layout, size_t m, size_t n, size_t k, Sleef_quad alpha, const Sleef_quad *A, size_t lda,
                   const Sleef_quad *B, size_t ldb, Sleef_quad beta, Sleef_quad *C, size_t ldc)
  {
    const qb_quad a = b200::bits(alpha), b = b200::bits(beta);
    qb_gemm(b200::layout_char(layout), 'N', 'N', (int64_t)m, (int64_t)n, (int64_t)k, &a, A, (int64_t)lda, B, (int64_t)ldb, &b, C, (int64_t)ldc);
  }
} // namespace QuadBLAS
#endif // QUADBLAS_ALGORITHMS_LEVEL3_HPP

// quadblas/algorithms/level2.hpp (B200 build) — QuadBLAS::gemv
// (/root/reference/include/quadblas/algorithms/level2.hpp:85-99; row-major :15-50, col-major :53-82).
// Forwards to qb_gemv (csrc/qb_level2.cu); m, n are the dimensions AFTER any transpose relabelling,
// exactly like the reference's free function.
#ifndef QUADBLAS_ALGORITHMS_LEVEL2_HPP
#define QUADBLAS_ALGORITHMS_LEVEL2_HPP
#include "level1.hpp"
#include "../core/types.hpp"
namespace QuadBLAS
{
  inline void gemv(Layout layout, size_t m, size_t n, Sleef_quad alpha, const Sleef_quad *A, size_t lda, const Sleef_quad *x,
                   size_t incx, Sleef_quad beta, Sleef_quad *y, size_t incy)
  {
    const qb_quad a = b200::bits(alpha), b = b200::bits(beta);
    qb_gemv(b200::layout_char(layout), (int64_t)m, (int64_t)n, &a, A, (int64_t)lda, x, (int64_t)incx, &b, y, (int64_t)incy);
  }
} // namespace QuadBLAS
#endif // QUADBLAS_ALGORITHMS_LEVEL2_HPP

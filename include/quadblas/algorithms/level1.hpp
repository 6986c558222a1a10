// quadblas/algorithms/level1.hpp (B200 build) — QuadBLAS::dot, QuadBLAS::dot_kernel_vectorized and QuadBLAS::axpy
// (/root/reference/include/quadblas/algorithms/level1.hpp:14-35, :80-137 and :190-223).
// Same signatures; the work is done by the CUDA kernels in csrc/qb_level1.cu through qb_dot /
// qb_axpy.  Pointers may be host or device memory (classified by the library).  On a CUDA failure
// dot returns NaN and the library's sticky error (qb_last_error) says why; nothing is computed on
// the host.
#ifndef QUADBLAS_ALGORITHMS_LEVEL1_HPP
#define QUADBLAS_ALGORITHMS_LEVEL1_HPP
#include "../core/platform.hpp"
#include <cstddef>
#include <cstring>
namespace QuadBLAS
{
  namespace b200
  {
    inline qb_quad bits(Sleef_quad v) { qb_quad r; std::memcpy(&r, &v, 16); return r; }
    inline Sleef_quad quad(qb_quad v) { Sleef_quad r; std::memcpy(&r, &v, 16); return r; }
    inline Sleef_quad nan_quad() { qb_quad r; r.lo = 0; r.hi = 0x7fff800000000000ULL; return quad(r); }
  }

  // level1.hpp:14-35 — the two-lane kernel over contiguous data (even / odd chains from +0, add(lane0, lane1), odd tail), in the
  // reference's order whatever the library mode; the reference's debug programs call it directly (tests/debug_test.cpp:79).
  inline Sleef_quad dot_kernel_vectorized(const Sleef_quad *x, const Sleef_quad *y, size_t n)
  {
    qb_quad r;
    if (qb_dot_kernel((int64_t)n, x, y, &r) != QB_OK) return b200::nan_quad();
    return b200::quad(r);
  }

  inline Sleef_quad dot(size_t n, const Sleef_quad *x, size_t incx, const Sleef_quad *y, size_t incy)
  {
    qb_quad r;
    if (qb_dot((int64_t)n, x, (int64_t)incx, y, (int64_t)incy, &r) != QB_OK) return b200::nan_quad();
    return b200::quad(r);
  }

  inline void axpy(size_t n, Sleef_quad alpha, const Sleef_quad *x, size_t incx, Sleef_quad *y, size_t incy)
  {
    const qb_quad a = b200::bits(alpha);
    qb_axpy((int64_t)n, &a, x, (int64_t)incx, y, (int64_t)incy);
  }
} // namespace QuadBLAS
#endif // QUADBLAS_ALGORITHMS_LEVEL1_HPP

// quadblas/simd/quad_vector.hpp (B200 build) — the reference's two-lane value type
// (/root/reference/include/quadblas/simd/quad_vector.hpp:11-151).  Only the reference's debug
// programs touch it (tests/debug_test.cpp:31-48); it is a host convenience over two Sleef_quad
// lanes and plays no part in the device kernels.
#ifndef QUADBLAS_SIMD_QUAD_VECTOR_HPP
#define QUADBLAS_SIMD_QUAD_VECTOR_HPP
#include "../core/platform.hpp"
namespace QuadBLAS
{
  class QuadVector
  {
    Sleef_quad lane_[2];

  public:
    QuadVector() = default;
    explicit QuadVector(Sleef_quad v) : lane_{v, v} {}
    QuadVector(Sleef_quad a, Sleef_quad b) : lane_{a, b} {}
    static QuadVector load(const Sleef_quad *p) { return QuadVector(p[0], p[1]); }
    void store(Sleef_quad *p) const { p[0] = lane_[0]; p[1] = lane_[1]; }
    Sleef_quad get(int i) const { return lane_[i]; }
    QuadVector operator+(const QuadVector &o) const { return QuadVector(Sleef_addq1_u05(lane_[0], o.lane_[0]), Sleef_addq1_u05(lane_[1], o.lane_[1])); }
    QuadVector operator*(const QuadVector &o) const { return QuadVector(Sleef_mulq1_u05(lane_[0], o.lane_[0]), Sleef_mulq1_u05(lane_[1], o.lane_[1])); }
    // this * b + c, one rounding per lane
    QuadVector fma(const QuadVector &b, const QuadVector &c) const
    {
      return QuadVector(Sleef_fmaq1_u05(lane_[0], b.lane_[0], c.lane_[0]), Sleef_fmaq1_u05(lane_[1], b.lane_[1], c.lane_[1]));
    }
    Sleef_quad horizontal_sum() const { return Sleef_addq1_u05(lane_[0], lane_[1]); }
  };
} // namespace QuadBLAS
#endif // QUADBLAS_SIMD_QUAD_VECTOR_HPP

// quadblas/interface/cpp_classes.hpp (B200 build) — QuadBLAS::Vector<Layout>, Matrix<Layout> and
// their aliases, API-compatible with /root/reference/include/quadblas/interface/cpp_classes.hpp:
// Vector :18-82 (owning / view constructors, operator[], size/stride/data, dot, axpy, norm),
// Matrix :85-155 (operator(), rows/cols/leading_dimension/data, gemv, gemm), aliases :158-168.
// Owning containers hold page-locked host memory (zero filled), so element access stays plain host
// access and every routine call stages through the CUDA library at full copy rate.  Quirk kept on
// purpose: an owning Matrix uses ld = cols for BOTH layouts (cpp_classes.hpp:96), so a non-square
// owning ColMajor matrix is only well-formed when rows <= cols, as in the reference.
#ifndef QUADBLAS_INTERFACE_CPP_CLASSES_HPP
#define QUADBLAS_INTERFACE_CPP_CLASSES_HPP
#include "../algorithms/level1.hpp"
#include "../algorithms/level2.hpp"
#include "../algorithms/level3.hpp"
#include "../core/types.hpp"
#include "../memory/allocation.hpp"
#include <cstddef>
#include <cstring>
namespace QuadBLAS
{
  namespace b200
  {
    // Move-only owner-or-view of a run of quads; shared by Vector and Matrix.
    class QuadSpan
    {
      Sleef_quad *p_ = nullptr;
      bool owned_ = false;

    public:
      explicit QuadSpan(size_t count) : p_(aligned_alloc<Sleef_quad>(count)), owned_(true)
      {
        if (p_ != nullptr) std::memset(static_cast<void *>(p_), 0, count * sizeof(Sleef_quad)); // +0.0 is all-zero bits
      }
      explicit QuadSpan(Sleef_quad *view) : p_(view) {}
      QuadSpan(QuadSpan &&o) noexcept : p_(o.p_), owned_(o.owned_) { o.owned_ = false; }
      QuadSpan(const QuadSpan &) = delete;
      QuadSpan &operator=(const QuadSpan &) = delete;
      ~QuadSpan() { if (owned_) aligned_free(p_); }
      Sleef_quad *get() const { return p_; }
    };
  }

  template <Layout layout>
  class Vector
  {
    b200::QuadSpan mem_;
    size_t n_, inc_;

  public:
    explicit Vector(size_t size) : mem_(size), n_(size), inc_(1) {}
    Vector(Sleef_quad *data, size_t size, size_t stride = 1) : mem_(data), n_(size), inc_(stride) {}
    Vector(Vector &&) noexcept = default;

    Sleef_quad &operator[](size_t i) { return mem_.get()[i * inc_]; }
    const Sleef_quad &operator[](size_t i) const { return mem_.get()[i * inc_]; }
    size_t size() const { return n_; }
    size_t stride() const { return inc_; }
    Sleef_quad *data() { return mem_.get(); }
    const Sleef_quad *data() const { return mem_.get(); }

    Sleef_quad dot(const Vector &other) const { return QuadBLAS::dot(n_, data(), inc_, other.data(), other.inc_); }
    // this <- alpha * other + this
    void axpy(Sleef_quad alpha, const Vector &other) { QuadBLAS::axpy(n_, alpha, other.data(), other.inc_, data(), inc_); }
    // sqrt(dot(x, x)) with the square root taken on the device too (qb_nrm2), full quad result
    Sleef_quad norm() const
    {
      qb_quad r;
      if (qb_nrm2((int64_t)n_, data(), (int64_t)inc_, &r) != QB_OK) return b200::nan_quad();
      return b200::quad(r);
    }
  };

  template <Layout layout>
  class Matrix
  {
    b200::QuadSpan mem_;
    size_t m_, n_, ld_;
    size_t at(size_t i, size_t j) const { return layout == Layout::RowMajor ? i * ld_ + j : j * ld_ + i; }

  public:
    Matrix(size_t rows, size_t cols) : mem_(rows * cols), m_(rows), n_(cols), ld_(cols) {}
    Matrix(Sleef_quad *data, size_t rows, size_t cols, size_t ld = 0) : mem_(data), m_(rows), n_(cols), ld_(ld ? ld : cols) {}
    Matrix(Matrix &&) noexcept = default;

    Sleef_quad &operator()(size_t i, size_t j) { return mem_.get()[at(i, j)]; }
    const Sleef_quad &operator()(size_t i, size_t j) const { return mem_.get()[at(i, j)]; }
    size_t rows() const { return m_; }
    size_t cols() const { return n_; }
    size_t leading_dimension() const { return ld_; }
    Sleef_quad *data() { return mem_.get(); }
    const Sleef_quad *data() const { return mem_.get(); }

    // y <- alpha * this * x + beta * y
    void gemv(Sleef_quad alpha, const Vector<layout> &x, Sleef_quad beta, Vector<layout> &y) const
    {
      QuadBLAS::gemv(layout, m_, n_, alpha, data(), ld_, x.data(), x.stride(), beta, y.data(), y.stride());
    }
    // C <- alpha * this * B + beta * C
    void gemm(Sleef_quad alpha, const Matrix &B, Sleef_quad beta, Matrix &C) const
    {
      QuadBLAS::gemm(layout, m_, B.cols(), n_, alpha, data(), ld_, B.data(), B.leading_dimension(), beta, C.data(), C.leading_dimension());
    }
  };

  using VectorRowMajor = Vector<Layout::RowMajor>;
  using VectorColMajor = Vector<Layout::ColMajor>;
  using MatrixRowMajor = Matrix<Layout::RowMajor>;
  using MatrixColMajor = Matrix<Layout::ColMajor>;
  template <Layout layout = Layout::RowMajor> using DefaultVector = Vector<layout>;
  template <Layout layout = Layout::RowMajor> using DefaultMatrix = Matrix<layout>;
} // namespace QuadBLAS
#endif // QUADBLAS_INTERFACE_CPP_CLASSES_HPP

// TEST: the drop-in C++ surface (include/quadblas/quadblas.hpp) and the reference-named C ABI against
// libqblas_b200.so, on a GPU.  Cases follow the reference's own known-answer tests
// (/root/reference/test_quadblas.cpp:205-224 dot 1..10, :290-312 norm, :327-350 3x3 gemv, :446-464 2x2
// gemm, :691-712 identity, :715-739 cancellation; tests/debug_test.cpp:34-47 QuadVector) plus naive
// single-chain loops built from the caller-side scalar ops (exact for small-integer data in any order).
#include <quadblas/quadblas.hpp>
#include <cstdio>
#include <cstring>
#include <vector>

static int fails = 0;
static bool same(Sleef_quad a, Sleef_quad b) { return std::memcmp(&a, &b, 16) == 0; }
#define CHECK(c) do { if (!(c)) { ++fails; std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); } } while (0)
static Sleef_quad Q(double d) { return Sleef_cast_from_doubleq1(d); }

int main()
{
  using namespace QuadBLAS;
  if (qb_init() != QB_OK) { std::printf("no usable CUDA device: %s\n", qb_last_error()); return 2; }
  std::printf("%s | %s | threads=%d\n", quadblas_get_version(), qb_build_info(), get_num_threads());
  CHECK(std::strcmp(quadblas_get_version(), "QuadBLAS 1.0.0 - High Performance Quad Precision BLAS") == 0);
  CHECK(std::strcmp(VERSION, "1.0.0") == 0 && ALIGNMENT == 32 && VECTOR_SIZE == 2);

  { // dot(1..10, ones) = 55, through Vector and through the C ABI (which rounds to double)
    DefaultVector<> x(10), y(10);
    for (int i = 0; i < 10; ++i) { x[i] = Q(i + 1); y[i] = Q(1.0); }
    CHECK(same(x.dot(y), Q(55.0)));
    CHECK(quadblas_qdot(10, x.data(), 1, y.data(), 1) == 55.0);
    CHECK(quadblas_is_aligned(x.data()));
    // norm(1..5) = sqrt(55)
    DefaultVector<> z(5);
    for (int i = 0; i < 5; ++i) z[i] = Q(i + 1);
    CHECK(same(z.norm(), Sleef_sqrtq1_u05(Q(55.0))));
    // axpy: y <- 2.5 x + y
    y.axpy(Q(2.5), x);
    for (int i = 0; i < 10; ++i) CHECK(same(y[i], Q(2.5 * (i + 1) + 1.0)));
    // strided view
    std::vector<Sleef_quad> raw(30, Q(0.0));
    for (int i = 0; i < 10; ++i) raw[3 * i] = Q(i + 1);
    VectorRowMajor v(raw.data(), 10, 3);
    CHECK(same(v.dot(x), Q(385.0)));
  }
  { // 3x3 gemv, rows sum to 6, 15, 24; then the README call A.gemv(1, x, 0, y)
    MatrixRowMajor A(3, 3);
    VectorRowMajor x(3), y(3);
    for (int i = 0; i < 3; ++i) { x[i] = Q(1.0); for (int j = 0; j < 3; ++j) A(i, j) = Q(3 * i + j + 1); }
    A.gemv(SLEEF_QUAD_C(1.0), x, SLEEF_QUAD_C(0.0), y);
    CHECK(same(y[0], Q(6.0)) && same(y[1], Q(15.0)) && same(y[2], Q(24.0)));
    // C ABI with trans: y <- A^T x
    std::vector<Sleef_quad> yt(3, Q(0.0));
    quadblas_qgemv('R', 'T', 3, 3, 1.0, A.data(), 3, x.data(), 1, 0.0, yt.data(), 1);
    CHECK(same(yt[0], Q(12.0)) && same(yt[1], Q(15.0)) && same(yt[2], Q(18.0)));
  }
  { // 2x2 gemm [[1,2],[3,4]] [[5,6],[7,8]] and col-major containers
    MatrixRowMajor A(2, 2), B(2, 2), C(2, 2);
    A(0, 0) = Q(1); A(0, 1) = Q(2); A(1, 0) = Q(3); A(1, 1) = Q(4);
    B(0, 0) = Q(5); B(0, 1) = Q(6); B(1, 0) = Q(7); B(1, 1) = Q(8);
    A.gemm(Q(1.0), B, Q(0.0), C);
    CHECK(same(C(0, 0), Q(19)) && same(C(0, 1), Q(22)) && same(C(1, 0), Q(43)) && same(C(1, 1), Q(50)));
    MatrixColMajor Ac(2, 2), Bc(2, 2), Cc(2, 2);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) { Ac(i, j) = A(i, j); Bc(i, j) = B(i, j); }
    Ac.gemm(Q(2.0), Bc, Q(0.0), Cc);
    CHECK(same(Cc(0, 0), Q(38)) && same(Cc(1, 1), Q(100)));
  }
  for (int mode = 0; mode < 2; ++mode) { // 150^3 small-integer data vs a naive chain, reference mode then fast (tensor) mode
    qb_set_mode(mode ? QB_MODE_FAST : QB_MODE_REFERENCE);
    qb_set_tensor_path(2);
    const size_t n = 150;
    MatrixRowMajor A(n, n), B(n, n), C(n, n);
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) { A(i, j) = Q((double)((i * 7 + j * 3) % 19) - 9.0); B(i, j) = Q((double)((i * 5 + j * 11) % 23) - 11.0); C(i, j) = Q(1.0); }
    A.gemm(Q(1.5), B, Q(0.5), C);
    int bad = 0;
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) {
      Sleef_quad s = Q(0.0);
      for (size_t l = 0; l < n; ++l) s = Sleef_fmaq1_u05(A(i, l), B(l, j), s);
      bad += !same(C(i, j), Sleef_fmaq1_u05(Q(1.5), s, Sleef_mulq1_u05(Q(0.5), Q(1.0))));
    }
    CHECK(bad == 0);
    // identity * A = A (test_quadblas.cpp:691-712)
    MatrixRowMajor I(n, n), R(n, n);
    for (size_t i = 0; i < n; ++i) I(i, i) = Q(1.0);
    I.gemm(Q(1.0), A, Q(0.0), R);
    bad = 0;
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) bad += !same(R(i, j), A(i, j));
    CHECK(bad == 0);
    // README:141-158 cancellation: (1e20, 1, -1e20) . (1, 1, 1) = 1 exactly
    std::vector<Sleef_quad> x = {Q(1e20), Q(1.0), Q(-1e20)}, y = {Q(1.0), Q(1.0), Q(1.0)};
    CHECK(same(dot(3, x.data(), 1, y.data(), 1), Q(1.0)));
  }
  qb_set_mode(QB_MODE_REFERENCE); qb_set_tensor_path(1);
  { // QuadVector (tests/debug_test.cpp:34-47)
    QuadVector a(Q(2), Q(3)), b(Q(4), Q(5)), c(Q(6), Q(8));
    CHECK(same((a + b).horizontal_sum(), Q(14)));
    QuadVector f = a.fma(b, c);
    CHECK(same(f.get(0), Q(14)) && same(f.get(1), Q(23)));
  }
  { // aligned_alloc / aligned_free and raw free functions
    Sleef_quad *p = aligned_alloc<Sleef_quad>(1000);
    CHECK(p != nullptr && quadblas_is_aligned(p));
    for (int i = 0; i < 1000; ++i) p[i] = Q(i + 1);
    CHECK(same(Sleef_mulq1_u05(dot(1000, p, 1, p, 1), Q(6.0)), Q(1000.0 * 1001.0 * 2001.0))); // sum i^2 = n(n+1)(2n+1)/6
    aligned_free(p);
  }
  if (qb_last_error_code() != 0) { ++fails; std::printf("sticky error: %s\n", qb_last_error()); }
  std::printf(fails ? "dropin_test: %d FAILED\n" : "dropin_test: all passed\n", fails);
  return fails ? 1 : 0;
}

/*
 * oracle/ref_wrap.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the UNMODIFIED reference headers where they lie (-I/root/reference/include, read at
 * build time, never copied) against oracle/shim/sleefquad.h and re-exports the reference's own
 * entry points under ref_* names with full binary128 results, so that
 *   (1) oracle/qoracle.c (the CPU restatement) can be pinned bitwise against the real reference
 *       loops, and
 *   (2) bench.py --impl reference can time the reference's SLEEF-shaped OpenMP path on the GPU
 *       box's host cores (arithmetic = libquadmath, because SLEEF 3.8 is unavailable offline).
 * Output goes to oracle/_ref/libqref.so (git-ignored, travels to the GPU box via gpurun).
 *
 * Entry points wrapped (reference file:line):
 *   QuadBLAS::gemm  include/quadblas/algorithms/level3.hpp:215
 *   QuadBLAS::gemv  include/quadblas/algorithms/level2.hpp:85
 *   QuadBLAS::dot   include/quadblas/algorithms/level1.hpp:80
 *   QuadBLAS::axpy  include/quadblas/algorithms/level1.hpp:190
 *   quadblas_q*     include/quadblas/interface/c_interface.hpp:21-146
 */
#include <quadblas/quadblas.hpp>
#include <cstring>

typedef Sleef_quad Q;

extern "C" {

/* ---- scalar ops exactly as the reference calls them ---- */
void ref_fma(const void *a, const void *b, const void *c, void *out)
{ *(Q *)out = Sleef_fmaq1_u05(*(const Q *)a, *(const Q *)b, *(const Q *)c); }
void ref_mul(const void *a, const void *b, void *out)
{ *(Q *)out = Sleef_mulq1_u05(*(const Q *)a, *(const Q *)b); }
void ref_add(const void *a, const void *b, void *out)
{ *(Q *)out = Sleef_addq1_u05(*(const Q *)a, *(const Q *)b); }
void ref_sqrt(const void *a, void *out) { *(Q *)out = Sleef_sqrtq1_u05(*(const Q *)a); }
void ref_from_double(double d, void *out) { *(Q *)out = Sleef_cast_from_doubleq1(d); }
double ref_to_double(const void *a) { return Sleef_cast_to_doubleq1(*(const Q *)a); }

/* vectorised scalar ops for bulk known-answer generation */
void ref_fma_n(long n, const void *a, const void *b, const void *c, void *out)
{
  const Q *qa = (const Q *)a, *qb = (const Q *)b, *qc = (const Q *)c; Q *qo = (Q *)out;
#pragma omp parallel for
  for (long i = 0; i < n; ++i) qo[i] = Sleef_fmaq1_u05(qa[i], qb[i], qc[i]);
}

/* ---- quad-typed routine wrappers (C++ surface, cpp_classes.hpp:66-81,143-154) ---- */
void ref_gemm(char layout, long m, long n, long k, const void *alpha, const void *A, long lda,
              const void *B, long ldb, const void *beta, void *C, long ldc)
{
  QuadBLAS::Layout L = (layout == 'C' || layout == 'c') ? QuadBLAS::Layout::ColMajor : QuadBLAS::Layout::RowMajor;
  QuadBLAS::gemm(L, (size_t)m, (size_t)n, (size_t)k, *(const Q *)alpha, (const Q *)A, (size_t)lda,
                 (const Q *)B, (size_t)ldb, *(const Q *)beta, (Q *)C, (size_t)ldc);
}
void ref_gemv(char layout, long m, long n, const void *alpha, const void *A, long lda,
              const void *x, long incx, const void *beta, void *y, long incy)
{
  QuadBLAS::Layout L = (layout == 'C' || layout == 'c') ? QuadBLAS::Layout::ColMajor : QuadBLAS::Layout::RowMajor;
  QuadBLAS::gemv(L, (size_t)m, (size_t)n, *(const Q *)alpha, (const Q *)A, (size_t)lda,
                 (const Q *)x, (size_t)incx, *(const Q *)beta, (Q *)y, (size_t)incy);
}
void ref_dot(long n, const void *x, long incx, const void *y, long incy, void *out)
{ *(Q *)out = QuadBLAS::dot((size_t)n, (const Q *)x, (size_t)incx, (const Q *)y, (size_t)incy); }
void ref_nrm2(long n, const void *x, long incx, void *out)
{ *(Q *)out = Sleef_sqrtq1_u05(QuadBLAS::dot((size_t)n, (const Q *)x, (size_t)incx, (const Q *)x, (size_t)incx)); }
void ref_axpy(long n, const void *alpha, const void *x, long incx, void *y, long incy)
{ QuadBLAS::axpy((size_t)n, *(const Q *)alpha, (const Q *)x, (size_t)incx, (Q *)y, (size_t)incy); }

/* ---- the reference C ABI itself, re-exported under ref_c_* (c_interface.hpp) ---- */
double ref_c_qdot(int n, void *x, int incx, void *y, int incy) { return quadblas_qdot(n, x, incx, y, incy); }
double ref_c_qnrm2(int n, void *x, int incx) { return quadblas_qnrm2(n, x, incx); }
void ref_c_qaxpy(int n, double alpha, void *x, int incx, void *y, int incy) { quadblas_qaxpy(n, alpha, x, incx, y, incy); }
void ref_c_qgemv(char layout, char trans, int m, int n, double alpha, void *A, int lda, void *x, int incx,
                 double beta, void *y, int incy)
{ quadblas_qgemv(layout, trans, m, n, alpha, A, lda, x, incx, beta, y, incy); }
void ref_c_qgemm(char layout, char ta, char tb, int m, int n, int k, double alpha, void *A, int lda, void *B,
                 int ldb, double beta, void *C, int ldc)
{ quadblas_qgemm(layout, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc); }
void ref_set_num_threads(int t) { quadblas_set_num_threads(t); }
int ref_get_num_threads(void) { return quadblas_get_num_threads(); }
const char *ref_get_version(void) { return quadblas_get_version(); }
int ref_is_aligned(const void *p) { return quadblas_is_aligned(p); }
const char *ref_arith(void) { return "libquadmath (SLEEF 3.8 unavailable offline)"; }

} /* extern "C" */

/*
 * tests/host/qwide_host.cpp — CPU harness for qblas_b200/csrc/qwide.cuh (host/device dual source).
 *
 * TEST INFRASTRUCTURE: builds the fast-mode window accumulator with g++ as a tiny shared library so
 * that tests/test_host_qwide.py can check it against exact rational arithmetic without a GPU.  The
 * GPU tests re-check the nvcc build of the same source through the C ABI.
 */
#include <cstdint>
#include <cstring>
#include "../../qblas_b200/csrc/q128.cuh"
#include "../../qblas_b200/csrc/q128_chain.cuh"
#include "../../qblas_b200/csrc/qwide.cuh"

using namespace qb;

extern "C" {

/* sum_i x[i*incx] * y[i*incy] with `lanes` window accumulators (element i goes to lane i % lanes,
 * as thread t of the kernels takes elements t, t + T, ...), merged in lane order, rounded once.
 * variant 0: qw_fma on every element; 1: qw_fma_rare on every element (the generic path);
 * 2: the kernels' hot-loop form qwa_fma (+ qwa_fma_rare when it declines) on a scratch column. */
void qwide_dot(int64_t n, const q128 *x, int64_t incx, const q128 *y, int64_t incy, int lanes, int variant, q128 *out,
               uint32_t *bad_out)
{
  qwide *acc = new qwide[lanes];
  uint32_t bad = 0;
  for (int l = 0; l < lanes; ++l) acc[l] = qw_zero();
  if (variant == 2) {
    qwacc *ha = new qwacc[lanes];
    uint32_t col[QWA_COL_WORDS * 3];            /* stride 3: exercises the strided addressing */
    qwa_col_init(col, 3);
    for (int l = 0; l < lanes; ++l) ha[l] = qwa_zero();
    for (int64_t i = 0; i < n; ++i) {
      qwacc &S = ha[i % lanes];
      if (qwa_fma(S, qop_load_n(x[i * incx]), qop_load_n(y[i * incy]), col, 3)) qwa_fma_rare(S, x[i * incx], y[i * incy], bad);
    }
    for (int l = 0; l < lanes; ++l) acc[l] = qwa_fold(ha[l]);
    delete[] ha;
    n = 0;
  }
  for (int64_t i = 0; i < n; ++i) {
    qwide &S = acc[i % lanes];
    if (variant == 0) qw_fma(S, qop_load(x[i * incx]), qop_load(y[i * incy]), bad);
    else S = qw_fma_rare(S, x[i * incx], y[i * incy], &bad);
  }
  qwide v = acc[0];
  for (int l = 1; l < lanes; ++l) qw_merge(v, acc[l]);
  *out = qw_finish(v, bad);
  *bad_out = bad;
  delete[] acc;
}

/* binary merge tree over per-element windows (stress for qw_merge / qw_shr with every anchor gap) */
void qwide_tree(int64_t n, const q128 *x, const q128 *y, q128 *out)
{
  qwide *acc = new qwide[n > 0 ? n : 1];
  uint32_t bad = 0;
  for (int64_t i = 0; i < n; ++i) { acc[i] = qw_zero(); qw_fma(acc[i], qop_load(x[i]), qop_load(y[i]), bad); }
  for (int64_t s = 1; s < n; s *= 2)
    for (int64_t i = 0; i + s < n; i += 2 * s) qw_merge(acc[i], acc[i + s]);
  *out = n > 0 ? qw_finish(acc[0], bad) : q_zero(0);
  delete[] acc;
}
}

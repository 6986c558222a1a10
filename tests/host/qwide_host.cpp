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
#include "../../qblas_b200/csrc/qslice.cuh"

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

/* n single-product windows (after `lanes`-way pre-accumulation: window l takes the products l, l + lanes, ...) summed the way the
 * block reductions of the level-1 kernels do it: one shift to the common anchor, 224-bit integer sum (qw_sum_aligned) */
void qwide_blocksum(int64_t n, const q128 *x, const q128 *y, int lanes, q128 *out)
{
  qwide *acc = new qwide[lanes > 0 ? lanes : 1];
  uint32_t bad = 0;
  for (int l = 0; l < lanes; ++l) acc[l] = qw_zero();
  for (int64_t i = 0; i < n; ++i) qw_fma(acc[i % lanes], qop_load(x[i]), qop_load(y[i]), bad);
  *out = qw_finish(qw_sum_aligned(acc, lanes), bad);
  delete[] acc;
}

/* the sliced FP64 accumulate (qslice.cuh) as k_gemv_row_f64 runs it on one row: thread t of `lanes` takes the elements t, t + lanes, ...
 * of the row a against x, flushes its columns every QS_TILE steps, the windows are merged in lane order and rounded once.
 * Returns 1 when the row passes the acceptance test (the kernel then stores this result), 0 when it would be recomputed by the
 * window kernel, -1 when x itself sends the whole call there (Inf / NaN / subnormal in x).  info = {anchor, EX, dmax, flags}. */
int qslice_dot(int64_t n, const q128 *a, const q128 *x, int lanes, q128 *out, int32_t *info)
{
  int32_t EX = 0;
  for (int64_t j = 0; j < n; ++j) {
    const qop o = qop_load(x[j]);
    if (o.e == 0x7fff || (o.e == 0 && (o.m0 | o.m1 | o.m2 | o.m3))) return -1;
    if (o.e > EX) EX = o.e;
  }
  qs_xrec *rec = new qs_xrec[n > 0 ? n : 1];
  for (int64_t j = 0; j < n; ++j) qs_xrec_make(x[j], EX, rec[j]);
  qwide v = qw_zero();
  int32_t anc = QS_ANCMIN, dmax = QS_EXNONE;
  uint32_t flags = 0;
  for (int t = 0; t < lanes; ++t) {
    qs_cols C = qs_cols_zero();
    qs_row S; S.anc = QS_ANCMIN; S.dmax = QS_EXNONE;
    uint64_t w[4] = {0, 0, 0, 0};
    double col[QS_XCOL];
    qs_col_init(col, 1);
    int steps = 0;
    for (int64_t j = t; j < n; j += lanes) {
      qs_col_set(col, 1, rec[j].X);
      const uint32_t w0 = (uint32_t)a[j].lo, w1 = (uint32_t)(a[j].lo >> 32), w2 = (uint32_t)a[j].hi, w3 = (uint32_t)(a[j].hi >> 32);
      const uint32_t e = (w3 >> 16) & 0x7fffu;
      /* as the kernel does it: the hot step for every element (an element above the anchor or a zero meets only zeros there),
       * then the rare ones again after qs_rare */
      uint32_t sh = (uint32_t)S.anc - e;
      sh = sh > QS_SHMAX ? QS_SHMAX : sh;
      qs_step(C, w0, w1, w2, w3, sh, col, 1);
      if ((uint32_t)(e - 1u) >= (uint32_t)S.anc) {
        sh = qs_rare(C, S, flags, e, w0, w1, w2, w3, w, 1);
        if (sh < QS_SHMAX) qs_step(C, w0, w1, w2, w3, sh, col, 1);
      }
      const int32_t d = (int32_t)e + rec[j].ex;
      S.dmax = d > S.dmax ? d : S.dmax;
      if (++steps == QS_TILE) { qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1); C = qs_cols_zero(); steps = 0; }
    }
    qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1);
    qw_merge(v, qs_to_qwide(w, 1, S.anc, EX));
    anc = S.anc > anc ? S.anc : anc;
    dmax = S.dmax > dmax ? S.dmax : dmax;
  }
  delete[] rec;
  *out = qw_finish(v, 0);
  info[0] = anc; info[1] = EX; info[2] = dmax; info[3] = (int32_t)flags;
  return qs_accept(anc, EX, dmax, flags) ? 1 : 0;
}

/* the sliced sum of squares (qs_square_step) as k_nrm2_f64 runs it: thread t of `lanes` takes the elements t, t + lanes, ...;
 * returns 0 when an Inf / NaN / nonzero subnormal sends the call to the window kernel, else 1 with *out = the rounded sum of squares */
int qslice_sumsq(int64_t n, const q128 *x, int lanes, q128 *out)
{
  qwide v = qw_zero();
  uint32_t flags = 0;
  for (int t = 0; t < lanes; ++t) {
    qs_sq_cols Q = qs_sq_zero();
    qs_row S; S.anc = QS_ANCMIN; S.dmax = 0;
    uint64_t w[4] = {0, 0, 0, 0};
    int steps = 0;
    for (int64_t j = t; j < n; j += lanes) {
      const uint32_t w0 = (uint32_t)x[j].lo, w1 = (uint32_t)(x[j].lo >> 32), w2 = (uint32_t)x[j].hi, w3 = (uint32_t)(x[j].hi >> 32);
      const uint32_t e = (w3 >> 16) & 0x7fffu;
      uint32_t sh = (uint32_t)S.anc - e;
      sh = sh > QS_SHMAX ? QS_SHMAX : sh;
      qs_square_step(Q, w0, w1, w2, w3, sh);
      if ((uint32_t)(e - 1u) >= (uint32_t)S.anc) {
        qs_cols C = qs_sq_columns(Q);
        Q = qs_sq_zero();
        sh = qs_rare(C, S, flags, e, w0, w1, w2, w3, w, 1, 2u);      /* flushes C when it moves the anchor */
        qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1);
        if (sh < QS_SHMAX) qs_square_step(Q, w0, w1, w2, w3, sh);
      }
      if (++steps == QS_TILE) { const qs_cols C = qs_sq_columns(Q); qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1); Q = qs_sq_zero(); steps = 0; }
    }
    const qs_cols C = qs_sq_columns(Q);
    qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1);
    qw_merge(v, qs_to_qwide(w, 1, S.anc, S.anc));
  }
  *out = qw_finish(v, 0);
  return (flags & QS_FALLBACK) ? 0 : 1;
}

/* the sliced dot product of two vectors (qs_dot_step) as k_dot_f64 runs it: thread t of `lanes` takes the pairs t, t + lanes, ... with
 * its own anchors for x and y; returns 1 when the acceptance test passes, 0 when the call goes to the window kernel */
int qslice_dot2(int64_t n, const q128 *x, const q128 *y, int lanes, q128 *out)
{
  qwide v = qw_zero();
  uint32_t flags = 0;
  int32_t ancsum = 0, dmax = QS_EXNONE;
  for (int t = 0; t < lanes; ++t) {
    qs_cols C = qs_cols_zero();
    qs_row SX, SY; SX.anc = SY.anc = QS_ANCMIN; SX.dmax = SY.dmax = 0;
    int32_t dm = QS_EXNONE;
    uint64_t w[4] = {0, 0, 0, 0};
    int steps = 0;
    for (int64_t j = t; j < n; j += lanes) {
      const uint32_t x0 = (uint32_t)x[j].lo, x1 = (uint32_t)(x[j].lo >> 32), x2 = (uint32_t)x[j].hi, x3 = (uint32_t)(x[j].hi >> 32);
      const uint32_t y0 = (uint32_t)y[j].lo, y1 = (uint32_t)(y[j].lo >> 32), y2 = (uint32_t)y[j].hi, y3 = (uint32_t)(y[j].hi >> 32);
      const uint32_t ex = (x3 >> 16) & 0x7fffu, ey = (y3 >> 16) & 0x7fffu;
      const bool rx = (uint32_t)(ex - 1u) >= (uint32_t)SX.anc, ry = (uint32_t)(ey - 1u) >= (uint32_t)SY.anc;
      uint32_t shx = (uint32_t)SX.anc - ex, shy = (uint32_t)SY.anc - ey;
      shx = shx > QS_SHMAX ? QS_SHMAX : shx; shy = shy > QS_SHMAX ? QS_SHMAX : shy;
      if (!rx && !ry) qs_dot_step(C, x0, x1, x2, x3, y0, y1, y2, y3, shx, shy);
      else {   /* a zero / subnormal / Inf / NaN factor or a new largest one: anchors first (each moves the window), then the product */
        if (rx) shx = qs_rare(C, SX, flags, ex, x0, x1, x2, x3, w, 1);
        if (ry) shy = qs_rare(C, SY, flags, ey, y0, y1, y2, y3, w, 1);
        if (shx < QS_SHMAX && shy < QS_SHMAX) qs_dot_step(C, x0, x1, x2, x3, y0, y1, y2, y3, shx, shy);
      }
      if (ex != 0 && ey != 0) { const int32_t d = (int32_t)ex + (int32_t)ey; dm = d > dm ? d : dm; }
      if (++steps == QS_TILE) { qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1); C = qs_cols_zero(); steps = 0; }
    }
    qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, 1);
    qw_merge(v, qs_to_qwide(w, 1, SX.anc, SY.anc));
    ancsum = SX.anc + SY.anc > ancsum ? SX.anc + SY.anc : ancsum;
    dmax = dm > dmax ? dm : dmax;
  }
  *out = qw_finish(v, 0);
  if (flags & QS_FALLBACK) return 0;
  return dmax >= ancsum - QS_ACCEPT ? 1 : 0;
}
}

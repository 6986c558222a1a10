/*
 * qslice.cuh — the FAST-mode accumulate of qgemv / qnrm2 on the FP64 pipe of sm_100a (B200: 64 DFMA per clock and SM).
 *
 * The window accumulate of qwide.cuh costs ~90 integer instructions per element, 28 of them on the 32-bit multiplier
 * (IMAD.WIDE / IMAD.HI: 4-5 clocks per warp instruction and sub-partition, nothing else issues meanwhile —
 * tools/exp/mb_pipes.cu), which caps a streaming kernel at 0.56 of the HBM rate.  Here the product is formed by DFMAs on
 * exact integer-valued doubles and — the larger saving — NOTHING is aligned, carried or normalised per element:
 *
 *   x_j   = sign * sum_l X_l 2^(-22 l) * 2^(EX - QBIAS - 21),  6 slices of 22 bits of the significand aligned to the largest
 *           exponent EX of x (made once per call by a small kernel, 64 bytes per element: qs_xrec);
 *   a     = sign * sum_i A_i 2^(-22 (i + q)) * 2^(anc - QBIAS - 21),  anc - e(a) = 22 q + r: the significand is shifted by
 *           r < 22 bits (5 funnel shifts) and cut at FIXED bit positions into 6 slices (11 integer instructions, 6 I2F);
 *           the slice offset q is not applied to a but to x: the x slices sit in a per-thread shared-memory column behind
 *           six zeros and are read at the dynamic index 6 - q (+ a second copy of -X that the sign of a selects);
 *   C_c  += sum_{i + l' = c} A_i X'_l' for the columns c = 0..5 only (21 DFMAs): every term is an integer < 2^44, a column
 *           sums at most 6 * QS_TILE = 384 of them, so the doubles stay exact (< 2^53);
 *   every QS_TILE elements (and when a larger element raises the anchor) the six columns are added, as integers, into a
 *           256-bit window W at fixed bit positions: value = W * 2^(anc + EX - 2 QBIAS - 152).
 *
 * What is dropped per element: the columns c >= 6 (< 2^24.6 window units), the two bits of a below its last slice and the
 * floor of x to 132 bits (< 2^22 units each): < 2^25 units = 2^-127 * 2^(anc - QBIAS) * 2^(EX - QBIAS), i.e. relative to
 * (largest |a| seen so far) x (largest |x_j|), NOT to the products actually formed.  The fast-mode contract
 * |s^ - s| <= gamma_n sum |a_j||x_j| therefore needs a check: with Dmax = max_j (e(a_j) + e(x_j)) over the
 * row (one VIADDMNMX per element), sum |a_j||x_j| >= 2^(Dmax - 2 QBIAS), and for n >= 128
 *       (n + 64) 2^(anc + EX - 127)  <=  (n - 1) 2^-113 2^Dmax      <=   Dmax >= anc + EX - QS_ACCEPT,  QS_ACCEPT = 12
 * (the 64 covers the per-thread conversions and merges).  A row that fails the test, or holds an Inf / NaN / nonzero
 * subnormal, is recomputed by the exact window kernel of qwide.cuh (k_gemv_row_wide) — same contract, slower; ordinary data
 * (the largest element of a row meets an x_j within 2^-12 of the largest, or any other product comes that close) never is.
 *
 * Dual host/device source (tests/host/qwide_host.cpp builds it with g++; tests/test_host_qslice.py checks it against exact
 * rational arithmetic).
 */
#pragma once
#include "qwide.cuh"
#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace qb {

constexpr int QS_NS = 6;                                      /* slices of x, slices of a, columns kept */
constexpr int QS_SL = 22;                                     /* bits per slice */
constexpr uint32_t QS_MK = (1u << QS_SL) - 1u;
constexpr uint32_t QS_SHMAX = QS_SL * QS_NS + (QS_SL - 1);   /* 153: q = 6, the element only meets the six zeros */
constexpr int QS_TILE = 64;                                   /* elements between two flushes: 6 * 64 * 2^44 < 2^53 */
constexpr int QS_XCOL = 4 * QS_NS;                            /* doubles per thread column: [0,6) zeros, [6,12) +X, [12,18) zeros, [18,24) -X */
constexpr int QS_ACCEPT = 12;
constexpr int32_t QS_EXNONE = -(1 << 20);                     /* e(x_j) of a zero x_j: never the largest e(a) + e(x) */
constexpr uint32_t QS_FALLBACK = 8u;                          /* next to the QW_* flags */
constexpr int32_t QS_ANCMIN = 160;                            /* the anchor every accumulator starts from (elements below it are simply far below the anchor):
                                                                 with anc > QS_SHMAX the kernels' shift min(anc - e, QS_SHMAX) is QS_SHMAX for a zero
                                                                 (e = 0) and for an element above the anchor (the difference wraps), without a special case */

struct qs_cols { double c0, c1, c2, c3, c4, c5; };

struct qs_xrec {                                              /* 64 bytes per element of x */
  double X[QS_NS];
  int32_t ex, pad0;
  double pad1;
};

QB_HD qs_cols qs_cols_zero() { qs_cols z; z.c0 = z.c1 = z.c2 = z.c3 = z.c4 = z.c5 = 0.0; return z; }

/* bits [pos, pos + len) of the little-endian words w[0..nw), len <= 32, zero beyond the array */
QB_HD uint32_t qs_bits(const uint32_t *w, int nw, int pos, int len)
{
  uint64_t v = 0;
  const int k = pos >> 5, o = pos & 31;
  if (k < nw && k >= 0) v = w[k];
  if (k + 1 < nw && k + 1 >= 0) v |= (uint64_t)w[k + 1] << 32;
  return (uint32_t)((v >> o) & ((len >= 32) ? 0xffffffffull : ((1ull << len) - 1ull)));
}

/* the record of one x_j for the anchor EX (the largest exponent field of x): floor(m * 2^(19 - (EX - e))) in 6 signed slices.
 * A zero — and a subnormal, which the caller must have flagged — gives zero slices. */
QB_HD void qs_xrec_make(q128 x, int32_t EX, qs_xrec &r)
{
  const qop o = qop_load(x);
  for (int l = 0; l < QS_NS; ++l) r.X[l] = 0.0;
  r.ex = QS_EXNONE; r.pad0 = 0; r.pad1 = 0.0;
  if (o.e == 0 || o.e == 0x7fff) return;
  r.ex = o.e;
  const int shx = EX - o.e;                                   /* >= 0 */
  if (shx >= 132) return;
  /* G = m << 19 (132 bits) in 5 words, then >> shx */
  uint32_t g[6] = {o.m0 << 19, fshl(o.m0, o.m1, 19), fshl(o.m1, o.m2, 19), fshl(o.m2, o.m3, 19), o.m3 >> 13, 0u};
  for (int l = 0; l < QS_NS; ++l) {
    const uint32_t s = qs_bits(g, 6, shx + 110 - QS_SL * l, QS_SL);
    r.X[l] = o.s ? -(double)s : (double)s;
  }
}

/* (w3 & 0xffff) | 0x10000: the top word of the significand with the implicit bit (one PRMT on the device) */
QB_HD uint32_t qs_top_word(uint32_t w3)
{
#if defined(__CUDA_ARCH__)
  return __byte_perm(w3, 0x00010000u, 0x7610);
#else
  return (w3 & 0xffffu) | 0x10000u;
#endif
}

/* one element: C_c += sum_i A_i X'_(c-i).  w0..w3 = the packed element, nsh = 21 - sh with sh = anc - e clamped to QS_SHMAX
 * (QS_SHMAX itself for an element that must not count) — the kernels get nsh from one fused add-max —, col = this thread's column
 * of QS_XCOL doubles at `stride`. */
QB_HD void qs_step_n(qs_cols &C, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int32_t nsh, const double *col, int stride)
{
  const uint32_t q = (uint32_t)(nsh * -2979 + 21 * 2979) >> 16;   /* sh / 22 for sh <= 153 */
  const uint32_t s = q * 22u + (uint32_t)nsh;                     /* 21 - sh % 22 */
  const uint32_t m3 = qs_top_word(w3);
  const uint32_t v0 = w0 << s, v1 = fshl(w0, w1, s), v2 = fshl(w1, w2, s), v3 = fshl(w2, m3, s), v4 = fshl(m3, 0u, s);
  const double d0 = (double)fshr(v3, v4, 16), d1 = (double)(fshr(v2, v3, 26) & QS_MK), d2 = (double)((v2 >> 4) & QS_MK),
               d3 = (double)(fshr(v1, v2, 14) & QS_MK), d4 = (double)(fshr(v0, v1, 24) & QS_MK), d5 = (double)((v0 >> 2) & QS_MK);
  const double *xq = col + ((int)(QS_NS - q) + (int)(w3 >> 31) * (2 * QS_NS)) * stride;
  const double x0 = xq[0], x1 = xq[stride], x2 = xq[2 * stride], x3 = xq[3 * stride], x4 = xq[4 * stride], x5 = xq[5 * stride];
  C.c0 = fma(d0, x0, C.c0);
  C.c1 = fma(d0, x1, fma(d1, x0, C.c1));
  C.c2 = fma(d0, x2, fma(d1, x1, fma(d2, x0, C.c2)));
  C.c3 = fma(d0, x3, fma(d1, x2, fma(d2, x1, fma(d3, x0, C.c3))));
  C.c4 = fma(d0, x4, fma(d1, x3, fma(d2, x2, fma(d3, x1, fma(d4, x0, C.c4)))));
  C.c5 = fma(d0, x5, fma(d1, x4, fma(d2, x3, fma(d3, x2, fma(d4, x1, fma(d5, x0, C.c5))))));
}
QB_HD void qs_step(qs_cols &C, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t sh, const double *col, int stride)
{
  qs_step_n(C, w0, w1, w2, w3, 21 - (int32_t)sh, col, stride);
}

/* one element of a sum of squares: C_c += sum_{i + l + 2q = c} A_i A_l with q = sh / 22 and the significand shifted by sh % 22.
 * Both factors are the same slices, so the column offset 2q is static per case and nothing goes through shared memory; the
 * symmetric pairs i < l are formed once and kept in accumulators of their own (o1..o5) that count twice when the columns are read
 * (qs_sq_columns): 12 DFMAs for q = 0, 6 for q = 1, 2 for q = 2, none beyond — such an element is below 2^-66 of the largest
 * one, its square below 2^-130 of the sum.  Per element a column takes at most 3 products < 2^44: QS_TILE elements stay exact. */
struct qs_sq_cols { double d0, o1, o2, d2, o3, o4, d4, o5; };

QB_HD qs_sq_cols qs_sq_zero() { qs_sq_cols z; z.d0 = z.o1 = z.o2 = z.d2 = z.o3 = z.o4 = z.d4 = z.o5 = 0.0; return z; }

QB_HD qs_cols qs_sq_columns(const qs_sq_cols &Q)
{
  qs_cols C;
  C.c0 = Q.d0; C.c1 = Q.o1 + Q.o1; C.c2 = Q.o2 + Q.o2 + Q.d2; C.c3 = Q.o3 + Q.o3; C.c4 = Q.o4 + Q.o4 + Q.d4; C.c5 = Q.o5 + Q.o5;
  return C;
}

QB_HD void qs_square_step(qs_sq_cols &Q, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t sh)
{
  if (sh >= 3u * QS_SL) return;
  const uint32_t q = (sh * 2979u) >> 16;
  const uint32_t s = 21u - (sh - q * 22u);
  const uint32_t m3 = qs_top_word(w3);
  const uint32_t v0 = w0 << s, v1 = fshl(w0, w1, s), v2 = fshl(w1, w2, s), v3 = fshl(w2, m3, s), v4 = fshl(m3, 0u, s);
  const double d0 = (double)fshr(v3, v4, 16), d1 = (double)(fshr(v2, v3, 26) & QS_MK);
  if (q == 0u) {
    const double d2 = (double)((v2 >> 4) & QS_MK), d3 = (double)(fshr(v1, v2, 14) & QS_MK), d4 = (double)(fshr(v0, v1, 24) & QS_MK),
                 d5 = (double)((v0 >> 2) & QS_MK);
    Q.d0 = fma(d0, d0, Q.d0);
    Q.o1 = fma(d0, d1, Q.o1);
    Q.o2 = fma(d0, d2, Q.o2); Q.d2 = fma(d1, d1, Q.d2);
    Q.o3 = fma(d0, d3, fma(d1, d2, Q.o3));
    Q.o4 = fma(d0, d4, fma(d1, d3, Q.o4)); Q.d4 = fma(d2, d2, Q.d4);
    Q.o5 = fma(d0, d5, fma(d1, d4, fma(d2, d3, Q.o5)));
  } else if (q == 1u) {
    const double d2 = (double)((v2 >> 4) & QS_MK), d3 = (double)(fshr(v1, v2, 14) & QS_MK);
    Q.d2 = fma(d0, d0, Q.d2);
    Q.o3 = fma(d0, d1, Q.o3);
    Q.o4 = fma(d0, d2, Q.o4); Q.d4 = fma(d1, d1, Q.d4);
    Q.o5 = fma(d0, d3, fma(d1, d2, Q.o5));
  } else {
    Q.d4 = fma(d0, d0, Q.d4);
    Q.o5 = fma(d0, d1, Q.o5);
  }
}

/* one product x * y of a dot product with BOTH factors sliced on the fly (no table): shx = ancx - e(x), shy = ancy - e(y), each
 * clamped to QS_SHMAX.  With sh = 22 q + r the significand is shifted by r and the slice offset q is static per case:
 * C_c += sum_{i + l + Q = c} X_i Y_l, Q = qx + qy = 0..5 (21, 15, 10, 6, 3, 1 DFMAs; nothing beyond).  The sign of the product is
 * put into the Y slices (one XOR on the high word of each double). */
QB_HD void qs_slices(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t r, double (&d)[QS_NS])
{
  const uint32_t s = 21u - r;
  const uint32_t m3 = qs_top_word(w3);
  const uint32_t v0 = w0 << s, v1 = fshl(w0, w1, s), v2 = fshl(w1, w2, s), v3 = fshl(w2, m3, s), v4 = fshl(m3, 0u, s);
  d[0] = (double)fshr(v3, v4, 16); d[1] = (double)(fshr(v2, v3, 26) & QS_MK); d[2] = (double)((v2 >> 4) & QS_MK);
  d[3] = (double)(fshr(v1, v2, 14) & QS_MK); d[4] = (double)(fshr(v0, v1, 24) & QS_MK); d[5] = (double)((v0 >> 2) & QS_MK);
}
QB_HD double qs_flip(double v, uint32_t signbit31)   /* v with its sign XORed by bit 31 of signbit31 */
{
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) ^ (int)(signbit31 & 0x80000000u), __double2loint(v));
#else
  return (signbit31 & 0x80000000u) ? -v : v;
#endif
}
QB_HD void qs_dot_step(qs_cols &C, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y0, uint32_t y1, uint32_t y2, uint32_t y3,
                       uint32_t shx, uint32_t shy)
{
  const uint32_t qx = (shx * 2979u) >> 16, qy = (shy * 2979u) >> 16, Q = qx + qy;
  if (Q >= (uint32_t)QS_NS) return;
  double a[QS_NS], b[QS_NS];
  qs_slices(x0, x1, x2, x3, shx - qx * 22u, a);
  qs_slices(y0, y1, y2, y3, shy - qy * 22u, b);
  const uint32_t sg = x3 ^ y3;
  for (int l = 0; l < QS_NS; ++l) b[l] = qs_flip(b[l], sg);
  if (Q == 0u) {
    C.c0 = fma(a[0], b[0], C.c0);
    C.c1 = fma(a[0], b[1], fma(a[1], b[0], C.c1));
    C.c2 = fma(a[0], b[2], fma(a[1], b[1], fma(a[2], b[0], C.c2)));
    C.c3 = fma(a[0], b[3], fma(a[1], b[2], fma(a[2], b[1], fma(a[3], b[0], C.c3))));
    C.c4 = fma(a[0], b[4], fma(a[1], b[3], fma(a[2], b[2], fma(a[3], b[1], fma(a[4], b[0], C.c4)))));
    C.c5 = fma(a[0], b[5], fma(a[1], b[4], fma(a[2], b[3], fma(a[3], b[2], fma(a[4], b[1], fma(a[5], b[0], C.c5))))));
  } else if (Q == 1u) {
    C.c1 = fma(a[0], b[0], C.c1);
    C.c2 = fma(a[0], b[1], fma(a[1], b[0], C.c2));
    C.c3 = fma(a[0], b[2], fma(a[1], b[1], fma(a[2], b[0], C.c3)));
    C.c4 = fma(a[0], b[3], fma(a[1], b[2], fma(a[2], b[1], fma(a[3], b[0], C.c4))));
    C.c5 = fma(a[0], b[4], fma(a[1], b[3], fma(a[2], b[2], fma(a[3], b[1], fma(a[4], b[0], C.c5)))));
  } else if (Q == 2u) {
    C.c2 = fma(a[0], b[0], C.c2);
    C.c3 = fma(a[0], b[1], fma(a[1], b[0], C.c3));
    C.c4 = fma(a[0], b[2], fma(a[1], b[1], fma(a[2], b[0], C.c4)));
    C.c5 = fma(a[0], b[3], fma(a[1], b[2], fma(a[2], b[1], fma(a[3], b[0], C.c5))));
  } else if (Q == 3u) {
    C.c3 = fma(a[0], b[0], C.c3);
    C.c4 = fma(a[0], b[1], fma(a[1], b[0], C.c4));
    C.c5 = fma(a[0], b[2], fma(a[1], b[1], fma(a[2], b[0], C.c5)));
  } else if (Q == 4u) {
    C.c4 = fma(a[0], b[0], C.c4);
    C.c5 = fma(a[0], b[1], fma(a[1], b[0], C.c5));
  } else {
    C.c5 = fma(a[0], b[0], C.c5);
  }
}

/* the thread column: zeros in front of both copies (once), then the slices of the current x_j */
QB_HD void qs_col_init(double *col, int stride)
{
  for (int k = 0; k < QS_NS; ++k) col[k * stride] = col[(2 * QS_NS + k) * stride] = 0.0;
}
QB_HD void qs_col_set(double *col, int stride, const double (&X)[QS_NS])
{
  for (int k = 0; k < QS_NS; ++k) { col[(QS_NS + k) * stride] = X[k]; col[(3 * QS_NS + k) * stride] = -X[k]; }
}

/* ---- the 256-bit window: 4 little-endian 64-bit limbs at `stride`, two's complement ---- */
QB_HD void qs_win_add(uint64_t *w, int stride, int limb, uint64_t lo, uint64_t hi, uint64_t ext)
{
  uint64_t add[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) add[k] = k < limb ? 0ull : (k == limb ? lo : (k == limb + 1 ? hi : ext));
  uint64_t c = 0;
  for (int k = 0; k < 4; ++k) {
    const uint64_t a = w[k * stride], s1 = a + add[k], s2 = s1 + c;
    c = (uint64_t)(s1 < a) + (uint64_t)(s2 < s1);
    w[k * stride] = s2;
  }
}

/* W += sum_c I_c 2^(22 (5 - c)), the columns being exact integers of magnitude < 2^53 */
QB_HD void qs_flush(double c0, double c1, double c2, double c3, double c4, double c5, uint64_t *w, int stride)
{
  const double c[QS_NS] = {c0, c1, c2, c3, c4, c5};
  for (int k = 0; k < QS_NS; ++k) {
    const int64_t v = (int64_t)c[k];
    if (v == 0) continue;
    const int sh = QS_SL * (QS_NS - 1 - k), limb = sh >> 6, off = sh & 63;
    const uint64_t ext = v < 0 ? ~0ull : 0ull;
    const uint64_t lo = (uint64_t)v << off;
    const uint64_t hi = off ? (uint64_t)(v >> (64 - off)) : ext;   /* arithmetic shift: the sign fills */
    qs_win_add(w, stride, limb, lo, hi, ext);
  }
}

/* W >>= d (arithmetic, floor) */
QB_HD void qs_win_shr(uint64_t *w, int stride, uint32_t d)
{
  uint64_t v[4] = {w[0], w[stride], w[2 * stride], w[3 * stride]};
  const uint64_t sg = (uint64_t)((int64_t)v[3] >> 63);
  const uint32_t lq = d >> 6, r = d & 63;
  for (int k = 0; k < 4; ++k) {
    const uint64_t lo = (k + lq < 4) ? v[k + lq] : sg;
    const uint64_t hi = (k + lq + 1 < 4) ? v[k + lq + 1] : sg;
    w[k * stride] = (d >= 256) ? sg : (r ? ((lo >> r) | (hi << (64 - r))) : lo);
  }
}

/* the window at anchors (anc, EX) as a qwide: W >> 24 (floor: < 2^24 units, counted in QS_ACCEPT), E = anc + EX + 31 */
QB_HD qwide qs_to_qwide(const uint64_t *w, int stride, int32_t anc, int32_t EX)
{
  const uint64_t a = w[0], b = w[stride], c = w[2 * stride], d = w[3 * stride];
  const uint64_t l0 = (a >> 24) | (b << 40), l1 = (b >> 24) | (c << 40), l2 = (c >> 24) | (d << 40);
  qwide s;
  s.w0 = (uint32_t)l0; s.w1 = (uint32_t)(l0 >> 32); s.w2 = (uint32_t)l1; s.w3 = (uint32_t)(l1 >> 32);
  s.w4 = (uint32_t)l2; s.w5 = (uint32_t)(l2 >> 32);
  s.E = anc + EX + 31;
  return s;
}

/* per (thread, row) state outside the columns */
struct qs_row {
  int32_t anc;                       /* anchor: the largest exponent field seen so far, at least QS_ANCMIN */
  int32_t dmax;                      /* max_j e(a_j) + e(x_j) */
};

/* what the hot loop does when (e - 1) >= anc as unsigned numbers: e = 0 (zero: nothing; subnormal: the row falls back),
 * e = 0x7fff (Inf / NaN: falls back) or a new largest element (the columns go to the window at the old anchor, the window is
 * shifted to the new one).  Returns the shift for qs_step. */
QB_HD uint32_t qs_rare(qs_cols &C, qs_row &S, uint32_t &flags, uint32_t e, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint64_t *w,
                       int stride, uint32_t anchors = 1u /* 2 for a sum of squares: both factors follow the anchor */)
{
  if (e == 0x7fffu) { flags |= QS_FALLBACK; return QS_SHMAX; }
  if (e == 0u) { if ((w0 | w1 | w2 | (w3 & 0xffffu)) != 0u) flags |= QS_FALLBACK; return QS_SHMAX; }
  qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, stride);
  C = qs_cols_zero();
  qs_win_shr(w, stride, anchors * (e - (uint32_t)S.anc));
  S.anc = (int32_t)e;
  return 0u;
}

/* the acceptance test of a whole row (see the header) */
QB_HD bool qs_accept(int32_t anc, int32_t EX, int32_t dmax, uint32_t flags)
{
  if (flags & QS_FALLBACK) return false;
  return dmax >= anc + EX - QS_ACCEPT;                        /* (a row of zeros fails it too: the window kernel returns its +0) */
}

} // namespace qb

/*
 * qwide.cuh — the FAST-mode accumulate for qdot / qnrm2 / qgemv:  S <- S + a*b  with NO rounding,
 * in a 192-bit two's-complement fixed-point window anchored at the largest product seen so far.
 *
 * The reference accumulates with one correctly rounded FMA per element
 * (/root/reference/include/quadblas/algorithms/level1.hpp:24,31; level2.hpp:43,79).  Re-running
 * that chain on the GPU costs ~170 integer instructions per element (normalise + RNE every step,
 * q128_chain.cuh) and caps a streaming kernel at the integer-issue rate, far below HBM.  Fast mode
 * only promises |r^ - r| <= gamma_k * sum|a_i||b_i|, so the per-step normalise and
 * round are dropped:
 *
 *   value(S) = W * 2^(E - 2*QW_EOFF + 64),   W = signed 192-bit integer (w5 = top limb)
 *   step     : P = ma * mb (exact, 226 bits), d = E - (ea + eb) >= 0,
 *              W += sign * floor(P / 2^(64 + d))                    (truncate the magnitude)
 *   re-anchor: a product with ea + eb > E first shifts W right (arithmetic) and raises E.
 *
 * A product at the anchor has its MSB at window bit 160/161, so every term is truncated at least
 * 160 bits below the largest product: error per term <= 2^-160 * max_i|a_i b_i| — forty-seven bits
 * below the u = 2^-113 of a single correctly rounded operation — and the window has 29 bits of
 * carry headroom (2^29 terms per accumulator; the kernels stay far below that).  One RNE rounding
 * happens at the very end (qw_round).  Zeros and subnormal operands take a slower exact path;
 * Inf/NaN operands only record the class of their product in `bad` (+Inf, -Inf, NaN incl. Inf*0),
 * and qw_finish turns the flags into the IEEE result of the chain: NaN if any NaN or Inf - Inf,
 * else the infinity seen.
 *
 * Dual host/device source (tests/host/qwide_host.cpp builds it with g++).
 */
#pragma once
#include "q128_chain.cuh"

namespace qb {

constexpr int32_t QW_EOFF = QBIAS + 112;      /* value(a) = ma * 2^(ea - QW_EOFF) */
constexpr int32_t QW_EMPTY = -(1 << 28);      /* anchor of an accumulator that has seen nothing */

/* `bad` flags collected next to an accumulator: which non-finite products were seen */
constexpr uint32_t QW_PINF = 1u, QW_NINF = 2u, QW_NAN = 4u;

struct qwide {
  uint32_t w0, w1, w2, w3, w4, w5;
  int32_t E;
};

QB_HD qwide qw_zero()
{
  qwide z;
  z.w0 = z.w1 = z.w2 = z.w3 = z.w4 = z.w5 = 0;
  z.E = QW_EMPTY;
  return z;
}

/* W >>= s (arithmetic, floor), s >= 0.  Generic form: only the rare paths and the tree use it. */
QB_HD void qw_shr(qwide &S, uint32_t s)
{
  uint32_t w[8] = {S.w0, S.w1, S.w2, S.w3, S.w4, S.w5, 0, 0};
  const uint32_t sg = (uint32_t)((int32_t)S.w5 >> 31);
  w[6] = w[7] = sg;
  if (s >= 192) {
    S.w0 = S.w1 = S.w2 = S.w3 = S.w4 = S.w5 = sg;
    return;
  }
  const uint32_t wq = s >> 5, r = s & 31;
  uint32_t o[6];
  for (int j = 0; j < 6; ++j) {
    const uint32_t lo = (j + wq < 6) ? w[(j + wq) & 7] : sg;
    const uint32_t hi = (j + wq + 1 < 6) ? w[(j + wq + 1) & 7] : sg;
    o[j] = fshr(lo, hi, r);
  }
  S.w0 = o[0]; S.w1 = o[1]; S.w2 = o[2]; S.w3 = o[3]; S.w4 = o[4]; S.w5 = o[5];
}

/* W += (x ^ mask) + (mask & 1): adds x (mask = 0) or subtracts it (mask = ~0) */
QB_HD void qw_addsub6(qwide &S, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5, uint32_t mask)
{
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %6, %6, %6;\n\t"              /* carry flag <- mask & 1 (mask is 0 or ~0); clobbers the copy */
      "addc.cc.u32 %0, %0, %7;\n\t"
      "addc.cc.u32 %1, %1, %8;\n\t"
      "addc.cc.u32 %2, %2, %9;\n\t"
      "addc.cc.u32 %3, %3, %10;\n\t"
      "addc.cc.u32 %4, %4, %11;\n\t"
      "addc.u32 %5, %5, %12;"
      : "+r"(S.w0), "+r"(S.w1), "+r"(S.w2), "+r"(S.w3), "+r"(S.w4), "+r"(S.w5), "+r"(mask)
      : "r"(x0 ^ mask), "r"(x1 ^ mask), "r"(x2 ^ mask), "r"(x3 ^ mask), "r"(x4 ^ mask), "r"(x5 ^ mask));
#else
  const uint32_t x[6] = {x0 ^ mask, x1 ^ mask, x2 ^ mask, x3 ^ mask, x4 ^ mask, x5 ^ mask};
  uint32_t *w[6] = {&S.w0, &S.w1, &S.w2, &S.w3, &S.w4, &S.w5};
  uint64_t c = mask & 1u;
  for (int j = 0; j < 6; ++j) {
    c += (uint64_t)*w[j] + x[j];
    *w[j] = (uint32_t)c;
    c >>= 32;
  }
#endif
}

/* t[0..5] = limbs 2+wq .. 7+wq of p[0..7] (zero beyond), wq in [0, 6]; then >> r, r in [0, 31] */
QB_HD void qw_align(const uint32_t *p, uint32_t wq, uint32_t r, uint32_t &f0, uint32_t &f1, uint32_t &f2, uint32_t &f3,
                    uint32_t &f4, uint32_t &f5)
{
  uint32_t t0, t1, t2, t3, t4, t5;
  {
    const bool w4 = (wq & 4u) != 0;
    t0 = w4 ? p[6] : p[2]; t1 = w4 ? p[7] : p[3]; t2 = w4 ? 0u : p[4]; t3 = w4 ? 0u : p[5];
    t4 = w4 ? 0u : p[6];   t5 = w4 ? 0u : p[7];
  }
  {
    const bool w2 = (wq & 2u) != 0;
    t0 = w2 ? t2 : t0; t1 = w2 ? t3 : t1; t2 = w2 ? t4 : t2; t3 = w2 ? t5 : t3;
    t4 = w2 ? 0u : t4; t5 = w2 ? 0u : t5;
  }
  {
    const bool w1 = (wq & 1u) != 0;
    t0 = w1 ? t1 : t0; t1 = w1 ? t2 : t1; t2 = w1 ? t3 : t2; t3 = w1 ? t4 : t3;
    t4 = w1 ? t5 : t4; t5 = w1 ? 0u : t5;
  }
  f0 = fshr(t0, t1, r); f1 = fshr(t1, t2, r); f2 = fshr(t2, t3, r); f3 = fshr(t3, t4, r); f4 = fshr(t4, t5, r);
  f5 = t5 >> r;
}

/* The rare path: exponent field 0 (zero / subnormal) or 0x7fff (Inf / NaN) in an operand, or a
 * product above the anchor.  Out of line, operands and accumulator by value. */
QB_HD_NOINLINE qwide qw_fma_rare(qwide S, q128 a, q128 b, uint32_t *bad)
{
  const uint32_t ea0 = (uint32_t)(a.hi >> 48) & 0x7fff, eb0 = (uint32_t)(b.hi >> 48) & 0x7fff;
  if (ea0 == 0x7fff || eb0 == 0x7fff) { /* class of the non-finite product: +Inf, -Inf or NaN */
    if (q_is_nan(a) || q_is_nan(b) || q_is_zero(a) || q_is_zero(b)) *bad |= QW_NAN;
    else *bad |= ((a.hi ^ b.hi) >> 63) ? QW_NINF : QW_PINF;
    return S;
  }
  if (q_is_zero(a) || q_is_zero(b)) return S;
  const qunp ua = q_unpack_finite(a), ub = q_unpack_finite(b);  /* subnormals normalised, e <= 0 */
  const int32_t ep = ua.e + ub.e;
  if (ep > S.E) {
    if (S.E != QW_EMPTY) qw_shr(S, (uint32_t)(ep - S.E));
    S.E = ep;
  }
  int32_t d = S.E - ep;
  d = d > 192 ? 192 : d;
  uint32_t p[8];
  mul4x4((uint32_t)ua.ml, (uint32_t)(ua.ml >> 32), (uint32_t)ua.mh, (uint32_t)(ua.mh >> 32),
         (uint32_t)ub.ml, (uint32_t)(ub.ml >> 32), (uint32_t)ub.mh, (uint32_t)(ub.mh >> 32), p);
  uint32_t f0, f1, f2, f3, f4, f5;
  qw_align(p, (uint32_t)d >> 5, (uint32_t)d & 31u, f0, f1, f2, f3, f4, f5);
  qw_addsub6(S, f0, f1, f2, f3, f4, f5, 0u - (ua.s ^ ub.s));
  return S;
}

/* S <- S + A*B (see the file header).  `bad` collects the Inf/NaN flag. */
QB_HD void qw_fma(qwide &S, const qop &A, const qop &B, uint32_t &bad)
{
  const bool normal = ((uint32_t)A.e - 1u < 0x7ffeu) && ((uint32_t)B.e - 1u < 0x7ffeu);
  const int32_t ep = A.e + B.e;
  int32_t d = S.E - ep;
  if (!normal || d < 0) {
    uint32_t bd = 0;
    S = qw_fma_rare(S, qop_pack(A), qop_pack(B), &bd);
    bad |= bd;
    return;
  }
  uint32_t p[8];
  mul4x4(A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, p);
  d = d > 192 ? 192 : d;
  uint32_t f0, f1, f2, f3, f4, f5;
  qw_align(p, (uint32_t)d >> 5, (uint32_t)d & 31u, f0, f1, f2, f3, f4, f5);
  qw_addsub6(S, f0, f1, f2, f3, f4, f5, 0u - (A.s ^ B.s));
}

/* S <- S + T (exact up to the window truncation of the lower-anchored one) */
QB_HD void qw_merge(qwide &S, qwide T)
{
  if (T.E == QW_EMPTY) return;
  if (S.E == QW_EMPTY) { S = T; return; }
  if (T.E > S.E) { qw_shr(S, (uint32_t)(T.E - S.E)); S.E = T.E; }
  else if (S.E > T.E) qw_shr(T, (uint32_t)(S.E - T.E));
  qw_addsub6(S, T.w0, T.w1, T.w2, T.w3, T.w4, T.w5, 0u);
  /* keep the carry headroom whatever the number of merged partials: once |W| >= 2^188 give back
   * 16 bits at the bottom (still >= 144 bits below the largest product) */
  const int32_t top = (int32_t)S.w5 >> 28;
  if (top != 0 && top != -1) { qw_shr(S, 16); S.E += 16; }
}

/* one RNE rounding of the window to binary128 (overflow -> Inf, gradual underflow handled by
 * q_round_pack).  An empty or exactly cancelled window is +0, as the reference's chain from +0
 * gives for an empty sum (level1.hpp:83-84) and for x + (-x) under RNE. */
QB_HD q128 qw_round(const qwide &S)
{
  if (S.E == QW_EMPTY) return q_zero(0);
  uint32_t w[6] = {S.w0, S.w1, S.w2, S.w3, S.w4, S.w5};
  const uint32_t sign = w[5] >> 31;
  if (sign) { /* magnitude = -W */
    uint64_t c = 1;
    for (int j = 0; j < 6; ++j) { c += (uint64_t)(~w[j]); w[j] = (uint32_t)c; c >>= 32; }
  }
  u256 R;
  R.w0 = 0;
  R.w1 = ((uint64_t)w[1] << 32) | w[0];
  R.w2 = ((uint64_t)w[3] << 32) | w[2];
  R.w3 = ((uint64_t)w[5] << 32) | w[4];
  if (u256_is_zero(R)) return q_zero(0);
  const int lz = u256_clz(R);
  R = u256_shl(R, (uint32_t)lz);
  /* value = Wmag * 2^(E - 2*EOFF + 64); R = Wmag << (64 + lz), MSB at bit 255:
   * value = R * 2^(er - QBIAS - 255)  =>  er = E - 2*EOFF - lz + QBIAS + 255 */
  const int32_t er = S.E - 2 * QW_EOFF - lz + QBIAS + 255;
  return q_round_pack(sign, er, R);
}

/* final result of a reduction: non-finite class from the flags, else the rounded window */
QB_HD q128 qw_finish(const qwide &S, uint32_t bad)
{
  if ((bad & QW_NAN) || (bad & (QW_PINF | QW_NINF)) == (QW_PINF | QW_NINF)) return q_nan();
  if (bad & QW_PINF) return q_inf(0);
  if (bad & QW_NINF) return q_inf(1);
  return qw_round(S);
}

} // namespace qb

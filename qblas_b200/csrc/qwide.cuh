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
 *   value(S) = W * 2^(E - 2*QW_EOFF + QW_SH0),   W = signed 192-bit integer (w5 = top limb), QW_SH0 = 65
 *   step     : P = ma * mb (exact, 226 bits), d = E - (ea + eb) >= 0,
 *              W += sign * floor(P / 2^(QW_SH0 + d))                (truncate the magnitude)
 *   re-anchor: a product with ea + eb > E first shifts W right (arithmetic) and raises E.
 *
 * A product at the anchor has its MSB at window bit 159/160 and the anchor is never more than
 * QW_SLACK = 24 bits above the largest product seen so far, so every term is truncated at least
 * 135 bits below the largest product: error per term < 2^-133 * max_i|a_i b_i| for every variant
 * below (the hot-loop form qwa_fma may be off by two window units, see there) — twenty bits below
 * the u = 2^-113 of a single correctly rounded operation — and the window has 30 bits of carry
 * headroom (2^30 terms per accumulator; the kernels stay far below that).  One RNE rounding
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
constexpr int32_t QW_SLACK = 24;              /* a re-anchor leaves this many bits of room: 32 lanes x several accumulators
                                                 per warp each meet a new largest product O(log n) times, and every one of
                                                 them sends the whole warp through the out-of-line path */
constexpr int32_t QW_SH0 = 65;                /* P >> QW_SH0 is what a product AT the anchor adds (= 32*2 + 1: see qwa_fma) */

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
  if (ep > S.E) { /* re-anchor QW_SLACK bits above the new largest product */
    const int32_t ne = ep + QW_SLACK;
    if (S.E != QW_EMPTY) qw_shr(S, (uint32_t)(ne - S.E));
    S.E = ne;
  }
  int32_t d = S.E - ep + (QW_SH0 - 64);
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
  d += QW_SH0 - 64;
  d = d > 192 ? 192 : d;
  uint32_t f0, f1, f2, f3, f4, f5;
  qw_align(p, (uint32_t)d >> 5, (uint32_t)d & 31u, f0, f1, f2, f3, f4, f5);
  qw_addsub6(S, f0, f1, f2, f3, f4, f5, 0u - (A.s ^ B.s));
}

/* p2..p7 = words 2..7 of a*b for 113-bit mantissas (a3, b3 < 2^17), without the column-0 product a0*b0
 * and without the low halves of the two column-1 products: the value is short of the exact
 * (a*b) >> 64 by less than 3 units of p2, i.e. 3 * 2^-160 of the product.  13 wide multiplies and 2
 * high multiplies instead of 16 wide ones; the two carry-save chains are ordered so that every carry
 * is consumed by the next multiply-add (no carry is ever parked in a register):
 *   even words:  C = (3:2) <- a0b2 + a1b1 + a2b0,  D = (5:4) <- a1b3 + a3b1 + a2b2 + carries(C),  F = (7:6) <- a3b3 + carry(D)
 *   odd words :  A = (4:3) <- a0b3 + a3b0 + a1b2 + a2b1,  B = (6:5) <- a2b3 + a3b2 + carries(A)
 * (a_i b_3 and a_3 b_j are below 2^49, which is why D's and B's first sums cannot carry). */
QB_HD void mul4x4_top6(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3, const uint32_t b0,
                       const uint32_t b1, const uint32_t b2, const uint32_t b3, uint32_t &p2, uint32_t &p3, uint32_t &p4,
                       uint32_t &p5, uint32_t &p6, uint32_t &p7)
{
#if defined(__CUDA_ARCH__)
  uint32_t e2, e3, e4, e5, e6, e7, o3, o4, o5, o6;
  asm("{\n\t"
      "mul.lo.u32      %0, %10, %16;\n\t"      "mul.hi.u32      %1, %10, %16;\n\t"       /* C  = a0*b2 */
      "mad.lo.cc.u32   %0, %11, %15, %0;\n\t"  "madc.hi.cc.u32  %1, %11, %15, %1;\n\t"   /* C += a1*b1 */
      "madc.lo.cc.u32  %2, %11, %17, 0;\n\t"   "madc.hi.u32     %3, %11, %17, 0;\n\t"    /* D  = a1*b3 + cy */
      "mad.lo.cc.u32   %0, %12, %14, %0;\n\t"  "madc.hi.cc.u32  %1, %12, %14, %1;\n\t"   /* C += a2*b0 */
      "madc.lo.cc.u32  %2, %13, %15, %2;\n\t"  "madc.hi.u32     %3, %13, %15, %3;\n\t"   /* D += a3*b1 + cy */
      "mad.lo.cc.u32   %2, %12, %16, %2;\n\t"  "madc.hi.cc.u32  %3, %12, %16, %3;\n\t"   /* D += a2*b2 */
      "madc.lo.cc.u32  %4, %13, %17, 0;\n\t"   "madc.hi.u32     %5, %13, %17, 0;\n\t"    /* F  = a3*b3 + cy */
      "mul.lo.u32      %6, %10, %17;\n\t"      "mul.hi.u32      %7, %10, %17;\n\t"       /* A  = a0*b3 */
      "mad.lo.cc.u32   %6, %13, %14, %6;\n\t"  "madc.hi.u32     %7, %13, %14, %7;\n\t"   /* A += a3*b0 */
      "mad.lo.cc.u32   %6, %11, %16, %6;\n\t"  "madc.hi.cc.u32  %7, %11, %16, %7;\n\t"   /* A += a1*b2 */
      "madc.lo.cc.u32  %8, %12, %17, 0;\n\t"   "madc.hi.u32     %9, %12, %17, 0;\n\t"    /* B  = a2*b3 + cy */
      "mad.lo.cc.u32   %6, %12, %15, %6;\n\t"  "madc.hi.cc.u32  %7, %12, %15, %7;\n\t"   /* A += a2*b1 */
      "madc.lo.cc.u32  %8, %13, %16, %8;\n\t"  "madc.hi.u32     %9, %13, %16, %9;\n\t"   /* B += a3*b2 + cy */
      "}"
      : "=&r"(e2), "=&r"(e3), "=&r"(e4), "=&r"(e5), "=&r"(e6), "=&r"(e7), "=&r"(o3), "=&r"(o4), "=&r"(o5), "=&r"(o6)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3));
  /* word 2 also takes the high halves of the column-1 products; then even + (odd << 32).  opaque(): keep
   * the two high multiplies as plain IMAD.HI (fused into a multiply-add they need a zeroed register pair) */
  const uint32_t h1 = opaque(__umulhi(a0, b1)), h2 = opaque(__umulhi(a1, b0));
  uint64_t t = (uint64_t)e2 + h1 + h2;
  p2 = (uint32_t)t; t = (t >> 32) + e3 + o3;
  p3 = (uint32_t)t; t = (t >> 32) + e4 + o4;
  p4 = (uint32_t)t; t = (t >> 32) + e5 + o5;
  p5 = (uint32_t)t; t = (t >> 32) + e6 + o6;
  p6 = (uint32_t)t; t = (t >> 32) + e7;
  p7 = (uint32_t)t;
#else
  typedef unsigned __int128 u128;
  const u128 c2 = (u128)((uint64_t)a0 * b2) + (uint64_t)a1 * b1 + (uint64_t)a2 * b0 +
                  (uint32_t)(((uint64_t)a0 * b1) >> 32) + (uint32_t)(((uint64_t)a1 * b0) >> 32);
  const u128 c3 = (u128)((uint64_t)a0 * b3) + (uint64_t)a1 * b2 + (uint64_t)a2 * b1 + (uint64_t)a3 * b0;
  const u128 c4 = (u128)((uint64_t)a1 * b3) + (uint64_t)a2 * b2 + (uint64_t)a3 * b1;
  const u128 c5 = (u128)((uint64_t)a2 * b3) + (uint64_t)a3 * b2;
  const u128 c6 = (u128)((uint64_t)a3 * b3);
  u128 t = c2;
  p2 = (uint32_t)t; t = (t >> 32) + c3;
  p3 = (uint32_t)t; t = (t >> 32) + c4;
  p4 = (uint32_t)t; t = (t >> 32) + c5;
  p5 = (uint32_t)t; t = (t >> 32) + c6;
  p6 = (uint32_t)t; t >>= 32;
  p7 = (uint32_t)t;
#endif
}

/* ------------------------------------------------------------------ hot-loop form (qdot / qnrm2 / qgemv kernels)
 * The generic step above spends most of its instructions on the ALU pipe (a 3-stage word multiplexer,
 * six funnel shifts, six XORs and a seven-add carry chain per element) while the multiplier pipe idles.
 * qwa_fma does the same update with
 *   - the whole-word part of the alignment through a per-thread shared-memory column (6 STS + 6 LDS at a
 *     dynamic word offset: the load/store pipe instead of 18 SELs),
 *   - the bit part and the accumulation fused into seven multiply-adds  W += t_j * 2^(32-r)  that land
 *     directly on the accumulator limbs.  Adjacent 64-bit products overlap by one word, so the window is
 *     kept as two carry-save halves, W = e + v (mod 2^192): e takes the products that start on an even
 *     word, v the ones that start on an odd word; they are only added together when the accumulator
 *     leaves the loop (qwa_fold),
 *   - subtraction by complementing the aligned words (t -> -t - 1, on the multiplier pipe) plus a
 *     carry-in of 1: the complemented zero words above the product are the two's-complement sign
 *     extension, and because M is a power of two the seven complemented products sum to exactly ~F,
 *     so a negative term subtracts exactly what the positive term of the same magnitude adds
 *     (x*y - x*y cancels to zero, as the README's 1e20 + 1 - 1e20 case needs).
 * With sh = QW_SH0 + d = 32*(wq+2) + r, r in [1,32], M = 2^(32-r):  P >> sh = (T*M) >> 64 where
 * T = P >> 32*(wq+1).  The word t_0 lies entirely below the window and is dropped, so the magnitude
 * added is F = floor(P/2^sh) or one unit less: |error| < 2 units = 2^-158 of a product at the anchor. */
struct qwacc {
  uint64_t e01, e23, e45;          /* even half, words (1:0) (3:2) (5:4): 64-bit so that they live in register pairs */
  uint32_t v0;                     /* odd half: word 0, words (2:1) (4:3), word 5 */
  uint64_t v12, v34;
  uint32_t v5;
  int32_t E;
};

constexpr int QWA_COL_WORDS = 12;  /* scratch column: 6 product words + 6 zero words (written once) */

QB_HD qwacc qwa_zero()
{
  qwacc z;
  z.e01 = z.e23 = z.e45 = z.v12 = z.v34 = 0;
  z.v0 = z.v5 = 0;
  z.E = QW_EMPTY;
  return z;
}

/* W = e + v */
QB_HD qwide qwa_fold(const qwacc &a)
{
  qwide s;
  s.w0 = (uint32_t)a.e01; s.w1 = (uint32_t)(a.e01 >> 32); s.w2 = (uint32_t)a.e23; s.w3 = (uint32_t)(a.e23 >> 32);
  s.w4 = (uint32_t)a.e45; s.w5 = (uint32_t)(a.e45 >> 32);
  s.E = a.E;
  qw_addsub6(s, a.v0, (uint32_t)a.v12, (uint32_t)(a.v12 >> 32), (uint32_t)a.v34, (uint32_t)(a.v34 >> 32), a.v5, 0u);
  return s;
}

QB_HD void qwa_set(qwacc &a, const qwide &s)
{
  a.e01 = ((uint64_t)s.w1 << 32) | s.w0; a.e23 = ((uint64_t)s.w3 << 32) | s.w2; a.e45 = ((uint64_t)s.w5 << 32) | s.w4;
  a.v12 = a.v34 = 0;
  a.v0 = a.v5 = 0;
  a.E = s.E;
}

/* the zero half of a scratch column; each thread owns col[k * stride], k < QWA_COL_WORDS */
QB_HD void qwa_col_init(uint32_t *col, uint32_t stride)
{
  for (int k = 6; k < QWA_COL_WORDS; ++k) col[k * stride] = 0u;
}

/* S <- S + A*B when that is the common case (both operands normal, product not above the anchor).
 * Otherwise S is left untouched and the function returns true: the caller then runs qwa_fma_rare.
 * Branch-free, so that several independent accumulators interleave in one instruction stream. */
QB_HD bool qwa_fma(qwacc &S, const qop &A, const qop &B, uint32_t *col, uint32_t stride)
{
  const bool normal = ((uint32_t)A.e - 1u < 0x7ffeu) && ((uint32_t)B.e - 1u < 0x7ffeu);
  const int32_t d = normal ? S.E - A.e - B.e : -1;
  const bool rare = d < 0;
  uint32_t p2, p3, p4, p5, p6, p7;
  mul4x4_top6(A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, p2, p3, p4, p5, p6, p7);
  col[0] = p2; col[stride] = p3; col[2 * stride] = p4; col[3 * stride] = p5; col[4 * stride] = p6;
  col[5 * stride] = p7;
  uint32_t du = (uint32_t)d;
  du = du > 223u ? 223u : du;                      /* a declined step (d < 0) lands here too: wq = 6 reads only zero words */
  const uint32_t wq = du >> 5;
  const uint32_t r = (du & 31u) + 1u;              /* bit part of the shift, 1..32 */
  const uint32_t M = 0x80000000u >> (du & 31u);    /* 2^(32-r) */
  const uint32_t mask = rare ? 0u : (0u - (A.s ^ B.s));   /* 0 for a declined step: zero words, nothing is added */
  const uint32_t *q = col + wq * stride;
  /* aligned words, complemented when subtracting (their zero extension becomes the sign extension) */
  const uint32_t t1 = q[0] ^ mask, t2 = q[stride] ^ mask, t3 = q[2 * stride] ^ mask, t4 = q[3 * stride] ^ mask,
                 t5 = q[4 * stride] ^ mask, t6 = q[5 * stride] ^ mask;
#if defined(__CUDA_ARCH__)
  const uint32_t t1s = __funnelshift_rc(t1, 0u, r);   /* hi(t1 * M) = t1 >> r (0 when r = 32) */
  asm("{\n\t"
      ".reg .u32 cy, a0, a1, a2, a3, a4, a5, b1, b2, b3, b4;\n\t"
      "mov.b64 {a0, a1}, %0;\n\t" "mov.b64 {a2, a3}, %1;\n\t" "mov.b64 {a4, a5}, %2;\n\t"
      "mov.b64 {b1, b2}, %4;\n\t" "mov.b64 {b3, b4}, %5;\n\t"
      "add.cc.u32      cy, %13, %13;\n\t"                                           /* carry <- 1 when subtracting */
      "madc.lo.cc.u32  a0, %8, %14, a0;\n\t"   "madc.hi.cc.u32 a1, %8, %14, a1;\n\t"    /* t2*M -> (e1:e0) */
      "madc.lo.cc.u32  a2, %10, %14, a2;\n\t"  "madc.hi.cc.u32 a3, %10, %14, a3;\n\t"   /* t4*M -> (e3:e2) */
      "madc.lo.cc.u32  a4, %12, %14, a4;\n\t"  "madc.hi.u32    a5, %12, %14, a5;\n\t"   /* t6*M -> (e5:e4) */
      "add.cc.u32      %3, %3, %7;\n\t"                                                 /* hi(t1*M) -> v0 */
      "madc.lo.cc.u32  b1, %9, %14, b1;\n\t"   "madc.hi.cc.u32 b2, %9, %14, b2;\n\t"    /* t3*M -> (v2:v1) */
      "madc.lo.cc.u32  b3, %11, %14, b3;\n\t"  "madc.hi.cc.u32 b4, %11, %14, b4;\n\t"   /* t5*M -> (v4:v3) */
      "madc.lo.u32     %6, %13, %14, %6;\n\t"                                           /* lo(t7*M) -> v5, t7 = sign words */
      "mov.b64 %0, {a0, a1};\n\t" "mov.b64 %1, {a2, a3};\n\t" "mov.b64 %2, {a4, a5};\n\t"
      "mov.b64 %4, {b1, b2};\n\t" "mov.b64 %5, {b3, b4};\n\t"
      "}"
      : "+l"(S.e01), "+l"(S.e23), "+l"(S.e45), "+r"(S.v0), "+l"(S.v12), "+l"(S.v34), "+r"(S.v5)
      : "r"(t1s), "r"(t2), "r"(t3), "r"(t4), "r"(t5), "r"(t6), "r"(mask), "r"(M));
#else
  (void)r;
  const uint32_t t[8] = {0u, t1, t2, t3, t4, t5, t6, mask};
  {
    uint64_t *e[3] = {&S.e01, &S.e23, &S.e45};
    uint64_t c = mask & 1u;
    for (int j = 0; j < 3; ++j) { /* t_{2j+2} * M at words (2j, 2j+1) */
      const uint64_t pr = (uint64_t)t[2 * j + 2] * M;
      const unsigned __int128 sum = (unsigned __int128)*e[j] + pr + c;
      *e[j] = (uint64_t)sum;
      c = (uint64_t)(sum >> 64);
    }
  }
  {
    uint64_t c = (uint64_t)S.v0 + (uint32_t)(((uint64_t)t[1] * M) >> 32);
    S.v0 = (uint32_t)c; c >>= 32;
    uint64_t *v[2] = {&S.v12, &S.v34};
    for (int j = 0; j < 2; ++j) { /* t_{2j+3} * M at words (2j+1, 2j+2) */
      const uint64_t pr = (uint64_t)t[2 * j + 3] * M;
      const unsigned __int128 sum = (unsigned __int128)*v[j] + pr + c;
      *v[j] = (uint64_t)sum;
      c = (uint64_t)(sum >> 64);
    }
    S.v5 = (uint32_t)(c + S.v5 + (uint32_t)((uint64_t)t[7] * M));
  }
#endif
  return rare;
}

/* the step qwa_fma declined: through the generic window code, out of line.  Everything goes through
 * memory (w = the 8-word record {w0..w5, E, bad} of the FOLDED accumulator) so that the call does not
 * pin the hot loop's registers to the ABI's argument registers. */
QB_HD_NOINLINE void qwa_fma_rare_mem(uint32_t *w, const q128 *a, const q128 *b)
{
  qwide s;
  s.w0 = w[0]; s.w1 = w[1]; s.w2 = w[2]; s.w3 = w[3]; s.w4 = w[4]; s.w5 = w[5]; s.E = (int32_t)w[6];
  uint32_t bad = 0;
  s = qw_fma_rare(s, *a, *b, &bad);
  w[0] = s.w0; w[1] = s.w1; w[2] = s.w2; w[3] = s.w3; w[4] = s.w4; w[5] = s.w5; w[6] = (uint32_t)s.E; w[7] |= bad;
}

QB_HD void qwa_fma_rare(qwacc &S, const q128 &a, const q128 &b, uint32_t &bad)
{
  const qwide f = qwa_fold(S);
  uint32_t w[8] = {f.w0, f.w1, f.w2, f.w3, f.w4, f.w5, (uint32_t)f.E, 0u};
  q128 ab[2] = {a, b};
  qwa_fma_rare_mem(w, &ab[0], &ab[1]);
  qwide g;
  g.w0 = w[0]; g.w1 = w[1]; g.w2 = w[2]; g.w3 = w[3]; g.w4 = w[4]; g.w5 = w[5]; g.E = (int32_t)w[6];
  qwa_set(S, g);
  bad |= w[7];
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

/* ---- sum of many windows at once (the block reductions of the level-1 kernels) -------------------------------------------------
 * A tree of pairwise qw_merge costs a dependent chain of merges (each with its own re-anchoring shift): 8 in a row for 128 threads,
 * twice per kernel — 18 us of a 64 us qdot at n = 10^7.  Instead every window is shifted ONCE to the largest anchor of the group
 * (the truncation qw_merge applies to the lower-anchored operand), extended to 224 bits, and the sum is plain two's-complement
 * integer addition: associative and exact, so warps add by shuffles in any order and the result is the same bits.  Groups of up
 * to 2^16 windows cannot overflow 224 bits (each |W| < 2^189). */

/* v shifted to the anchor E >= v.E (arithmetic, floor) as 7 words; an empty window is zero.  Registers only (select network). */
QB_HD void qw_align7(const qwide &v, int32_t E, uint32_t (&o)[7])
{
  if (v.E == QW_EMPTY) {
    for (int j = 0; j < 7; ++j) o[j] = 0u;
    return;
  }
  const uint32_t sg = (uint32_t)((int32_t)v.w5 >> 31);
  uint32_t s = (uint32_t)(E - v.E);
  if (s >= 192u) {                                        /* everything below the new window: 0 or -1 (floor) */
    for (int j = 0; j < 7; ++j) o[j] = sg;
    return;
  }
  uint32_t t0 = v.w0, t1 = v.w1, t2 = v.w2, t3 = v.w3, t4 = v.w4, t5 = v.w5;
  if (s & 128u) { t0 = t4; t1 = t5; t2 = sg; t3 = sg; t4 = sg; t5 = sg; }
  if (s & 64u) { t0 = t2; t1 = t3; t2 = t4; t3 = t5; t4 = sg; t5 = sg; }
  if (s & 32u) { t0 = t1; t1 = t2; t2 = t3; t3 = t4; t4 = t5; t5 = sg; }
  o[0] = fshr(t0, t1, s); o[1] = fshr(t1, t2, s); o[2] = fshr(t2, t3, s); o[3] = fshr(t3, t4, s); o[4] = fshr(t4, t5, s);
  o[5] = fshr(t5, sg, s); o[6] = sg;
}
/* a += b (224-bit two's complement) */
QB_HD void qw7_add(uint32_t (&a)[7], const uint32_t (&b)[7])
{
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %0, %7;\n\t"
      "addc.cc.u32 %1, %1, %8;\n\t"
      "addc.cc.u32 %2, %2, %9;\n\t"
      "addc.cc.u32 %3, %3, %10;\n\t"
      "addc.cc.u32 %4, %4, %11;\n\t"
      "addc.cc.u32 %5, %5, %12;\n\t"
      "addc.u32 %6, %6, %13;"
      : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6])
      : "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]));
#else
  uint64_t c = 0;
  for (int j = 0; j < 7; ++j) { c += (uint64_t)a[j] + b[j]; a[j] = (uint32_t)c; c >>= 32; }
#endif
}
/* the 224-bit sum at anchor E back as a window: while it does not leave the carry headroom of qw_merge (|W| < 2^188) 16 bits are
 * given back at the bottom, as qw_merge does */
QB_HD qwide qw_from7(const uint32_t (&a)[7], int32_t E)
{
  qwide S = qw_zero();
  if (E == QW_EMPTY) return S;
  uint32_t w[7] = {a[0], a[1], a[2], a[3], a[4], a[5], a[6]};
  for (int it = 0; it < 3; ++it) {
    const int64_t top = (int64_t)(((uint64_t)w[6] << 32) | w[5]) >> 28;   /* bits 188.. as a signed number */
    if (top == 0 || top == -1) break;
    const uint32_t sg = (uint32_t)((int32_t)w[6] >> 31);
    for (int j = 0; j < 6; ++j) w[j] = fshr(w[j], w[j + 1], 16);
    w[6] = fshr(w[6], sg, 16);
    E += 16;
  }
  S.w0 = w[0]; S.w1 = w[1]; S.w2 = w[2]; S.w3 = w[3]; S.w4 = w[4]; S.w5 = w[5]; S.E = E;
  return S;
}
/* the sum of n windows, the way the block reductions compute it (host tests; the kernels spread the same additions over a CTA) */
QB_HD qwide qw_sum_aligned(const qwide *p, int64_t n)
{
  int32_t E = QW_EMPTY;
  for (int64_t i = 0; i < n; ++i) E = p[i].E > E ? p[i].E : E;
  uint32_t a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    uint32_t b[7];
    qw_align7(p[i], E, b);
    qw7_add(a, b);
  }
  return qw_from7(a, E);
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
  /* value = Wmag * 2^(E - 2*EOFF + SH0); R = Wmag << (64 + lz), MSB at bit 255:
   * value = R * 2^(er - QBIAS - 255)  =>  er = E - 2*EOFF + SH0 - 64 - lz + QBIAS + 255 */
  const int32_t er = S.E - 2 * QW_EOFF + (QW_SH0 - 64) - lz + QBIAS + 255;
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

/*
 * q128_chain.cuh — the hot-loop primitive: acc <- RNE(a*b + acc) on 32-bit limbs.
 *
 * Every flop of the reference goes through one correctly rounded binary128 FMA
 * (Sleef_fmaq2_u05sse2 at /root/reference/include/quadblas/algorithms/level3.hpp:84 and
 * level1.hpp:24; Sleef_fmaq1_u05 at level1.hpp:31, level2.hpp:43,79, level3.hpp:36,97).  All the
 * routines are long chains  s <- fma(a_l, b_l, s), so the accumulator is kept UNPACKED between
 * steps and only the common case is inlined:
 *
 *   fast path  : a, b normal; s normal or zero; product at most 2^30 above s; fewer than 14 bits
 *                of cancellation; result normal.  Everything is done in a 160-bit frame anchored
 *                at the accumulator (5 x u32), with the low part of the 226-bit product jammed
 *                into bit 0.  Bit-exact (see the jamming argument in DESIGN.md §q128).
 *   slow path  : anything else (zeros, subnormals, Inf/NaN, massive cancellation, product far
 *                above the accumulator, overflow/underflow) re-runs the step through the generic
 *                q_fma of q128.cuh.  Same bits, just slower, so correctness never depends on
 *                which path ran.
 *
 * Dual host/device source: tests/host/q128_host_test.cpp checks q_fma_fast == fmaq bitwise on the
 * CPU; the GPU tests re-check the nvcc build.
 */
#pragma once
#include "q128.cuh"

#if defined(QB_CHAIN_STATS) && !defined(__CUDA_ARCH__)
static long qb_chain_fast_hits = 0, qb_chain_slow_hits = 0; /* host test harness only */
#endif

namespace qb {

/* operand: raw IEEE fields, mantissa right-aligned with the implicit bit made explicit */
struct qop {
  uint32_t m0, m1, m2, m3; /* 113-bit mantissa, bit 112 = m3 bit 16 (set iff exponent field != 0) */
  int32_t e;               /* biased exponent field 0..0x7fff */
  uint32_t s;              /* sign 0/1 */
};

/* accumulator: normal value  -> m = mantissa << 15 (MSB at m3 bit 31), e in [1, 0x7ffe]
 *              zero          -> m = 0, e = 0
 *              anything else -> e = -1 and m holds the packed q128 words (handled by the slow path) */
struct qacc {
  uint32_t m0, m1, m2, m3;
  int32_t e;
  uint32_t s;
};

QB_HD qop qop_load(q128 a)
{
  qop o;
  uint32_t h = (uint32_t)(a.hi >> 32);
  o.m0 = (uint32_t)a.lo;
  o.m1 = (uint32_t)(a.lo >> 32);
  o.m2 = (uint32_t)a.hi;
  o.e = (int32_t)((h >> 16) & 0x7fff);
  o.s = h >> 31;
  o.m3 = (h & 0xffffu) | (o.e ? 0x10000u : 0u);
  return o;
}

QB_HD q128 qop_pack(const qop &o)
{
  q128 r;
  r.lo = ((uint64_t)o.m1 << 32) | o.m0;
  uint32_t h = (o.s << 31) | ((uint32_t)o.e << 16) | (o.m3 & 0xffffu);
  r.hi = ((uint64_t)h << 32) | o.m2;
  return r;
}

QB_HD qacc qacc_zero()
{
  qacc z;
  z.m0 = z.m1 = z.m2 = z.m3 = 0;
  z.e = 0;
  z.s = 0;
  return z;
}

QB_HD qacc qacc_from(q128 a)
{
  qacc r;
  uint32_t h = (uint32_t)(a.hi >> 32);
  uint32_t ef = (h >> 16) & 0x7fff;
  r.s = h >> 31;
  if (ef - 1u < 0x7ffeu) { /* normal */
    uint32_t w0 = (uint32_t)a.lo, w1 = (uint32_t)(a.lo >> 32), w2 = (uint32_t)a.hi, w3 = (h & 0xffffu) | 0x10000u;
    r.m3 = (w3 << 15) | (w2 >> 17);
    r.m2 = (w2 << 15) | (w1 >> 17);
    r.m1 = (w1 << 15) | (w0 >> 17);
    r.m0 = w0 << 15;
    r.e = (int32_t)ef;
  } else if (((a.hi & 0x7fffffffffffffffULL) | a.lo) == 0) {
    r.m0 = r.m1 = r.m2 = r.m3 = 0;
    r.e = 0;
  } else {
    r.m0 = (uint32_t)a.lo; r.m1 = (uint32_t)(a.lo >> 32); r.m2 = (uint32_t)a.hi; r.m3 = h;
    r.e = -1;
  }
  return r;
}

QB_HD q128 qacc_pack(const qacc &a)
{
  q128 r;
  if (a.e > 0) {
    uint32_t w0 = (a.m0 >> 15) | (a.m1 << 17);
    uint32_t w1 = (a.m1 >> 15) | (a.m2 << 17);
    uint32_t w2 = (a.m2 >> 15) | (a.m3 << 17);
    uint32_t w3 = ((a.m3 >> 15) & 0xffffu) | ((uint32_t)a.e << 16) | (a.s << 31);
    r.lo = ((uint64_t)w1 << 32) | w0;
    r.hi = ((uint64_t)w3 << 32) | w2;
  } else if (a.e == 0) {
    r.lo = 0;
    r.hi = (uint64_t)a.s << 63;
  } else {
    r.lo = ((uint64_t)a.m1 << 32) | a.m0;
    r.hi = ((uint64_t)a.m3 << 32) | a.m2;
  }
  return r;
}

/* ------------------------------------------------------------------ limb primitives */
QB_HD uint32_t fshr(uint32_t lo, uint32_t hi, uint32_t r) /* (hi:lo) >> r, r in [0,31] */
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, r);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (r & 31));
#endif
}
QB_HD uint32_t fshl(uint32_t lo, uint32_t hi, uint32_t r) /* high word of (hi:lo) << r, r in [0,31] */
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, r);
#else
  return (uint32_t)(((((uint64_t)hi << 32) | lo) << (r & 31)) >> 32);
#endif
}
QB_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

/* 4x4 limb product: p[0..7] = a[0..3] * b[0..3] (exact, 256-bit) */
QB_HD void mul4x4(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                  const uint32_t b0, const uint32_t b1, const uint32_t b2, const uint32_t b3, uint32_t *p)
{
#if defined(__CUDA_ARCH__)
  /* even/odd column accumulators built from mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into
   * IMAD.WIDE.U32(.X) with the carry in a predicate: 16 wide multiplies + carry catches */
  uint32_t e0, e1, e2, e3, e4, e5, e6, e7; /* even chain: columns 0,2,4,6 */
  uint32_t o1, o2, o3, o4, o5, o6, o7;     /* odd chain: columns 1,3,5 (o7 = carry sink) */
  asm("{\n\t"
      /* row 0 */
      "mul.lo.u32 %0, %15, %19;\n\t"  "mul.hi.u32 %1, %15, %19;\n\t"      /* a0*b0 -> e1:e0 */
      "mul.lo.u32 %2, %15, %21;\n\t"  "mul.hi.u32 %3, %15, %21;\n\t"      /* a0*b2 -> e3:e2 */
      "mul.lo.u32 %8, %15, %20;\n\t"  "mul.hi.u32 %9, %15, %20;\n\t"      /* a0*b1 -> o2:o1 */
      "mul.lo.u32 %10, %15, %22;\n\t" "mul.hi.u32 %11, %15, %22;\n\t"     /* a0*b3 -> o4:o3 */
      /* row 1: a1*b1, a1*b3 on the even chain; a1*b0, a1*b2 on the odd chain */
      "mad.lo.cc.u32 %2, %16, %20, %2;\n\t"  "madc.hi.cc.u32 %3, %16, %20, %3;\n\t"
      "madc.lo.cc.u32 %4, %16, %22, 0;\n\t"  "madc.hi.u32 %5, %16, %22, 0;\n\t"
      "mad.lo.cc.u32 %8, %16, %19, %8;\n\t"  "madc.hi.cc.u32 %9, %16, %19, %9;\n\t"
      "madc.lo.cc.u32 %10, %16, %21, %10;\n\t" "madc.hi.cc.u32 %11, %16, %21, %11;\n\t"
      "addc.u32 %12, 0, 0;\n\t"
      /* row 2: a2*b0, a2*b2 even; a2*b1, a2*b3 odd */
      "mad.lo.cc.u32 %2, %17, %19, %2;\n\t"  "madc.hi.cc.u32 %3, %17, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, %17, %21, %4;\n\t" "madc.hi.cc.u32 %5, %17, %21, %5;\n\t"
      "addc.u32 %6, 0, 0;\n\t"
      "mad.lo.cc.u32 %10, %17, %20, %10;\n\t" "madc.hi.cc.u32 %11, %17, %20, %11;\n\t"
      "madc.lo.cc.u32 %12, %17, %22, %12;\n\t" "madc.hi.u32 %13, %17, %22, 0;\n\t"
      /* row 3: a3*b1, a3*b3 even; a3*b0, a3*b2 odd */
      "mad.lo.cc.u32 %4, %18, %20, %4;\n\t"  "madc.hi.cc.u32 %5, %18, %20, %5;\n\t"
      "madc.lo.cc.u32 %6, %18, %22, %6;\n\t" "madc.hi.u32 %7, %18, %22, 0;\n\t"
      "mad.lo.cc.u32 %10, %18, %19, %10;\n\t" "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"
      "madc.lo.cc.u32 %12, %18, %21, %12;\n\t" "madc.hi.cc.u32 %13, %18, %21, %13;\n\t"
      "addc.u32 %14, 0, 0;\n\t"
      /* merge: even + (odd << 32) */
      "add.cc.u32 %1, %1, %8;\n\t"  "addc.cc.u32 %2, %2, %9;\n\t"  "addc.cc.u32 %3, %3, %10;\n\t"
      "addc.cc.u32 %4, %4, %11;\n\t" "addc.cc.u32 %5, %5, %12;\n\t" "addc.cc.u32 %6, %6, %13;\n\t"
      "addc.u32 %7, %7, %14;\n\t"
      "}"
      : "=&r"(e0), "=&r"(e1), "=&r"(e2), "=&r"(e3), "=&r"(e4), "=&r"(e5), "=&r"(e6), "=&r"(e7),
        "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4), "=&r"(o5), "=&r"(o6), "=&r"(o7)
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3));
  (void)o1; (void)o2; (void)o3; (void)o4; (void)o5; (void)o6; (void)o7;
  p[0] = e0; p[1] = e1; p[2] = e2; p[3] = e3; p[4] = e4; p[5] = e5; p[6] = e6; p[7] = e7;
#else
  const uint32_t a[4] = {a0, a1, a2, a3}, b[4] = {b0, b1, b2, b3};
  for (int i = 0; i < 8; ++i) p[i] = 0;
  for (int i = 0; i < 4; ++i) {
    uint64_t carry = 0;
    for (int j = 0; j < 4; ++j) {
      uint64_t t = (uint64_t)a[i] * b[j] + p[i + j] + carry;
      p[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    p[i + 4] = (uint32_t)carry;
  }
#endif
}

/* ---- multi-limb carry chains (PTX add.cc/addc.cc on the device, uint64 on the host) ---- */
/* g[0..4] = (s0..s3, 0) + (x0..x4); returns the carry out of limb 4 */
QB_HD uint32_t add5(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t x0, uint32_t x1, uint32_t x2,
                    uint32_t x3, uint32_t x4, uint32_t &g0, uint32_t &g1, uint32_t &g2, uint32_t &g3, uint32_t &g4)
{
  uint32_t co;
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %6, %10;\n\t"
      "addc.cc.u32 %1, %7, %11;\n\t"
      "addc.cc.u32 %2, %8, %12;\n\t"
      "addc.cc.u32 %3, %9, %13;\n\t"
      "addc.cc.u32 %4, %14, 0;\n\t"
      "addc.u32 %5, 0, 0;"
      : "=r"(g0), "=r"(g1), "=r"(g2), "=r"(g3), "=r"(g4), "=r"(co)
      : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(x4));
#else
  uint64_t c;
  c = (uint64_t)s0 + x0; g0 = (uint32_t)c; c >>= 32;
  c += (uint64_t)s1 + x1; g1 = (uint32_t)c; c >>= 32;
  c += (uint64_t)s2 + x2; g2 = (uint32_t)c; c >>= 32;
  c += (uint64_t)s3 + x3; g3 = (uint32_t)c; c >>= 32;
  c += (uint64_t)x4; g4 = (uint32_t)c; c >>= 32;
  co = (uint32_t)c;
#endif
  return co;
}
/* (g0..g4) += inc (0/1), no carry out by construction */
QB_HD void inc5(uint32_t &g0, uint32_t &g1, uint32_t &g2, uint32_t &g3, uint32_t &g4, uint32_t inc)
{
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %0, %5;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+r"(g0), "+r"(g1), "+r"(g2), "+r"(g3), "+r"(g4)
      : "r"(inc));
#else
  uint64_t c;
  c = (uint64_t)g0 + inc; g0 = (uint32_t)c; c >>= 32;
  c += g1; g1 = (uint32_t)c; c >>= 32;
  c += g2; g2 = (uint32_t)c; c >>= 32;
  c += g3; g3 = (uint32_t)c; c >>= 32;
  g4 += (uint32_t)c;
#endif
}
/* (n0..n3) += rc; returns carry out */
QB_HD uint32_t inc4c(uint32_t &n0, uint32_t &n1, uint32_t &n2, uint32_t &n3, uint32_t rc)
{
  uint32_t co;
#if defined(__CUDA_ARCH__)
  asm("add.cc.u32 %0, %0, %5;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.cc.u32 %3, %3, 0;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(n0), "+r"(n1), "+r"(n2), "+r"(n3), "=r"(co)
      : "r"(rc));
#else
  uint64_t c;
  c = (uint64_t)n0 + rc; n0 = (uint32_t)c; c >>= 32;
  c += n1; n1 = (uint32_t)c; c >>= 32;
  c += n2; n2 = (uint32_t)c; c >>= 32;
  c += n3; n3 = (uint32_t)c; c >>= 32;
  co = (uint32_t)c;
#endif
  return co;
}

/* The slow path: same step through the generic routine.  Kept out of line (arguments by value,
 * i.e. in registers) so that the hot loop stays small; it is also the only path that can produce
 * or consume non-normal accumulators. */
QB_HD_NOINLINE q128 q_fma_slow_packed(q128 a, q128 b, q128 c) { return q_fma(a, b, c); }

/* S <- RNE(A*B + S), bit-exact IEEE-754 binary128 FMA (Sleef_fmaq1_u05 semantics) */
QB_HD void qacc_fma(qacc &S, const qop &A, const qop &B)
{
  /* ---- product (always computed; the slow path recomputes from the packed operands) ---- */
  uint32_t p[8];
  mul4x4(A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, p);

  const bool ab_normal = ((uint32_t)A.e - 1u < 0x7ffeu) && ((uint32_t)B.e - 1u < 0x7ffeu);
  const bool s_zero = (S.e == 0);
  const int32_t ep = A.e + B.e - QBIAS;       /* exponent of the product if its MSB is bit 224 */
  const uint32_t sp = A.s ^ B.s;
  const int32_t es = s_zero ? ep : S.e;
  const uint32_t ss = s_zero ? sp : S.s;
  int32_t sh = es - ep + 97;                  /* frame bit 127 <-> accumulator MSB */
  bool bad = !ab_normal || (S.e < 0) || (sh < 67);
  sh = sh > 255 ? 255 : sh;
  const uint32_t wq = ((uint32_t)sh >> 5) - 2u;   /* extra whole words, 0..5 (garbage if bad) */
  const uint32_t r = (uint32_t)sh & 31u;

  /* ---- P >> sh into the 160-bit frame, low bits jammed ---- */
  /* p[0], p[1] always fall below the frame */
  uint32_t lost = p[0] | p[1];
  uint32_t t0, t1, t2, t3, t4, t5;
  { /* stage A: by 4 words */
    const bool w4 = (wq & 4u) != 0;
    lost |= w4 ? (p[2] | p[3] | p[4] | p[5]) : 0u;
    t0 = w4 ? p[6] : p[2]; t1 = w4 ? p[7] : p[3]; t2 = w4 ? 0u : p[4]; t3 = w4 ? 0u : p[5];
    t4 = w4 ? 0u : p[6];   t5 = w4 ? 0u : p[7];
  }
  { /* stage B: by 2 words */
    const bool w2 = (wq & 2u) != 0;
    lost |= w2 ? (t0 | t1) : 0u;
    t0 = w2 ? t2 : t0; t1 = w2 ? t3 : t1; t2 = w2 ? t4 : t2; t3 = w2 ? t5 : t3;
    t4 = w2 ? 0u : t4; t5 = w2 ? 0u : t5;
  }
  { /* stage C: by 1 word */
    const bool w1 = (wq & 1u) != 0;
    lost |= w1 ? t0 : 0u;
    t0 = w1 ? t1 : t0; t1 = w1 ? t2 : t1; t2 = w1 ? t3 : t2; t3 = w1 ? t4 : t3;
    t4 = w1 ? t5 : t4; t5 = w1 ? 0u : t5;
  }
  lost |= t0 & ~(0xffffffffu << r);
  uint32_t f0 = fshr(t0, t1, r), f1 = fshr(t1, t2, r), f2 = fshr(t2, t3, r), f3 = fshr(t3, t4, r),
           f4 = fshr(t4, t5, r);
  f0 |= (lost != 0);

  /* ---- S +- P with end-around carry (one's complement subtract) ---- */
  const bool sub = (ss != sp);
  const uint32_t mk = sub ? 0xffffffffu : 0u;
  uint32_t g0, g1, g2, g3, g4;
  const uint32_t cout = add5(S.m0, S.m1, S.m2, S.m3, f0 ^ mk, f1 ^ mk, f2 ^ mk, f3 ^ mk, f4 ^ mk, g0, g1, g2, g3, g4);
  const bool neg = sub && !cout;                /* |P| > |S| : magnitude = ~T, sign flips */
  const uint32_t nm = neg ? 0xffffffffu : 0u;
  g0 ^= nm; g1 ^= nm; g2 ^= nm; g3 ^= nm; g4 ^= nm;
  inc5(g0, g1, g2, g3, g4, sub ? cout : 0u);
  const uint32_t sr = neg ? sp : ss;

  /* ---- normalise: MSB to frame bit 159 ---- */
  const bool top0 = (g4 == 0);
  const uint32_t h4 = top0 ? g3 : g4, h3 = top0 ? g2 : g3, h2 = top0 ? g1 : g2, h1 = top0 ? g0 : g1,
                 h0 = top0 ? 0u : g0;
  const int lzw = clz32(h4);                    /* 32 if h4 == 0 -> bad */
  const int lz = lzw + (top0 ? 32 : 0);
  bad = bad || (lz > 45);                       /* > 13 bits of cancellation (or exact zero) */
  const uint32_t q = (uint32_t)lzw & 31u;
  uint32_t n3 = fshl(h3, h4, q), n2 = fshl(h2, h3, q), n1 = fshl(h1, h2, q), n0 = fshl(h0, h1, q);
  const uint32_t rest = h0 << q;
  int32_t en = es + 32 - lz;

  /* ---- round to nearest even at bit 15 of n0 ---- */
  const uint32_t lsb = (n0 >> 15) & 1u;
  n0 |= (rest != 0);
  const uint32_t rco = inc4c(n0, n1, n2, n3, 0x3fffu + lsb);
  n0 &= 0xffff8000u;
  if (rco) { n3 = 0x80000000u; en += 1; }       /* mantissa rounded up to 2^113 */
  bad = bad || ((uint32_t)en - 1u >= 0x7ffeu);  /* overflow / subnormal result */

#if defined(QB_CHAIN_STATS) && !defined(__CUDA_ARCH__)
  __atomic_fetch_add(bad ? &qb_chain_slow_hits : &qb_chain_fast_hits, 1L, __ATOMIC_RELAXED);
#endif
  if (bad) {
    S = qacc_from(q_fma_slow_packed(qop_pack(A), qop_pack(B), qacc_pack(S)));
  } else {
    S.m0 = n0; S.m1 = n1; S.m2 = n2; S.m3 = n3; S.e = en; S.s = sr;
  }
}

/* packed convenience wrapper (tests, epilogues) */
QB_HD q128 q_fma_fast(q128 a, q128 b, q128 c)
{
  qacc s = qacc_from(c);
  qacc_fma(s, qop_load(a), qop_load(b));
  return qacc_pack(s);
}

} // namespace qb

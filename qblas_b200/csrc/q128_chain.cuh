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

/* hot-path variant: the implicit bit is set unconditionally.  Only valid where a zero / subnormal
 * operand (exponent field 0) is detected from .e and handled from the PACKED value (qwa_fma). */
QB_HD qop qop_load_n(q128 a)
{
  qop o;
  uint32_t h = (uint32_t)(a.hi >> 32);
  o.m0 = (uint32_t)a.lo;
  o.m1 = (uint32_t)(a.lo >> 32);
  o.m2 = (uint32_t)a.hi;
  o.e = (int32_t)((h >> 16) & 0x7fff);
  o.s = h >> 31;
  o.m3 = (h & 0xffffu) | 0x10000u;
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

/* ---- pipe-balancing primitives -------------------------------------------------------------
 * ncu on the first version of the GEMM showed the ALU pipe (LOP3/SHF/SEL/IADD3) at 80 % and the
 * FMA pipe (IMAD) at 16 % (profiles/r1a_gemm_ncu_full.txt).  The two pipes issue in parallel, so
 * shifts and conditional complements are expressed as integer multiply-adds below: same bits,
 * but the work lands on the idle pipe. */

/* Hide a value's provenance from the compiler (no instruction is emitted).  Without it NVVM
 * strength-reduces "multiply by a power of two" back into SHF and "x * +-1 + c" into LOP3. */
QB_HD uint32_t opaque(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(x));
#endif
  return x;
}

/* conditional complement as ONE IMAD: with (sg, ad) = (1, 0) it returns x, with (0xffffffff,
 * 0xffffffff) it returns -x - 1 == ~x.  cnot_coef builds the pair from a 0/1 flag. */
struct cnot_coef { uint32_t sg, ad; };
QB_HD cnot_coef cnot_make(uint32_t flag01)
{
  cnot_coef c;
  c.ad = opaque(0u - flag01);
  c.sg = opaque(c.ad | 1u);
  return c;
}
QB_HD uint32_t cnot(uint32_t x, const cnot_coef &c) { return x * c.sg + c.ad; }

QB_HD uint32_t umulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

/* Multi-limb funnel shifts on the multiplier (IMAD.HI + IMAD, no 64-bit register pairs):
 *   right by r in [1,32], M = 2^(32-r):  out = (hi >> r) | (lo... ) = lo32(hi_limb * M) + hi32(lo_limb * M)
 *   left  by q in [0,31], M = 2^q     :  out = lo32(hi_limb * M) + hi32(lo_limb * M)
 * i.e. the same formula; the two terms never overlap, so + is |. */
QB_HD uint32_t mfunnel(uint32_t lo_limb, uint32_t hi_limb, uint32_t M) { return hi_limb * M + umulhi32(lo_limb, M); }

/* The same shifts and the conditional complement on the ALU (SHF / LOP3).  tools/exp/mb_pipes.cu measured, per SM
 * sub-partition on sm_100a: IMAD 2 clk, IMAD.HI 4 clk, SHF / LOP3 2 clk, with little overlap between the pipes - so one
 * SHF (2 clk) beats the IMAD.HI + IMAD pair (6 clk) that mfunnel costs.  QB_CHAIN_SHIFT_ON_ALU selects it. */
#ifndef QB_CHAIN_SHIFT_ON_ALU
#define QB_CHAIN_SHIFT_ON_ALU 1
#endif
/* low word of (hi:lo) >> r, r in [1, 32] */
QB_HD uint32_t afunnel_r(uint32_t lo_limb, uint32_t hi_limb, uint32_t r)
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo_limb, hi_limb, r);
#else
  return r >= 32 ? hi_limb : (uint32_t)((((uint64_t)hi_limb << 32) | lo_limb) >> r);
#endif
}
/* high word of (hi:lo) << q, q in [0, 31] */
QB_HD uint32_t afunnel_l(uint32_t lo_limb, uint32_t hi_limb, uint32_t q) { return fshl(lo_limb, hi_limb, q); }

/* number of trailing zero bits of a 113-bit mantissa (normal operands have m3 bit 16 set) */
QB_HD uint32_t qop_tz(const qop &o)
{
  uint32_t w = o.m3, off = 96;
  if (o.m2) { w = o.m2; off = 64; }
  if (o.m1) { w = o.m1; off = 32; }
  if (o.m0) { w = o.m0; off = 0; }
#if defined(__CUDA_ARCH__)
  return off + (uint32_t)(__ffs((int)w) - 1);
#else
  return off + (w ? (uint32_t)__builtin_ctz(w) : 32u);
#endif
}

/* Per-thread scratch column for the whole-word part of the alignment shift.  On the GPU it lives
 * in shared memory, laid out [word][thread] (stride = threads per CTA, a multiple of 32) so that
 * the dynamically indexed loads are bank-conflict free whatever each lane's shift is; words 6..11
 * must be zero and are never written.  This moves ~25 SEL/LOP3 per FMA from the saturated ALU pipe
 * to the idle LSU pipe. */
struct qscratch {
  uint32_t *col;
  uint32_t stride;
};

/* S <- RNE(A*B + S), bit-exact IEEE-754 binary128 FMA (Sleef_fmaq1_u05 semantics).
 * USE_SCRATCH: whole-word alignment through the scratch column + jam bit from trailing-zero counts
 * (tzsum = tz(A) + tz(B) = tz(A*B)); otherwise a select network with explicit OR tracking. */
template <bool USE_SCRATCH>
QB_HD void qacc_fma_t(qacc &S, const qop &A, const qop &B, const qscratch &scr, uint32_t tzsum)
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
  int32_t sh = es - ep + 97;                  /* P >> sh puts the product into the frame (bit 127 <-> MSB of S) */
  bool bad = !ab_normal || (S.e < 0) || (sh < 67);
  sh = sh > 255 ? 255 : sh;
  sh = sh < 67 ? 67 : sh;                     /* keep the scratch index in range on the bad path */
  /* sh = 32 * (wq + 2) + r with r in [1, 32]: the bit part is done by multiplying with 2^(32-r) */
  const uint32_t shm1 = (uint32_t)sh - 1u;
  const uint32_t wq = (shm1 >> 5) - 2u;       /* extra whole words, 0..5 */
  const uint32_t M = opaque(0x80000000u >> (shm1 & 31u)); /* 2^(32-r) */

  uint32_t t0, t1, t2, t3, t4, t5, jam;
  if (USE_SCRATCH) {
    /* whole-word part: t[j] = p[2 + wq + j] by dynamic indexing of the thread's scratch column */
    uint32_t *c = scr.col;
    const uint32_t st = scr.stride;
    c[0] = p[2]; c[st] = p[3]; c[2 * st] = p[4]; c[3 * st] = p[5]; c[4 * st] = p[6]; c[5 * st] = p[7];
    const uint32_t *q = c + wq * st;
    t0 = q[0]; t1 = q[st]; t2 = q[2 * st]; t3 = q[3 * st]; t4 = q[4 * st]; t5 = q[5 * st];
    /* bits of P below the frame are non-zero iff tz(P) < sh, and tz(a*b) = tz(a) + tz(b) */
    jam = (tzsum < (uint32_t)sh) ? 1u : 0u;
  } else {
    uint32_t lost = p[0] | p[1];
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
    lost |= t0 * M;                              /* the bits of t0 that the bit part shifts out */
    jam = (lost != 0) ? 1u : 0u;
  }
  /* ---- bit part: (t >> r) on the multiplier ---- */
#if QB_CHAIN_SHIFT_ON_ALU
  const uint32_t rr = (shm1 & 31u) + 1u;
  const uint32_t f0 = afunnel_r(t0, t1, rr) | jam, f1 = afunnel_r(t1, t2, rr), f2 = afunnel_r(t2, t3, rr), f3 = afunnel_r(t3, t4, rr),
                 f4 = afunnel_r(t4, t5, rr);
#else
  const uint32_t f0 = mfunnel(t0, t1, M) | jam, f1 = mfunnel(t1, t2, M), f2 = mfunnel(t2, t3, M), f3 = mfunnel(t3, t4, M),
                 f4 = mfunnel(t4, t5, M);
#endif

  /* ---- S +- P with end-around carry (one's complement subtract) ---- */
  const uint32_t sub = ss ^ sp;                 /* 0/1 */
  uint32_t g0, g1, g2, g3, g4;
#if QB_CHAIN_SHIFT_ON_ALU
  const uint32_t mk = 0u - sub;
  const uint32_t cout = add5(S.m0, S.m1, S.m2, S.m3, f0 ^ mk, f1 ^ mk, f2 ^ mk, f3 ^ mk, f4 ^ mk, g0, g1, g2, g3, g4);
  const uint32_t neg = sub & (cout ^ 1u);       /* |P| > |S| : magnitude = ~T, sign flips */
  const uint32_t nm = 0u - neg;
  g0 ^= nm; g1 ^= nm; g2 ^= nm; g3 ^= nm; g4 ^= nm;
#else
  const cnot_coef mk = cnot_make(sub);
  const uint32_t cout = add5(S.m0, S.m1, S.m2, S.m3, cnot(f0, mk), cnot(f1, mk), cnot(f2, mk), cnot(f3, mk), cnot(f4, mk),
                             g0, g1, g2, g3, g4);
  const uint32_t neg = sub & (cout ^ 1u);       /* |P| > |S| : magnitude = ~T, sign flips */
  const cnot_coef nm = cnot_make(neg);
  g0 = cnot(g0, nm); g1 = cnot(g1, nm); g2 = cnot(g2, nm); g3 = cnot(g3, nm); g4 = cnot(g4, nm);
#endif
  inc5(g0, g1, g2, g3, g4, sub & cout);
  const uint32_t sr = ss ^ neg;                 /* sign flips iff the magnitudes swapped (then sp == ss ^ 1) */

  /* ---- normalise: MSB to frame bit 159.  lz in [0, 45] is split as q1 + q2 (each <= 31) and both
   *      left shifts run on the multiplier ---- */
  const int lz4 = clz32(g4);
  const int lz = (g4 != 0) ? lz4 : 32 + clz32(g3);
  bad = bad || (lz > 45);                       /* > 13 bits of cancellation (or exact zero) */
  const uint32_t q1 = (uint32_t)(lz > 31 ? 31 : lz) & 31u, q2 = (uint32_t)(lz > 31 ? lz - 31 : 0) & 31u;
#if QB_CHAIN_SHIFT_ON_ALU
  const uint32_t h0 = g0 << q1, h1 = afunnel_l(g0, g1, q1), h2 = afunnel_l(g1, g2, q1), h3 = afunnel_l(g2, g3, q1), h4 = afunnel_l(g3, g4, q1);
  const uint32_t rest = h0 << q2;
  uint32_t n0 = afunnel_l(h0, h1, q2), n1 = afunnel_l(h1, h2, q2), n2 = afunnel_l(h2, h3, q2), n3 = afunnel_l(h3, h4, q2);
#else
  const uint32_t M1 = opaque(1u << q1), M2 = opaque(1u << q2);
  const uint32_t h0 = g0 * M1, h1 = mfunnel(g0, g1, M1), h2 = mfunnel(g1, g2, M1), h3 = mfunnel(g2, g3, M1),
                 h4 = mfunnel(g3, g4, M1);
  const uint32_t rest = h0 * M2;
  uint32_t n0 = mfunnel(h0, h1, M2), n1 = mfunnel(h1, h2, M2), n2 = mfunnel(h2, h3, M2), n3 = mfunnel(h3, h4, M2);
#endif
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

/* register-only form (select network) */
QB_HD void qacc_fma(qacc &S, const qop &A, const qop &B)
{
  qscratch none; none.col = nullptr; none.stride = 0;
  qacc_fma_t<false>(S, A, B, none, 0u);
}
/* scratch-column form: tzsum = qop_tz(A) + qop_tz(B) */
QB_HD void qacc_fma_sc(qacc &S, const qop &A, const qop &B, const qscratch &scr, uint32_t tzsum)
{
  qacc_fma_t<true>(S, A, B, scr, tzsum);
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * The branch-free step of the reference-order qgemm (k_gemm_nb, qb_level3.cu).
 *
 * tools/exp/mb_pipes.cu measured what a warp instruction costs one SM sub-partition of sm_100a in dispatch clocks — IMAD.WIDE 4-5,
 * IMAD / LOP3 / SHF / LDS / STS 2, IADD3 / ISETP / SEL 1 — and the sum of these over the hot path of qacc_fma_t reproduces the
 * measured 277 clocks per qFMA: the kernel is bound by that sum, not by one pipe.  This step is the same arithmetic with fewer and
 * cheaper instructions:
 *   - the operands arrive STAGED (qstaged: implicit bit set, and one meta word e | sign << 16 | zero << 18 | abnormal << 20 |
 *     tz << 23 made once per element when the tile is staged): ONE add of the two meta words gives the exponent sum, the sign of
 *     the product, tz(a) + tz(b) and the operand classes at once;
 *   - the product is added or subtracted in two's complement by predicated carry chains (IADD3, 1 clock each) and a predicated
 *     negation replaces the XOR masks, the end-around increment and the conditional complement;
 *   - normalisation is one 5-word select (the sum has its leading one in word 4 or in word 3) and ONE funnel-shift pass;
 *   - no branch and no call: a step that needs the generic routine (cancellation of more than 13 bits, a product far above the
 *     accumulator, an abnormal operand or result) leaves S untouched and returns 1; the caller redoes those steps out of line after
 *     the independent steps of one k index, so the hot loop is one basic block.
 * A zero operand is not abnormal: the product is an exact zero, tz = 255 keeps the sticky bit clear, and S comes out unchanged
 * (a zero accumulator is always +0 here — -0 is kept in the packed form — and (+0) + (+-0) = +0).
 * Same bits as q_fma for every step that returns 0 (tests/host/q128_host_test.cpp). */
struct qacc2 {
  uint32_t m0, m1, m2, m3;   /* normal: mantissa << 15; e in [1, 0x7ffe].  +0: all zero, e = 0.  Anything else: e = QACC2_PACKED, m = the packed words */
  int32_t e;
  uint32_t sw;               /* sign in bit 16 (the other bits are don't-care) */
};
struct qstaged { uint32_t m0, m1, m2, m3, meta; };

constexpr uint32_t QM_SIGN = 1u << 16, QM_ZERO = 1u << 18, QM_ABN = 1u << 20, QM_FLAGS = 0xfu << 18;
constexpr int QM_TZ = 23;
constexpr int32_t QACC2_PACKED = -0x10000;   /* so far below every exponent sum that the shift test of the step declines it */
struct qnbctx {              /* per-thread constants of the step */
  uint32_t *col;             /* scratch column (12 words at `stride`, words 6..11 zero) */
  uint32_t stride;
  uint32_t zero;             /* 0 that the compiler cannot see through (keeps the negation on IADD3.X with an inverted operand) */
};

QB_HD qstaged qstage(q128 v)
{
  const qop o = qop_load(v);
  qstaged r;
  r.m0 = o.m0; r.m1 = o.m1; r.m2 = o.m2; r.m3 = o.m3;
  r.meta = (uint32_t)o.e | (o.s << 16);
  if ((uint32_t)o.e - 1u < 0x7ffeu) r.meta |= qop_tz(o) << QM_TZ;
  else if (o.e == 0 && (o.m0 | o.m1 | o.m2 | o.m3) == 0u) r.meta |= QM_ZERO | (255u << QM_TZ);
  else r.meta |= QM_ABN;
  return r;
}
QB_HD q128 qstaged_pack(const qstaged &a)
{
  q128 r;
  r.lo = ((uint64_t)a.m1 << 32) | a.m0;
  const uint32_t h = ((a.meta & QM_SIGN) << 15) | ((a.meta & 0x7fffu) << 16) | (a.m3 & 0xffffu);
  r.hi = ((uint64_t)h << 32) | a.m2;
  return r;
}
QB_HD qacc2 qacc2_zero()
{
  qacc2 z;
  z.m0 = z.m1 = z.m2 = z.m3 = 0; z.e = 0; z.sw = 0;
  return z;
}
QB_HD qacc2 qacc2_from(q128 a)
{
  qacc2 r;
  const uint32_t h = (uint32_t)(a.hi >> 32);
  const uint32_t ef = (h >> 16) & 0x7fff;
  r.sw = h >> 15;                          /* bit 31 -> bit 16 */
  if (ef - 1u < 0x7ffeu) {
    const uint32_t w0 = (uint32_t)a.lo, w1 = (uint32_t)(a.lo >> 32), w2 = (uint32_t)a.hi, w3 = (h & 0xffffu) | 0x10000u;
    r.m3 = (w3 << 15) | (w2 >> 17);
    r.m2 = (w2 << 15) | (w1 >> 17);
    r.m1 = (w1 << 15) | (w0 >> 17);
    r.m0 = w0 << 15;
    r.e = (int32_t)ef;
  } else if ((a.hi | a.lo) == 0) {
    r.m0 = r.m1 = r.m2 = r.m3 = 0;
    r.e = 0;
  } else {
    r.m0 = (uint32_t)a.lo; r.m1 = (uint32_t)(a.lo >> 32); r.m2 = (uint32_t)a.hi; r.m3 = h;
    r.e = QACC2_PACKED;
  }
  return r;
}
QB_HD q128 qacc2_pack(const qacc2 &a)
{
  q128 r;
  if (a.e > 0) {
    const uint32_t w0 = (a.m0 >> 15) | (a.m1 << 17);
    const uint32_t w1 = (a.m1 >> 15) | (a.m2 << 17);
    const uint32_t w2 = (a.m2 >> 15) | (a.m3 << 17);
    const uint32_t w3 = ((a.m3 >> 15) & 0xffffu) | ((uint32_t)a.e << 16) | ((a.sw & QM_SIGN) << 15);
    r.lo = ((uint64_t)w1 << 32) | w0;
    r.hi = ((uint64_t)w3 << 32) | w2;
  } else if (a.e == 0) {
    r.lo = 0;
    r.hi = 0;
  } else {
    r.lo = ((uint64_t)a.m1 << 32) | a.m0;
    r.hi = ((uint64_t)a.m3 << 32) | a.m2;
  }
  return r;
}

/* (g0..g4) = (s0..s3, 0) +- (f0..f4) in two's complement (minus iff subw != 0); bo = 0xffffffff iff the difference is negative.
 * The borrow is read back inside the subtraction chain only: reading it with subc after the ADD chain (whose carry is known to be
 * clear) would save an instruction, but ptxas keeps the flag of a subtraction in the inverted sense and does not convert between the
 * two — tried, wrong results on the device. */
QB_HD void addsub5(uint32_t s0, uint32_t s1, uint32_t s2, uint32_t s3, uint32_t f0, uint32_t f1, uint32_t f2, uint32_t f3, uint32_t f4,
                   uint32_t subw, uint32_t &g0, uint32_t &g1, uint32_t &g2, uint32_t &g3, uint32_t &g4, uint32_t &bo)
{
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %15, 0;\n\t"
      "@!p add.cc.u32 %0, %6, %10;\n\t"
      "@!p addc.cc.u32 %1, %7, %11;\n\t"
      "@!p addc.cc.u32 %2, %8, %12;\n\t"
      "@!p addc.cc.u32 %3, %9, %13;\n\t"
      "@!p addc.u32 %4, %14, 0;\n\t"
      "@!p mov.u32 %5, 0;\n\t"
      "@p sub.cc.u32 %0, %6, %10;\n\t"
      "@p subc.cc.u32 %1, %7, %11;\n\t"
      "@p subc.cc.u32 %2, %8, %12;\n\t"
      "@p subc.cc.u32 %3, %9, %13;\n\t"
      "@p subc.cc.u32 %4, 0, %14;\n\t"
      "@p subc.u32 %5, 0, 0;\n\t"
      "}"
      : "=&r"(g0), "=&r"(g1), "=&r"(g2), "=&r"(g3), "=&r"(g4), "=&r"(bo)
      : "r"(s0), "r"(s1), "r"(s2), "r"(s3), "r"(f0), "r"(f1), "r"(f2), "r"(f3), "r"(f4), "r"(subw));
#else
  const uint32_t s[5] = {s0, s1, s2, s3, 0u}, f[5] = {f0, f1, f2, f3, f4};
  uint32_t g[5];
  if (subw == 0u) {
    uint64_t c = 0;
    for (int i = 0; i < 5; ++i) { c += (uint64_t)s[i] + f[i]; g[i] = (uint32_t)c; c >>= 32; }
    bo = 0u;
  } else {
    int64_t c = 0;
    for (int i = 0; i < 5; ++i) { c += (int64_t)s[i] - (int64_t)f[i]; g[i] = (uint32_t)c; c >>= 32; }
    bo = c < 0 ? 0xffffffffu : 0u;
  }
  g0 = g[0]; g1 = g[1]; g2 = g[2]; g3 = g[3]; g4 = g[4];
#endif
}
/* (g0..g4) = -(g0..g4) iff bo != 0 (z = 0 in a register) */
QB_HD void condneg5(uint32_t &g0, uint32_t &g1, uint32_t &g2, uint32_t &g3, uint32_t &g4, uint32_t bo, uint32_t z)
{
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .pred p;\n\t"
      "setp.ne.u32 p, %5, 0;\n\t"
      "@p sub.cc.u32 %0, %6, %0;\n\t"
      "@p subc.cc.u32 %1, %6, %1;\n\t"
      "@p subc.cc.u32 %2, %6, %2;\n\t"
      "@p subc.cc.u32 %3, %6, %3;\n\t"
      "@p subc.u32 %4, %6, %4;\n\t"
      "}"
      : "+r"(g0), "+r"(g1), "+r"(g2), "+r"(g3), "+r"(g4)
      : "r"(bo), "r"(z));
#else
  (void)z;
  if (bo) {
    uint32_t g[5] = {g0, g1, g2, g3, g4};
    int64_t c = 0;
    for (int i = 0; i < 5; ++i) { c -= (int64_t)g[i]; g[i] = (uint32_t)c; c >>= 32; }
    g0 = g[0]; g1 = g[1]; g2 = g[2]; g3 = g[3]; g4 = g[4];
  }
#endif
}
/* the leading one of (g4..g0) is in g4 or, if g4 = 0, in g3: move it to w4 (predicated moves: the copies may go to either pipe);
 * returns 32 when the words moved up (the whole-word part of the leading-zero count), else 0 */
QB_HD uint32_t lead5(uint32_t g0, uint32_t g1, uint32_t g2, uint32_t g3, uint32_t g4, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3,
                     uint32_t &w4)
{
#if defined(__CUDA_ARCH__)
  uint32_t up;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.eq.u32 p, %10, 0;\n\t"
      "mov.u32 %0, %6;\n\t" "mov.u32 %1, %7;\n\t" "mov.u32 %2, %8;\n\t" "mov.u32 %3, %9;\n\t" "mov.u32 %4, %10;\n\t"
      "@p mov.u32 %4, %9;\n\t" "@p mov.u32 %3, %8;\n\t" "@p mov.u32 %2, %7;\n\t" "@p mov.u32 %1, %6;\n\t" "@p mov.u32 %0, 0;\n\t"
      "selp.u32 %5, 32, 0, p;\n\t"
      "}"
      : "=&r"(w0), "=&r"(w1), "=&r"(w2), "=&r"(w3), "=&r"(w4), "=&r"(up)
      : "r"(g0), "r"(g1), "r"(g2), "r"(g3), "r"(g4));
  return up;
#else
  const bool top = (g4 != 0u);
  w4 = top ? g4 : g3; w3 = top ? g3 : g2; w2 = top ? g2 : g1; w1 = top ? g1 : g0; w0 = top ? g0 : 0u;
  return top ? 0u : 32u;
#endif
}

QB_HD uint32_t umin32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return min(a, b);
#else
  return a < b ? a : b;
#endif
}

/* The last stage of the step: the rounding increment, the carry into the exponent and the new sign are written into S by
 * PREDICATED instructions, so a declined step leaves S untouched without a select per word.  Declined (returns true) when
 *   sh_raw < 67 | lz > 45 | en outside [1, 0x7ffd] (one short of the largest exponent: the carry of the rounding is not known yet) |
 *   an abnormal operand — unless S = +0 meets an exact zero product (fl: a zero operand and no abnormal one), which leaves S = +0. */
QB_HD bool round_commit(qacc2 &S, uint32_t n0, uint32_t n1, uint32_t n2, uint32_t n3, uint32_t rc, int32_t en, uint32_t wsum, uint32_t subw,
                        uint32_t bo, int32_t sh_raw, uint32_t lz, uint32_t fl)
{
#if defined(__CUDA_ARCH__)
  uint32_t ret;
  asm("{\n\t.reg .pred p, z;\n\t.reg .u32 t, t3, c;\n\t"
      "setp.eq.s32 z, %4, 0;\n\t"                          /* S = +0 ... */
      "setp.lt.and.u32 z, %17, 0xfffff, z;\n\t"                /* ... and (fl - 1) < ABN - 1 */
      "setp.lt.s32 p, %14, 67;\n\t"
      "setp.gt.or.u32 p, %15, 45, p;\n\t"
      "setp.ge.or.u32 p, %16, 0x7ffd, p;\n\t"              /* en - 1 */
      "setp.ge.or.u32 p, %18, 0x100000, p;\n\t"                 /* fl >= ABN */
      "@!p add.cc.u32 t, %7, %11;\n\t"
      "@!p addc.cc.u32 %1, %8, 0;\n\t"
      "@!p addc.cc.u32 %2, %9, 0;\n\t"
      "@!p addc.cc.u32 t3, %10, 0;\n\t"
      "addc.u32 c, 0, 0;\n\t"                                /* unpredicated: only predicated instructions read it */
      "@!p and.b32 %0, t, 0xffff8000;\n\t"
      "@!p mad.lo.u32 %3, c, 0x80000000, t3;\n\t"          /* a mantissa that rounded up to 2^113 left 0 */
      "@!p add.s32 %4, %12, c;\n\t"
      "@!p lop3.b32 %5, %13, %19, %20, 0xb4;\n\t"          /* wsum ^ (subw & ~bo): sign(S) unless the magnitudes swapped */
      "and.pred p, p, !z;\n\t"
      "selp.u32 %6, 1, 0, p;\n\t"
      "}"
      : "+r"(S.m0), "+r"(S.m1), "+r"(S.m2), "+r"(S.m3), "+r"(S.e), "+r"(S.sw), "=r"(ret)
      : "r"(n0), "r"(n1), "r"(n2), "r"(n3), "r"(rc), "r"(en), "r"(wsum), "r"(sh_raw), "r"(lz), "r"((uint32_t)en - 1u), "r"(fl - 1u),
        "r"(fl), "r"(subw), "r"(bo));
  static_assert(QM_ABN == 0x100000u, "the literals of the asm block");
  return ret != 0u;
#else
  const bool zz = (S.e == 0) && (fl - 1u < QM_ABN - 1u);
  const bool bad = (sh_raw < 67) || (lz > 45u) || ((uint32_t)en - 1u >= 0x7ffdu) || (fl >= QM_ABN);
  if (!bad) {
    const uint32_t rco = inc4c(n0, n1, n2, n3, rc);
    S.m0 = n0 & 0xffff8000u; S.m1 = n1; S.m2 = n2; S.m3 = n3 + (rco << 31); S.e = en + (int32_t)rco; S.sw = wsum ^ (subw & ~bo);
  }
  return bad && !zz;
#endif
}

/* S <- RNE(A*B + S) for staged operands (wsum = A.meta + B.meta); returns false and updates S, or returns true and leaves S as it was */
QB_HD bool qacc_fma_nb(qacc2 &S, const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3, const uint32_t b0,
                       const uint32_t b1, const uint32_t b2, const uint32_t b3, const uint32_t wsum, const qnbctx &cx)
{
  uint32_t p[8];
  mul4x4(a0, a1, a2, a3, b0, b1, b2, b3, p);

  const int32_t esum = (int32_t)(wsum & 0xffffu);
  const bool s_zero = (S.e == 0);
  int32_t sh = S.e - esum + (QBIAS + 97);     /* es - ep + 97: P >> sh puts the product into the frame */
  sh = s_zero ? 97 : sh;
  const int32_t sh_raw = sh;                  /* < 67: product more than 2^30 above S, or S in the packed form */
  sh = sh > 255 ? 255 : sh;
  sh = sh < 67 ? 67 : sh;
  /* whole words of the shift through the scratch column (words 6..11 are zero), the bit part by funnel shifts (sh mod 32) */
  uint32_t *c = cx.col;
  const uint32_t st = cx.stride;
  c[0] = p[2]; c[st] = p[3]; c[2 * st] = p[4]; c[3 * st] = p[5]; c[4 * st] = p[6]; c[5 * st] = p[7];
  const uint32_t *q = c + (((uint32_t)sh >> 5) - 2u) * st;
  const uint32_t t0 = q[0], t1 = q[st], t2 = q[2 * st], t3 = q[3 * st], t4 = q[4 * st], t5 = q[5 * st];
  const uint32_t jam = ((wsum >> QM_TZ) < (uint32_t)sh) ? 1u : 0u;   /* bits of P below the frame: tz(a b) = tz(a) + tz(b) */
  const uint32_t r = (uint32_t)sh;
  const uint32_t f0 = fshr(t0, t1, r) | jam, f1 = fshr(t1, t2, r), f2 = fshr(t2, t3, r), f3 = fshr(t3, t4, r), f4 = fshr(t4, t5, r);

  /* S +- P, magnitude and sign */
  uint32_t subw = (S.sw ^ wsum) & QM_SIGN;
  subw = s_zero ? 0u : subw;
  uint32_t g0, g1, g2, g3, g4, bo;
  addsub5(S.m0, S.m1, S.m2, S.m3, f0, f1, f2, f3, f4, subw, g0, g1, g2, g3, g4, bo);
  condneg5(g0, g1, g2, g3, g4, bo, cx.zero);

  /* normalise: the leading one is in word 4 or, after a cancellation of up to 13 bits, in word 3 */
  uint32_t w0, w1, w2, w3, w4;
  const uint32_t up = lead5(g0, g1, g2, g3, g4, w0, w1, w2, w3, w4);
  const uint32_t lzq = (uint32_t)clz32(w4);
  const uint32_t lz = lzq + up;                          /* > 45: more than 13 bits cancelled */
  const uint32_t n3 = fshl(w3, w4, lzq), n2 = fshl(w2, w3, lzq), n1 = fshl(w1, w2, lzq);
  uint32_t n0 = fshl(w0, w1, lzq);
  const uint32_t rest = w0 << (lzq & 31u);
  const int32_t es = s_zero ? esum - QBIAS : S.e;
  const int32_t en = es + 32 - (int32_t)lz;              /* before the carry of the rounding */
  const uint32_t rc = 0x3fffu + ((n0 >> 15) & 1u);       /* round to nearest even at bit 15 of n0 */
  n0 |= umin32(rest, 1u);
  const uint32_t fl = wsum & QM_FLAGS;
  return round_commit(S, n0, n1, n2, n3, rc, en, wsum, subw, bo, sh_raw, lz, fl);
}

/* test wrapper: *bad = 1 means "not handled" and the return value is c itself */
QB_HD q128 q_fma_fast_nb(q128 a, q128 b, q128 c, uint32_t *col, uint32_t stride, uint32_t *bad)
{
  qacc2 s = qacc2_from(c);
  const qstaged A = qstage(a), B = qstage(b);
  qnbctx cx; cx.col = col; cx.stride = stride; cx.zero = 0u;
  *bad = qacc_fma_nb(s, A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, A.meta + B.meta, cx) ? 1u : 0u;
  return qacc2_pack(s);
}

/* packed convenience wrappers (tests, epilogues) */
QB_HD q128 q_fma_fast(q128 a, q128 b, q128 c)
{
  qacc s = qacc_from(c);
  qacc_fma(s, qop_load(a), qop_load(b));
  return qacc_pack(s);
}
/* scratch-column variant; `col` must have 12 words at the given stride with words 6..11 zero */
QB_HD q128 q_fma_fast_sc(q128 a, q128 b, q128 c, uint32_t *col, uint32_t stride)
{
  qacc s = qacc_from(c);
  const qop A = qop_load(a), B = qop_load(b);
  qscratch scr; scr.col = col; scr.stride = stride;
  qacc_fma_sc(s, A, B, scr, qop_tz(A) + qop_tz(B));
  return qacc_pack(s);
}

} // namespace qb

/*
 * q128.cuh — software IEEE-754 binary128 arithmetic on integer limbs (the scalar boundary).
 *
 * Replaces the reference's only arithmetic dependency, SLEEF's libsleefquad
 * (Sleef_fmaq1_u05 / Sleef_mulq1_u05 / Sleef_addq1_u05 / Sleef_sqrtq1_u05 and the double<->quad
 * casts; call sites listed in SURVEY.md §8 a15, e.g. /root/reference/include/quadblas/
 * algorithms/level1.hpp:31, level2.hpp:48, level3.hpp:84,107, interface/c_interface.hpp:30,42).
 * "_u05" = error <= 0.5 ULP = the correctly rounded round-to-nearest-even result, which is
 * unique, so these routines are specified to return exactly those bits for every finite /
 * infinite / subnormal / signed-zero input.  NaN results are canonical quiet NaNs (payloads are
 * outside the parity contract).
 *
 * Layout of a q128: little-endian IEEE binary128 = Sleef_quad = __float128:
 *   lo = mantissa bits 0..63,  hi = sign(1) | biased exponent(15) | mantissa bits 64..111 (48).
 *
 * The code is plain C++ on uint64_t so that the very same source is compiled
 *   - by nvcc for sm_100a (the product), and
 *   - by g++ inside tests/host/ (a CPU bit-exactness harness against libquadmath; tests only).
 * The generic routines here are the reference semantics; the hot kernels use the chain-form
 * primitives in q128_chain.cuh, which are tested bitwise against these.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QB_HD __host__ __device__ __forceinline__
#define QB_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define QB_HD static inline __attribute__((always_inline))
#define QB_HD_NOINLINE static __attribute__((noinline))
#endif

struct __attribute__((aligned(16))) q128 {
  uint64_t lo, hi;
};

namespace qb {

static constexpr int QBIAS = 16383;
static constexpr uint64_t Q_MANT_HI_MASK = 0x0000ffffffffffffULL;
static constexpr uint64_t Q_IMPLICIT = 0x0001000000000000ULL;
static constexpr uint64_t Q_EXP_INF_HI = 0x7fff000000000000ULL;
static constexpr uint64_t Q_QNAN_HI = 0x7fff800000000000ULL;

/* ------------------------------------------------------------------ small integer helpers */
QB_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

QB_HD void mul64x64(uint64_t a, uint64_t b, uint64_t &lo, uint64_t &hi)
{
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}

/* 256-bit little-endian frame */
struct u256 {
  uint64_t w0, w1, w2, w3;
};

QB_HD bool u256_is_zero(const u256 &a) { return (a.w0 | a.w1 | a.w2 | a.w3) == 0; }

QB_HD u256 u256_add(const u256 &a, const u256 &b)
{
  u256 r;
  uint64_t c;
  r.w0 = a.w0 + b.w0; c = r.w0 < a.w0;
  r.w1 = a.w1 + b.w1; uint64_t c1 = r.w1 < a.w1; r.w1 += c; c1 |= r.w1 < c; c = c1;
  r.w2 = a.w2 + b.w2; c1 = r.w2 < a.w2; r.w2 += c; c1 |= r.w2 < c; c = c1;
  r.w3 = a.w3 + b.w3 + c;
  return r;
}

/* r = a - b, returns borrow out */
QB_HD u256 u256_sub(const u256 &a, const u256 &b, uint64_t &borrow)
{
  u256 r;
  uint64_t bw, b1;
  r.w0 = a.w0 - b.w0; bw = a.w0 < b.w0;
  r.w1 = a.w1 - b.w1; b1 = a.w1 < b.w1; b1 |= r.w1 < bw; r.w1 -= bw; bw = b1;
  r.w2 = a.w2 - b.w2; b1 = a.w2 < b.w2; b1 |= r.w2 < bw; r.w2 -= bw; bw = b1;
  r.w3 = a.w3 - b.w3; b1 = a.w3 < b.w3; b1 |= r.w3 < bw; r.w3 -= bw; bw = b1;
  borrow = bw;
  return r;
}

QB_HD u256 u256_neg(const u256 &a)
{
  u256 z = {0, 0, 0, 0};
  uint64_t bw;
  return u256_sub(z, a, bw);
}

/* logical right shift by s in [0, 255]; bits shifted out are OR-ed ("jammed") into bit 0 */
QB_HD u256 u256_shr_jam(u256 a, uint32_t s)
{
  uint64_t lost = 0;
  if (s >= 128) { lost |= a.w0 | a.w1; a.w0 = a.w2; a.w1 = a.w3; a.w2 = 0; a.w3 = 0; s -= 128; }
  if (s >= 64)  { lost |= a.w0; a.w0 = a.w1; a.w1 = a.w2; a.w2 = a.w3; a.w3 = 0; s -= 64; }
  if (s) {
    lost |= a.w0 << (64 - s);
    a.w0 = (a.w0 >> s) | (a.w1 << (64 - s));
    a.w1 = (a.w1 >> s) | (a.w2 << (64 - s));
    a.w2 = (a.w2 >> s) | (a.w3 << (64 - s));
    a.w3 = a.w3 >> s;
  }
  a.w0 |= (lost != 0);
  return a;
}

/* left shift by s in [0, 255] */
QB_HD u256 u256_shl(u256 a, uint32_t s)
{
  if (s >= 128) { a.w3 = a.w1; a.w2 = a.w0; a.w1 = 0; a.w0 = 0; s -= 128; }
  if (s >= 64)  { a.w3 = a.w2; a.w2 = a.w1; a.w1 = a.w0; a.w0 = 0; s -= 64; }
  if (s) {
    a.w3 = (a.w3 << s) | (a.w2 >> (64 - s));
    a.w2 = (a.w2 << s) | (a.w1 >> (64 - s));
    a.w1 = (a.w1 << s) | (a.w0 >> (64 - s));
    a.w0 = a.w0 << s;
  }
  return a;
}

QB_HD int u256_clz(const u256 &a)
{
  if (a.w3) return clz64(a.w3);
  if (a.w2) return 64 + clz64(a.w2);
  if (a.w1) return 128 + clz64(a.w1);
  return 192 + clz64(a.w0);
}

/* ------------------------------------------------------------------ classification */
QB_HD bool q_is_nan(q128 a) { return ((a.hi & 0x7fffffffffffffffULL) > Q_EXP_INF_HI) || (((a.hi & 0x7fffffffffffffffULL) == Q_EXP_INF_HI) && a.lo != 0); }
QB_HD bool q_is_inf(q128 a) { return ((a.hi & 0x7fffffffffffffffULL) == Q_EXP_INF_HI) && a.lo == 0; }
QB_HD bool q_is_zero(q128 a) { return ((a.hi & 0x7fffffffffffffffULL) | a.lo) == 0; }
QB_HD q128 q_make(uint64_t hi, uint64_t lo) { q128 r; r.lo = lo; r.hi = hi; return r; }
QB_HD q128 q_nan() { return q_make(Q_QNAN_HI, 0); }
QB_HD q128 q_inf(uint32_t sign) { return q_make(((uint64_t)sign << 63) | Q_EXP_INF_HI, 0); }
QB_HD q128 q_zero(uint32_t sign) { return q_make((uint64_t)sign << 63, 0); }
QB_HD q128 q_one() { return q_make(0x3fff000000000000ULL, 0); }
QB_HD q128 q_neg(q128 a) { a.hi ^= 0x8000000000000000ULL; return a; }
QB_HD q128 q_abs(q128 a) { a.hi &= 0x7fffffffffffffffULL; return a; }

/* Unpacked finite value: (-1)^s * m * 2^(e - QBIAS - 112), m = mh:ml (113 bits, bit 112 set unless
 * the value is zero), e may be <= 0 for normalised subnormals. */
struct qunp {
  uint64_t mh, ml;
  int32_t e;
  uint32_t s;
};

/* unpack a finite q128; subnormals are normalised (e goes <= 0); zero -> m = 0 */
QB_HD qunp q_unpack_finite(q128 a)
{
  qunp u;
  u.s = (uint32_t)(a.hi >> 63);
  u.e = (int32_t)((a.hi >> 48) & 0x7fff);
  u.mh = a.hi & Q_MANT_HI_MASK;
  u.ml = a.lo;
  if (u.e != 0) {
    u.mh |= Q_IMPLICIT;
  } else if (u.mh | u.ml) {
    /* subnormal: value = m * 2^(1 - QBIAS - 112); bring the MSB to bit 112 */
    int lz = u.mh ? clz64(u.mh) - 15 : 49 + clz64(u.ml); /* shift needed */
    if (lz >= 64) { u.mh = u.ml << (lz - 64); u.ml = 0; }
    else { u.mh = (u.mh << lz) | (u.ml >> (64 - lz)); u.ml <<= lz; }
    u.e = 1 - lz;
  }
  return u;
}

/* Round-and-pack a 256-bit magnitude R (non-zero, MSB at bit 255 after the caller normalised it)
 * whose value is R * 2^(er - QBIAS - 255); sticky information below bit 0 is already jammed into
 * bit 0.  Handles overflow to Inf and gradual underflow.  RNE. */
QB_HD q128 q_round_pack(uint32_t sign, int32_t er, u256 R)
{
  if (er >= 0x7fff) return q_inf(sign);
  int32_t ebase = er - 1;
  if (er <= 0) {
    uint32_t sh = (uint32_t)(1 - er);
    if (sh > 255) sh = 255; /* MSB at 255 shifted by 255 leaves only bit 0: pure sticky */
    R = u256_shr_jam(R, sh);
    ebase = 0;
  }
  /* mantissa = bits 255..143, guard = bit 142, sticky = bits 141..0 */
  uint64_t mh = R.w3 >> 15;                       /* 49 bits */
  uint64_t ml = (R.w3 << 49) | (R.w2 >> 15);      /* 64 bits */
  uint64_t guard = (R.w2 >> 14) & 1;
  uint64_t sticky = ((R.w2 & 0x3fffULL) | R.w1 | R.w0) != 0;
  uint64_t inc = guard & (sticky | (ml & 1));
  /* bits = (ebase << 112) + m113 + inc : the implicit bit bumps the exponent field by one */
  uint64_t hi = ((uint64_t)ebase << 48) + mh;
  uint64_t lo = ml + inc;
  hi += (lo < inc);
  return q_make(((uint64_t)sign << 63) | hi, lo);
}

/* ------------------------------------------------------------------ fused multiply-add */
/* Correctly rounded a*b + c (single rounding, RNE).  Replaces Sleef_fmaq1_u05. */
QB_HD q128 q_fma(q128 a, q128 b, q128 c)
{
  const uint32_t ea0 = (uint32_t)(a.hi >> 48) & 0x7fff;
  const uint32_t eb0 = (uint32_t)(b.hi >> 48) & 0x7fff;
  const uint32_t ec0 = (uint32_t)(c.hi >> 48) & 0x7fff;
  const uint32_t sp = (uint32_t)((a.hi ^ b.hi) >> 63);
  const uint32_t sc = (uint32_t)(c.hi >> 63);

  if (ea0 == 0x7fff || eb0 == 0x7fff || ec0 == 0x7fff) {
    if (q_is_nan(a) || q_is_nan(b) || q_is_nan(c)) return q_nan();
    if (ea0 == 0x7fff || eb0 == 0x7fff) {           /* a or b infinite */
      if (q_is_zero(a) || q_is_zero(b)) return q_nan(); /* Inf * 0 */
      if (ec0 == 0x7fff && sc != sp) return q_nan();    /* Inf - Inf */
      return q_inf(sp);
    }
    return c; /* c infinite, product finite */
  }

  qunp ua = q_unpack_finite(a), ub = q_unpack_finite(b), uc = q_unpack_finite(c);

  if ((ua.mh | ua.ml) == 0 || (ub.mh | ub.ml) == 0) { /* exact zero product */
    if ((uc.mh | uc.ml) == 0) return q_zero(sp == sc ? sc : 0);
    return c;
  }

  /* 113 x 113 -> 226-bit product P = p3:p2:p1:p0, in [2^224, 2^226) */
  uint64_t p0, p1, p2, p3;
  {
    uint64_t l0, h0, l1, h1, l2, h2, l3, h3;
    mul64x64(ua.ml, ub.ml, l0, h0);
    mul64x64(ua.mh, ub.ml, l1, h1);
    mul64x64(ua.ml, ub.mh, l2, h2);
    mul64x64(ua.mh, ub.mh, l3, h3);
    p0 = l0;
    uint64_t t = h0 + l1; uint64_t cy = t < h0;
    p1 = t + l2; cy += p1 < t;
    t = h1 + h2;                      /* h1,h2 < 2^49: no overflow */
    uint64_t t2 = t + l3; uint64_t cy2 = t2 < t;
    p2 = t2 + cy; cy2 += p2 < cy;
    p3 = h3 + cy2;
  }
  /* frame of the product: P << 29, reference position (MSB of a full-size product) = bit 254 */
  u256 FP;
  FP.w0 = p0 << 29;
  FP.w1 = (p1 << 29) | (p0 >> 35);
  FP.w2 = (p2 << 29) | (p1 >> 35);
  FP.w3 = (p3 << 29) | (p2 >> 35);
  int32_t Ep = ua.e + ub.e - QBIAS + 1;

  /* frame of the addend: mc << 142 (MSB at bit 254) */
  u256 FC;
  FC.w0 = 0; FC.w1 = 0;
  FC.w2 = uc.ml << 14;
  FC.w3 = (uc.mh << 14) | (uc.ml >> 50);
  const bool c_zero = (uc.mh | uc.ml) == 0;
  int32_t Ec = c_zero ? -(1 << 24) : uc.e;

  int32_t d = Ec - Ep;
  u256 X, Y;
  int32_t E;
  uint32_t sX, sY;
  if (d >= 0) {
    X = FC; sX = sc; sY = sp; E = Ec;
    Y = u256_shr_jam(FP, d > 255 ? 255u : (uint32_t)d); /* FP < 2^255: shift 255 leaves jam only */
  } else {
    X = FP; sX = sp; sY = sc; E = Ep;
    uint32_t nd = (uint32_t)(-d);
    Y = u256_shr_jam(FC, nd > 255 ? 255u : nd);
  }

  u256 R;
  uint32_t sr = sX;
  if (sX == sY) {
    R = u256_add(X, Y);
  } else {
    uint64_t borrow;
    R = u256_sub(X, Y, borrow);
    if (borrow) { R = u256_neg(R); sr = sY; }
    if (u256_is_zero(R)) return q_zero(0); /* exact cancellation: +0 under RNE */
  }
  int lz = u256_clz(R);
  R = u256_shl(R, (uint32_t)lz);
  return q_round_pack(sr, E + 1 - lz, R);
}

/* Sleef_mulq1_u05 == fma(a, b, -0) and Sleef_addq1_u05 == fma(a, 1, b) bit-for-bit, signed
 * zeros included.  These are epilogue-only in every routine. */
QB_HD q128 q_mul(q128 a, q128 b) { return q_fma(a, b, q_zero(1)); }
QB_HD q128 q_add(q128 a, q128 b) { return q_fma(a, q_one(), b); }
QB_HD q128 q_sub(q128 a, q128 b) { return q_fma(a, q_one(), q_neg(b)); }

/* ------------------------------------------------------------------ square root */
/* Correctly rounded sqrt (RNE).  Replaces Sleef_sqrtq1_u05 (c_interface.hpp:42, cpp_classes.hpp:80).
 * One call per qnrm2, so a plain restoring integer square root is fine. */
QB_HD_NOINLINE q128 q_sqrt(q128 a)
{
  if (q_is_nan(a)) return q_nan();
  if (q_is_zero(a)) return a;              /* sqrt(+-0) = +-0 */
  if (a.hi >> 63) return q_nan();          /* negative */
  if (q_is_inf(a)) return a;
  qunp u = q_unpack_finite(a);
  int32_t eu = u.e - QBIAS;                /* value = m * 2^(eu - 112), m in [2^112, 2^113) */
  u256 N;                                  /* N = m << (114 + odd) so that isqrt(N) has 114 bits */
  N.w0 = 0; N.w1 = 0; N.w2 = 0; N.w3 = 0;
  u256 M; M.w0 = u.ml; M.w1 = u.mh; M.w2 = 0; M.w3 = 0;
  uint32_t odd = (uint32_t)(eu & 1);
  N = u256_shl(M, 114 + odd);
  int32_t ehalf = (eu - (int32_t)odd) / 2; /* exact: eu - odd is even (works for negatives) */
  /* restoring sqrt: res accumulates the root, bit walks down the powers of 4 */
  u256 res = {0, 0, 0, 0};
  u256 bit = {0, 0, 0, 0};
  bit.w3 = 1ULL << 34;                     /* 2^226 <= N < 2^228 */
  for (int i = 0; i < 114; ++i) {
    u256 t = u256_add(res, bit);
    uint64_t bw;
    u256 dlt = u256_sub(N, t, bw);
    /* res >>= 1 */
    res.w0 = (res.w0 >> 1) | (res.w1 << 63);
    res.w1 = (res.w1 >> 1) | (res.w2 << 63);
    res.w2 = (res.w2 >> 1) | (res.w3 << 63);
    res.w3 >>= 1;
    if (!bw) { N = dlt; res = u256_add(res, bit); }
    /* bit >>= 2 */
    bit.w0 = (bit.w0 >> 2) | (bit.w1 << 62);
    bit.w1 = (bit.w1 >> 2) | (bit.w2 << 62);
    bit.w2 = (bit.w2 >> 2) | (bit.w3 << 62);
    bit.w3 >>= 2;
  }
  /* res = floor(sqrt(N)) in [2^113, 2^114): 113 mantissa bits + 1 guard; remainder N != 0 -> sticky */
  u256 R = u256_shl(res, 255 - 113);       /* MSB to bit 255 */
  R.w0 |= !u256_is_zero(N);
  /* value = res * 2^(ehalf - 113) = R * 2^(ehalf - 255) -> er = ehalf + QBIAS */
  return q_round_pack(0, ehalf + QBIAS, R);
}

/* ------------------------------------------------------------------ double <-> quad */
/* Exact widening; replaces Sleef_cast_from_doubleq1 (c_interface.hpp:54,73-74,104-105). */
QB_HD q128 q_from_double_bits(uint64_t d)
{
  uint64_t sign = d & 0x8000000000000000ULL;
  uint32_t e = (uint32_t)(d >> 52) & 0x7ff;
  uint64_t m = d & 0x000fffffffffffffULL;
  if (e == 0x7ff) return q_make(sign | Q_EXP_INF_HI | (m >> 4) | (m ? 0x0000800000000000ULL : 0), m << 60);
  if (e == 0) {
    if (m == 0) return q_make(sign, 0);
    int lz = clz64(m) - 11;                /* bring MSB to bit 52 */
    m <<= lz;
    e = (uint32_t)(1 - lz);
    m &= 0x000fffffffffffffULL;
    return q_make(sign | ((uint64_t)((int32_t)e - 1023 + QBIAS) << 48) | (m >> 4), m << 60);
  }
  return q_make(sign | ((uint64_t)(e - 1023 + QBIAS) << 48) | (m >> 4), m << 60);
}

/* RNE narrowing; replaces Sleef_cast_to_doubleq1 (c_interface.hpp:30,43). */
QB_HD uint64_t q_to_double_bits(q128 a)
{
  uint64_t sign = a.hi & 0x8000000000000000ULL;
  uint32_t e = (uint32_t)(a.hi >> 48) & 0x7fff;
  uint64_t mh = a.hi & Q_MANT_HI_MASK, ml = a.lo;
  if (e == 0x7fff) {
    if (mh | ml) return sign | 0x7ff8000000000000ULL | (mh << 4) | (ml >> 60);
    return sign | 0x7ff0000000000000ULL;
  }
  if (e == 0) return sign; /* quad subnormals/zero are far below double's range */
  /* 113-bit mantissa -> 53 bits: keep top 53, guard, sticky */
  uint64_t m = ((mh | Q_IMPLICIT) << 4) | (ml >> 60);      /* 53 bits */
  uint64_t rest = ml << 4;                                  /* 60 remaining bits, left aligned */
  int32_t ed = (int32_t)e - QBIAS + 1023;
  if (ed >= 0x7ff) return sign | 0x7ff0000000000000ULL;
  if (ed <= 0) {
    /* double subnormal (or underflow to zero): shift (m:rest) right by sh = 1 - ed, keep sticky */
    int sh = 1 - ed;
    if (sh > 53) return sign;                              /* < half of the smallest subnormal */
    uint64_t lost = rest << (64 - sh);
    rest = (m << (64 - sh)) | (rest >> sh);
    m >>= sh;
    uint64_t guard = rest >> 63;
    uint64_t sticky = ((rest << 1) | lost) != 0;
    uint64_t inc = guard & (sticky | (m & 1));
    return sign | (m + inc); /* a carry into the exponent field yields the min normal: correct */
  }
  uint64_t guard = rest >> 63;
  uint64_t sticky = (rest << 1) != 0;
  uint64_t inc = guard & (sticky | (m & 1));
  uint64_t bits = ((uint64_t)(ed - 1) << 52) + m + inc;     /* implicit bit bumps exponent */
  return sign | bits;
}

} // namespace qb

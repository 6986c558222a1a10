/*
 * qb_crt.cuh — residue (Chinese-remainder) arithmetic of the fast-mode tensor-core qgemm.
 *
 * What it is for: QuadBLAS::gemm (/root/reference/include/quadblas/algorithms/level3.hpp:215-336) in
 * QB_MODE_FAST.  Rows of op(A) and columns of op(B) are block fixed point, x = X * 2^(base - 16495)
 * with exact integers |X| < 2^W (qb_ozaki.cu, scan/plan), so every inner product is an exact integer
 *     I_ij = sum_l XA_il XB_lj,   |I| < k 2^(W_A + W_B).
 * The digit-diagonal scheme of qb_ozaki.cu spends S_A*S_B (or the ~S^2/2 leading) int8 GEMMs on it.
 * Here the integer is computed modulo N pairwise coprime moduli p_i <= 256 instead:
 *     residues : a_i = XA mod p_i, b_i = XB mod p_i as symmetric int8 (one byte plane per modulus)
 *     mma      : R_i = (a_i b_i^T) mod p_i         -- ONE int8 GEMM per modulus (N, not S_A*S_B)
 *     fold     : I = CRT(R_0 .. R_{N-1}) in (-P/2, P/2), P = prod p_i > 2 k 2^(W_A+W_B), exact,
 *                then one correctly rounded conversion to binary128.
 * Full 113-bit mantissas with the exponent spread of U(-1,1) rows (W ~ 139) at k = 8192 need
 * P > 2^292: 41 moduli, against 324 digit-plane products (all diagonals) or 136 (16 diagonals).
 * (The modular formulation of exact int8 matrix products is the published "Ozaki scheme II" idea; the
 * arithmetic below - dp4a residues, grouped two-level reconstruction - is this library's own.)
 *
 * Host/device dual source: nvcc compiles it into qb_ozaki.cu, g++ compiles the same functions in
 * tests/host/crt_host.cpp for the CPU checks against Python integers (tests/test_host_crt.py).
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include "q128.cuh"

#if defined(__CUDACC__)
#define QCRT_HD __host__ __device__ __forceinline__
#else
#define QCRT_HD static inline __attribute__((always_inline))
#endif
#if defined(__CUDA_ARCH__)
#define QCRT_UNROLL _Pragma("unroll")
#else
#define QCRT_UNROLL
#endif

namespace qb {
namespace crt {

static constexpr int NM = 49;       /* pairwise coprime moduli <= 256, largest first (greedy) */
static constexpr int NMP = 52;      /* padded to whole groups of 4 */
static constexpr int NGMAX = 13;    /* groups of <= 4 moduli, product of a group < 2^32 */
static constexpr int NLMAX = 14;    /* 32-bit limbs: NG + 1 */
static constexpr int NWMAX = 6;     /* 32-bit words of |X| (W <= 192) */
static constexpr int WMAX = 32 * NWMAX;

static const int MODULI[NM] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191,
                               181, 179, 173, 167, 163, 157, 151, 149, 139, 137, 131, 127, 113, 109, 107, 103, 101,
                               97,  89,  83,  79,  73,  71,  67,  61,  59,  53,  47,  43,  41,  37,  29};

/* per-modulus constants, the same for every plan (device copy lives in __constant__ memory) */
struct Tables {
  uint32_t p[NMP];      /* modulus (padding entries: 1) */
  uint32_t half[NMP];   /* (p-1)/2, 128 for p = 256: symmetric residues are r - half for r = (x + half) mod p */
  uint32_t cneg[NMP];   /* p*ceil(2^21/p) + 2*half: v_neg = cneg - v_pos is == -x + half (mod p) and stays >= 0 */
  uint32_t minv[NMP];   /* ceil(2^32/p): umulhi(v, minv) == floor(v/p) exactly for v < 2^22 */
  uint32_t off[NMP];    /* p*ceil(2^30/p): makes an int32 accumulator (|v| <= 2^30) non-negative, == 0 mod p */
  uint32_t finv[NMP];   /* floor(2^32/p): umulhi(u, finv) in {floor(u/p)-1, floor(u/p)} for u < 2^32 */
  uint32_t pw[NWMAX][NMP]; /* byte b of pw[j][i] = 256^(4j+b) mod p_i  (dp4a against the words of |X|) */
  uint32_t seedn[NWMAX][NMP]; /* seed of a negative element with NW = index + 1 words: half + ((-2^(32 NW)) mod p) */
};

/* per-N constants of the reconstruction (kernel parameter) */
struct Plan {
  int N, NG;
  uint32_t G[NGMAX];          /* group moduli G_g = product of the group's p */
  uint32_t flo[NGMAX], fhi[NGMAX]; /* f_g = floor(2^64 / G_g) */
  uint32_t gshr[NGMAX], gshl[NGMAX], ginv[NGMAX]; /* device quotient: G' = G << sh in [2^31, 2^32); a = (v << sh) >> 11 = (v >> gshr) << gshl,
                                                     ginv = floor((2^63 - 1) / G'); floor(a ginv / 2^52) is floor(v / G) or one less */
  uint32_t cp[NMP];           /* c'_j: t_g = (sum_{j in g} r_j c'_j) mod G_g is the top-level CRT digit, already
                                 multiplied by (P/G_g)^-1 mod G_g */
  uint32_t M[NGMAX][NLMAX];   /* M_g = P / G_g */
  uint32_t P[NLMAX], hP[NLMAX]; /* P and floor(P/2) */
  double bits;                /* log2(P) */
};

/* ------------------------------------------------------------------ portable primitives */
QCRT_HD uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, c);
#else
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
  return c;
#endif
}
QCRT_HD uint32_t mulhi_u(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
QCRT_HD uint64_t mul64hi_u(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

/* ------------------------------------------------------------------ residues */
/* Residue of (-1)^sign * X modulo p_i as r in [0, p) with r - half_i the symmetric int8 residue; X = sum_{j<NW} w[j] 2^(32 j).
 * The caller passes the words of |X| for sign = 0 and of 2^(32 NW) - |X| (two's complement, once per element) for sign = 1;
 * the seed then carries c2 = (-2^(32 NW)) mod p_i, so that no per-modulus negation is needed.  Branch-free:
 * acc <= half + p + 24*255*255 < 2^21, and umulhi(acc, ceil(2^32/p)) is the exact quotient for acc < 2^22 (Tables). */
template <int NW>
QCRT_HD uint32_t residue_sym(const uint32_t (&w)[NWMAX], uint32_t sign, int i, const Tables &T)
{
  uint32_t acc = sign ? T.seedn[NW - 1][i] : T.half[i];
QCRT_UNROLL
  for (int j = 0; j < NW; ++j) acc = dp4a_u(w[j], T.pw[j][i], acc);
  const uint32_t q = mulhi_u(acc, T.minv[i]);
  return acc - (q * T.p[i] + T.half[i]);        /* symmetric residue in [-half, p-1-half]; the low byte is the int8 */
}
/* in place: the NW words of 2^(32 NW) - X */
template <int NW>
QCRT_HD void negate_words(uint32_t (&w)[NWMAX])
{
  uint32_t carry = 1;
QCRT_UNROLL
  for (int j = 0; j < NW; ++j) { const uint32_t t = ~w[j] + carry; carry = (carry && t == 0) ? 1u : 0u; w[j] = t; }
}
/* int8 residue (as the low byte) of (-1)^sign * X from the words of |X| (one element at a time: tests, reference form) */
template <int NW>
QCRT_HD uint32_t residue_byte(const uint32_t (&w)[NWMAX], uint32_t sign, int i, const Tables &T)
{
  uint32_t v[NWMAX];
QCRT_UNROLL
  for (int j = 0; j < NWMAX; ++j) v[j] = w[j];
  bool zero = true;
QCRT_UNROLL
  for (int j = 0; j < NW; ++j) zero = zero && v[j] == 0;
  if (sign && !zero) negate_words<NW>(v);
  return residue_sym<NW>(v, (sign && !zero) ? 1u : 0u, i, T) & 0xffu;
}

/* ------------------------------------------------------------------ element -> integer words */
/* logical right shift without the sticky jam of u256_shr_jam: truncation toward zero of the magnitude */
QCRT_HD u256 u256_shr_trunc(u256 a, uint32_t s)
{
  if (s >= 256) { a.w0 = a.w1 = a.w2 = a.w3 = 0; return a; }
  if (s >= 128) { a.w0 = a.w2; a.w1 = a.w3; a.w2 = 0; a.w3 = 0; s -= 128; }
  if (s >= 64)  { a.w0 = a.w1; a.w1 = a.w2; a.w2 = a.w3; a.w3 = 0; s -= 64; }
  if (s) {
    a.w0 = (a.w0 >> s) | (a.w1 << (64 - s));
    a.w1 = (a.w1 >> s) | (a.w2 << (64 - s));
    a.w2 = (a.w2 >> s) | (a.w3 << (64 - s));
    a.w3 = a.w3 >> s;
  }
  return a;
}
/* The element as the words the residue kernels consume (crt_load4 in qb_ozaki.cu calls this; tests/host/crt_host.cpp runs the
 * same source on the CPU).  x = (-1)^s M 2^(ee - 16495) (M < 2^113, ee = biased exponent, 1 for subnormals) and
 * base = emax(row) + 113 - W, so X = trunc(M 2^(ee - base)) < 2^W.  When W covers the row's whole bit span (base <= the lowest set
 * bit of the row) X is exact; when the planner capped W, the bits of small elements below 2^base are dropped (truncation toward
 * zero, |x / 2^(base-16495) - X| < 1) and the reconstruction kernel checks every C element against the dropped mass (k_crt_fold).
 * Words: |X| for s = 0, 2^(32 NW) - |X| for s = 1 and X != 0 (residue_sym<NW>).  Zeros, elements truncated to zero and Inf/NaN
 * (their rows go to the fix-up kernel) give all-zero words and sign 0. */
template <int NW>
QCRT_HD void element_words(q128 a, int base, uint32_t (&w)[NWMAX], uint32_t &sign)
{
QCRT_UNROLL
  for (int j = 0; j < NWMAX; ++j) w[j] = 0;
  sign = 0;
  const uint32_t ef = (uint32_t)(a.hi >> 48) & 0x7fffu;
  const uint64_t mhi = (a.hi & Q_MANT_HI_MASK) | (ef ? Q_IMPLICIT : 0);
  if (!(a.lo | mhi) || ef == 0x7fffu) return;
  const int ee = ef ? (int)ef : 1;
  u256 v; v.w0 = a.lo; v.w1 = mhi; v.w2 = 0; v.w3 = 0;
  const int sh = ee - base;
  v = sh >= 0 ? u256_shl(v, (uint32_t)sh) : u256_shr_trunc(v, (uint32_t)(-sh));
  if (!(v.w0 | v.w1 | v.w2)) return;            /* truncated to zero: a plain zero (sign 0) */
  w[0] = (uint32_t)v.w0; w[1] = (uint32_t)(v.w0 >> 32);
  w[2] = (uint32_t)v.w1; w[3] = (uint32_t)(v.w1 >> 32);
  w[4] = (uint32_t)v.w2; w[5] = (uint32_t)(v.w2 >> 32);
  sign = (uint32_t)(a.hi >> 63);
  if (sign) negate_words<NW>(w);
}

/* what the tensor kernel's epilogue does with one int32 accumulator v (|v| <= 2^30): v mod p_i in [0, p) */
QCRT_HD uint32_t acc_mod(int32_t v, int i, const Tables &T)
{
  const uint32_t u = (uint32_t)v + T.off[i];
  uint32_t r = u - mulhi_u(u, T.finv[i]) * T.p[i];
  if (r >= T.p[i]) r -= T.p[i];
  return r;
}

/* ------------------------------------------------------------------ capped windows: the per-element acceptance test */
/* When the planner caps a window below a row's bit span (element_words above), every dropped tail is < 1 unit of the row's
 * scale, so the integer I^ the moduli reconstruct differs from the exact scaled inner product I by
 *     |I - I^| < E = tA sum_l |XB_lj| + tB sum_l |XA_il| + tA tB k  <=  k (tA 2^WB + tB 2^WA + tA tB)
 * (tA / tB = 1 when row i of A / column j of B lost bits).  The rounded result c^ = RNE(I^) scale then has
 * |c^ - c| <= (u |I^| + E) scale while (|A||B|)_ij >= |c| >= (|I^| - E) scale, so the fast-mode contract
 * |c^ - c| <= k u (|A||B|)_ij holds as soon as |I^| >= E (1 + k u) / ((k - 1) u); with k / (k - 1) <= 2 that is implied by
 * |I^| >= 2^T, T = WB + 115 (only A truncated), WA + 115 (only B), max(WA, WB) + 116 (both).  Returns T (the lowest
 * acceptable position of the leading bit of |I^|), -1 (every value passes, zero included) when nothing was truncated, and an
 * unreachable value for k < 2
 * (a single product must be correctly rounded: every truncated element goes to the fix-up). */
QCRT_HD int accept_msb(bool tA, bool tB, int WA, int WB, long long k)
{
  if (!tA && !tB) return -1;
  if (k < 2) return 1 << 20;
  if (tA && tB) return (WA > WB ? WA : WB) + 116;
  return tA ? WB + 115 : WA + 115;
}
/* position of the leading bit of an NL-limb magnitude, -1 for zero */
template <int NL>
QCRT_HD int limbs_msb(const uint32_t (&L)[NL])
{
  int top = -1; uint32_t tv = 1;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) if (L[l]) { top = l; tv = L[l]; }
  if (top < 0) return -1;
#if defined(__CUDA_ARCH__)
  return 32 * top + 31 - __clz((int)tv);
#else
  return 32 * top + 31 - __builtin_clz(tv);
#endif
}

/* ------------------------------------------------------------------ reconstruction */
/* a > b, or a >= b when or_equal */
template <int NL>
QCRT_HD bool limbs_cmp(const uint32_t (&a)[NL], const uint32_t *b, bool or_equal)
{
  bool res = or_equal;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l)
    if (a[l] != b[l]) res = a[l] > b[l];
  return res;
}

/* r[j] in [0, p_j) for j < N (entries j >= N ignored)  ->  |I| as NG+1 limbs and its sign, where
 * I == r_j (mod p_j) and |I| < P/2. */
template <int NG>
QCRT_HD void reconstruct(const uint32_t (&r)[NMP], const Plan &pl, uint32_t (&Y)[NG + 1], uint32_t &neg)
{
  constexpr int NL = NG + 1;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) Y[l] = 0;
  uint64_t s_lo = 0; uint32_t s_hi = 0;     /* S = sum t_g f_g  (96 bits); floor(S / 2^64) estimates floor(X / P) */
QCRT_UNROLL
  for (int g = 0; g < NG; ++g) {
    uint64_t v = 0;
QCRT_UNROLL
    for (int b = 0; b < 4; ++b)
      if (4 * g + b < pl.N) v += (uint64_t)r[4 * g + b] * pl.cp[4 * g + b];
    const uint64_t f = ((uint64_t)pl.fhi[g] << 32) | pl.flo[g];
    uint64_t t64 = v - mul64hi_u(v, f) * pl.G[g];
    if (t64 >= pl.G[g]) t64 -= pl.G[g];
    const uint32_t t = (uint32_t)t64;
    /* X += t * M_g */
    uint32_t carry = 0;
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) {
      const uint64_t a = (uint64_t)t * pl.M[g][l] + Y[l] + carry;
      Y[l] = (uint32_t)a; carry = (uint32_t)(a >> 32);
    }
    /* S += t * f */
    const uint64_t lo = (uint64_t)t * pl.flo[g], hi = (uint64_t)t * pl.fhi[g];
    const uint64_t add_lo = lo + (hi << 32);
    const uint32_t c0 = add_lo < lo ? 1u : 0u;
    s_lo += add_lo;
    s_hi += (uint32_t)(hi >> 32) + c0 + (s_lo < add_lo ? 1u : 0u);
  }
  /* Y = X - qhat P in [0, 2P): qhat underestimates floor(X/P) by at most 1 (sum of NG truncations < NG 2^32 << 2^64) */
  {
    const uint32_t qh = s_hi;
    uint32_t carry = 0; uint32_t borrow = 0;
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) {
      const uint64_t m = (uint64_t)qh * pl.P[l] + carry;
      carry = (uint32_t)(m >> 32);
      const uint64_t d = (uint64_t)Y[l] - (uint32_t)m - borrow;
      Y[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63);
    }
  }
  if (limbs_cmp<NL>(Y, pl.P, true)) {
    uint32_t borrow = 0;
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) { const uint64_t d = (uint64_t)Y[l] - pl.P[l] - borrow; Y[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
  }
  neg = limbs_cmp<NL>(Y, pl.hP, false) ? 1u : 0u;
  if (neg) { /* |I| = P - Y */
    uint32_t borrow = 0;
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) { const uint64_t d = (uint64_t)pl.P[l] - Y[l] - borrow; Y[l] = (uint32_t)d; borrow = (uint32_t)(d >> 63); }
  }
}


/* ------------------------------------------------------------------ the one rounding */
QCRT_HD int clz32(uint32_t x) /* x != 0 */
{
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return __builtin_clz(x);
#endif
}
QCRT_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, int n) /* high word of (hi:lo) << n, 0 <= n < 32 */
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, hi, n);
#else
  return n ? (hi << n) | (lo >> (32 - n)) : hi;
#endif
}
/* |I| (NL limbs) times 2^Eb -> binary128 with ONE rounding (RNE; gradual underflow and overflow in q_round_pack).
 * I = 0 gives +0: an exact zero sum is +0 (a +0-seeded chain never yields -0). */
template <int NL>
QCRT_HD q128 limbs_to_q(const uint32_t (&L)[NL], uint32_t neg, int Eb)
{
  int top = -1;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) if (L[l]) top = l;
  if (top < 0) return q_zero(0);
  uint32_t buf[NL + 8];
QCRT_UNROLL
  for (int l = 0; l < 8; ++l) buf[l] = 0;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) buf[8 + l] = L[l];
  uint32_t w[9];
QCRT_UNROLL
  for (int k = 0; k < 9; ++k) w[k] = buf[top + k];   /* limbs top-8 .. top */
  uint32_t sticky = 0;
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) if (l < top - 8) sticky |= L[l];
  const int lz = clz32(w[8]);
  uint32_t R[8];
QCRT_UNROLL
  for (int k = 0; k < 8; ++k) R[k] = funnel_l(w[k], w[k + 1], lz);
  sticky |= w[0] << lz;
  if (lz == 0) sticky |= w[0];
  u256 Rq;
  Rq.w0 = ((uint64_t)R[1] << 32) | R[0] | (sticky != 0);
  Rq.w1 = ((uint64_t)R[3] << 32) | R[2];
  Rq.w2 = ((uint64_t)R[5] << 32) | R[4];
  Rq.w3 = ((uint64_t)R[7] << 32) | R[6];
  const int p = 32 * top + 31 - lz;  /* MSB position of |I| */
  return q_round_pack(neg, p + Eb + QBIAS, Rq);
}

/* ------------------------------------------------------------------ reconstruction, device form */
#if defined(__CUDACC__) || defined(QCRT_HOST_EMULATE_PTX)
/* Same result as reconstruct<NG>() (the reference form above), written on PTX carry chains: the products t * M_g
 * go into an even-column and an odd-column accumulator as mad.lo.cc / madc.hi.cc pairs (ptxas fuses each pair into one
 * IMAD.WIDE with carry), the group digit t_g uses a 32-bit reciprocal of the normalised group modulus.
 * tests/host/crt_host.cpp defines QCRT_HOST_EMULATE_PTX: the same source then runs on the CPU with the carry flag emulated
 * (one flag per thread), so every group count NG = 1..13 of this form is checked against Python integers without a GPU. */
namespace ptx {
#if defined(__CUDACC__)
#define QCRT_DI __device__ __forceinline__
QCRT_DI uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
QCRT_DI uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
QCRT_DI uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
QCRT_DI uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
QCRT_DI uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
QCRT_DI uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
QCRT_DI uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
QCRT_DI uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
QCRT_DI uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
QCRT_DI uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#else
#define QCRT_DI static inline
static thread_local uint32_t cc_cf = 0;   /* CC.CF: carry out of add / mad, borrow out of sub */
QCRT_DI uint32_t lo32(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
QCRT_DI uint32_t hi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
QCRT_DI uint32_t add3(uint32_t a, uint32_t b, uint32_t cin, bool set) { const uint64_t t = (uint64_t)a + b + cin; if (set) cc_cf = (uint32_t)(t >> 32); return (uint32_t)t; }
QCRT_DI uint32_t sub3(uint32_t a, uint32_t b, uint32_t bin, bool set) { const uint64_t t = (uint64_t)a - b - bin; if (set) cc_cf = (uint32_t)(t >> 63); return (uint32_t)t; }
QCRT_DI uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(lo32(a, b), c, 0, true); }
QCRT_DI uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(lo32(a, b), c, cc_cf, true); }
QCRT_DI uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add3(hi32(a, b), c, cc_cf, true); }
QCRT_DI uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return add3(hi32(a, b), c, cc_cf, false); }
QCRT_DI uint32_t add_cc(uint32_t a, uint32_t b) { return add3(a, b, 0, true); }
QCRT_DI uint32_t addc_cc(uint32_t a, uint32_t b) { return add3(a, b, cc_cf, true); }
QCRT_DI uint32_t addc(uint32_t a, uint32_t b) { return add3(a, b, cc_cf, false); }
QCRT_DI uint32_t sub_cc(uint32_t a, uint32_t b) { return sub3(a, b, 0, true); }
QCRT_DI uint32_t subc_cc(uint32_t a, uint32_t b) { return sub3(a, b, cc_cf, true); }
QCRT_DI uint32_t subc(uint32_t a, uint32_t b) { return sub3(a, b, cc_cf, false); }
#endif

/* acc (NL limbs, even or odd columns only) += t * m[l] for l = L0, L0 + 2, ... < NM; carries stay inside NL limbs */
template <int NL, int NMUL, int L0>
QCRT_DI void mad_columns(uint32_t (&acc)[NL], uint32_t t, const uint32_t *m)
{
  if (L0 >= NMUL) return;
  acc[L0] = mad_lo_cc(t, m[L0], acc[L0]);
  acc[L0 + 1] = madc_hi_cc(t, m[L0], acc[L0 + 1]);
  int last = L0 + 1;
QCRT_UNROLL
  for (int l = L0 + 2; l < NMUL; l += 2) {
    acc[l] = madc_lo_cc(t, m[l], acc[l]);
    acc[l + 1] = madc_hi_cc(t, m[l], acc[l + 1]);
    last = l + 1;
  }
QCRT_UNROLL
  for (int l = 0; l < NL; ++l)
    if (l > last) acc[l] = (l == NL - 1) ? addc(acc[l], 0u) : addc_cc(acc[l], 0u);
}
} // namespace ptx

template <int NG>
QCRT_DI void reconstruct_dev(const uint32_t (&r)[NMP], const Plan &pl, uint32_t (&Y)[NG + 1], uint32_t &neg)
{
  constexpr int NL = NG + 1;
  uint32_t E[NL], O[NL];
QCRT_UNROLL
  for (int l = 0; l < NL; ++l) { E[l] = 0; O[l] = 0; }
  uint32_t s0 = 0, s1 = 0, s2 = 0;          /* S = sum t_g f_g (96 bits) */
QCRT_UNROLL
  for (int g = 0; g < NG; ++g) {
    uint64_t v = 0;
QCRT_UNROLL
    for (int b = 0; b < 4; ++b) v += (uint64_t)r[4 * g + b] * pl.cp[4 * g + b];   /* cp = 0 beyond N */
    /* t = v mod G: quotient (< 2^10) from the top 31 bits of v 2^sh and a 32-bit reciprocal of G 2^sh in [2^31, 2^32) */
    const uint32_t a = (uint32_t)(v >> pl.gshr[g]) << pl.gshl[g];  /* (v << sh) >> 11 < 2^31 */
    const uint32_t qh = mulhi_u(a, pl.ginv[g]) >> 20;               /* floor(v / G) or one less */
    uint64_t t64 = v - (uint64_t)qh * pl.G[g];
    if (t64 >= pl.G[g]) t64 -= pl.G[g];
    const uint32_t t = (uint32_t)t64;
    ptx::mad_columns<NL, NG, 0>(E, t, pl.M[g]);
    ptx::mad_columns<NL, NG, 1>(O, t, pl.M[g]);
    s0 = ptx::mad_lo_cc(t, pl.flo[g], s0); s1 = ptx::madc_hi_cc(t, pl.flo[g], s1); s2 = ptx::addc(s2, 0u);
    s1 = ptx::mad_lo_cc(t, pl.fhi[g], s1); s2 = ptx::madc_hi(t, pl.fhi[g], s2);
  }
  /* Y = E + O - qhat P  in [0, 2P) */
  {
    uint32_t QE[NL], QO[NL];
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) { QE[l] = 0; QO[l] = 0; }
    ptx::mad_columns<NL, NG, 0>(QE, s2, pl.P);
    ptx::mad_columns<NL, NG, 1>(QO, s2, pl.P);
    Y[0] = ptx::add_cc(E[0], O[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) Y[l] = (l == NL - 1) ? ptx::addc(E[l], O[l]) : ptx::addc_cc(E[l], O[l]);
    Y[0] = ptx::sub_cc(Y[0], QE[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) Y[l] = (l == NL - 1) ? ptx::subc(Y[l], QE[l]) : ptx::subc_cc(Y[l], QE[l]);
    Y[0] = ptx::sub_cc(Y[0], QO[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) Y[l] = (l == NL - 1) ? ptx::subc(Y[l], QO[l]) : ptx::subc_cc(Y[l], QO[l]);
  }
  /* Y >= P: subtract P once */
  {
    uint32_t T[NL];
    T[0] = ptx::sub_cc(Y[0], pl.P[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) T[l] = ptx::subc_cc(Y[l], pl.P[l]);
    const uint32_t below = ptx::subc(0u, 0u);   /* all ones when Y < P */
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) Y[l] = below ? Y[l] : T[l];
  }
  /* |I| = min(Y, P - Y); negative when P - Y is the smaller one (Y = P/2 cannot happen: |I| < P/2) */
  {
    uint32_t U[NL];
    U[0] = ptx::sub_cc(pl.P[0], Y[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) U[l] = (l == NL - 1) ? ptx::subc(pl.P[l], Y[l]) : ptx::subc_cc(pl.P[l], Y[l]);
    (void)ptx::sub_cc(U[0], Y[0]);
QCRT_UNROLL
    for (int l = 1; l < NL; ++l) (void)ptx::subc_cc(U[l], Y[l]);
    const uint32_t u_below = ptx::subc(0u, 0u); /* all ones when U < Y */
    neg = u_below ? 1u : 0u;
QCRT_UNROLL
    for (int l = 0; l < NL; ++l) Y[l] = u_below ? U[l] : Y[l];
  }
}
#endif

/* ------------------------------------------------------------------ host: tables and plans */
namespace host {

static inline int64_t inv_mod(int64_t a, int64_t m) /* a^-1 mod m, gcd = 1 */
{
  int64_t g = m, x = 0, y = 1, aa = ((a % m) + m) % m;
  while (aa) { const int64_t q = g / aa; int64_t t = g - q * aa; g = aa; aa = t; t = x - q * y; x = y; y = t; }
  return ((x % m) + m) % m;
}

static inline void build_tables(Tables &T)
{
  memset(&T, 0, sizeof(T));
  for (int i = 0; i < NMP; ++i) {
    const uint32_t p = i < NM ? (uint32_t)MODULI[i] : 1u;
    T.p[i] = p;
    T.half[i] = p == 256 ? 128u : (p - 1) / 2;
    T.cneg[i] = p * (((1u << 21) + p - 1) / p) + 2 * T.half[i];
    T.minv[i] = (uint32_t)((((uint64_t)1 << 32) + p - 1) / p);
    T.off[i] = p * (((1u << 30) + p - 1) / p);
    T.finv[i] = p == 1 ? 0xffffffffu : (uint32_t)(((uint64_t)1 << 32) / p);
    uint32_t pw = 1 % p;
    for (int j = 0; j < NWMAX; ++j) {
      uint32_t word = 0;
      for (int b = 0; b < 4; ++b) { word |= pw << (8 * b); pw = (pw * 256u) % p; }
      T.pw[j][i] = i < NM ? word : 0u;
      /* after the loop body above pw == 256^(4(j+1)) mod p == 2^(32 (j+1)) mod p */
      T.seedn[j][i] = T.half[i] + (p - pw % p) % p;
    }
  }
}

/* little-endian multi-limb helpers on fixed NLMAX + 1 limbs */
struct Big { uint32_t w[NLMAX + 2]; };
static inline Big big_one() { Big b; memset(&b, 0, sizeof(b)); b.w[0] = 1; return b; }
static inline void big_mul_small(Big &a, uint32_t m)
{
  uint64_t c = 0;
  for (int l = 0; l < NLMAX + 2; ++l) { const uint64_t t = (uint64_t)a.w[l] * m + c; a.w[l] = (uint32_t)t; c = t >> 32; }
}
static inline uint32_t big_div_small(Big &a, uint32_t d) /* a /= d, returns the remainder */
{
  uint64_t r = 0;
  for (int l = NLMAX + 1; l >= 0; --l) { const uint64_t t = (r << 32) | a.w[l]; a.w[l] = (uint32_t)(t / d); r = t % d; }
  return (uint32_t)r;
}
static inline uint32_t big_mod_small(const Big &a, uint32_t d) { Big t = a; return big_div_small(t, d); }
static inline double big_log2(const Big &a)
{
  int top = NLMAX + 1;
  while (top > 0 && !a.w[top]) --top;
  double v = 0;
  for (int l = top; l >= 0 && l > top - 3; --l) v = v * 4294967296.0 + a.w[l];
  int sh = top >= 2 ? top - 2 : 0;
  return __builtin_log2(v) + 32.0 * sh;
}

/* plan for the first N moduli (1 <= N <= NM) */
static inline void build_plan(int N, Plan &pl)
{
  memset(&pl, 0, sizeof(pl));
  pl.N = N; pl.NG = (N + 3) / 4;
  Big P = big_one();
  for (int i = 0; i < N; ++i) big_mul_small(P, (uint32_t)MODULI[i]);
  for (int l = 0; l < NLMAX; ++l) pl.P[l] = P.w[l];
  { Big h = P; big_div_small(h, 2); for (int l = 0; l < NLMAX; ++l) pl.hP[l] = h.w[l]; }
  pl.bits = big_log2(P);
  for (int g = 0; g < pl.NG; ++g) {
    uint64_t G = 1;
    for (int b = 0; b < 4 && 4 * g + b < N; ++b) G *= (uint64_t)MODULI[4 * g + b];
    pl.G[g] = (uint32_t)G;
    const unsigned __int128 two64 = (unsigned __int128)1 << 64;
    const uint64_t f = G == 1 ? ~(uint64_t)0 : (uint64_t)(two64 / G);
    pl.flo[g] = (uint32_t)f; pl.fhi[g] = (uint32_t)(f >> 32);
    {
      int lg = 0;
      while ((G >> (lg + 1)) != 0) ++lg;
      const int sh = 31 - lg;                    /* G << sh in [2^31, 2^32) */
      pl.gshr[g] = sh < 11 ? (uint32_t)(11 - sh) : 0u;
      pl.gshl[g] = sh > 11 ? (uint32_t)(sh - 11) : 0u;
      pl.ginv[g] = (uint32_t)((((uint64_t)1 << 63) - 1) / (G << sh));
    }
    Big M = P;
    big_div_small(M, (uint32_t)G);
    for (int l = 0; l < NLMAX; ++l) pl.M[g][l] = M.w[l];
    const uint64_t y = G == 1 ? 0 : (uint64_t)inv_mod((int64_t)big_mod_small(M, (uint32_t)G), (int64_t)G);
    for (int b = 0; b < 4 && 4 * g + b < N; ++b) {
      const uint64_t p = (uint64_t)MODULI[4 * g + b], Gp = G / p;
      const uint64_t e = (uint64_t)(((unsigned __int128)Gp * (uint64_t)inv_mod((int64_t)(Gp % p), (int64_t)p)) % G); /* == 1 mod p, 0 mod the others */
      pl.cp[4 * g + b] = (uint32_t)(((unsigned __int128)e * y) % G);
    }
  }
}

/* smallest N with prod_{i<N} p_i > 2^need_bits, 0 if even all NM moduli are not enough */
static inline int moduli_for_bits(int need_bits)
{
  double bits = 0;
  for (int i = 0; i < NM; ++i) {
    bits += __builtin_log2((double)MODULI[i]);
    if (bits > need_bits + 1e-6) return i + 1;
  }
  return 0;
}

/* Windows and moduli count of one qgemm.  WA_nat / WB_nat = widest bit span of a row of op(A) / column of op(B) (scan).  The
 * windows are the spans themselves when the moduli can cover them inside the `wcap` budget (WA + WB <= 2 wcap): then every
 * element is represented exactly and so is every inner product.  Otherwise the wider window is cut (both to wcap when both
 * exceed it) and the low bits of small elements are dropped (element_words); accept_msb() is the per-element test that keeps
 * the result inside the fast-mode contract.  false: even all the moduli cannot cover 2 x 117 bits at this k. */
struct Windows { int WA, WB, N, truncA, truncB; };
static inline bool plan_windows(int WA_nat, int WB_nat, long long k, int wcap, Windows &w)
{
  int lk = 0;
  while (((long long)1 << lk) < k) ++lk;
  double all = 0;
  for (int i = 0; i < NM; ++i) all += __builtin_log2((double)MODULI[i]);
  int maxsum = (int)(all - 1e-6) - lk - 1;              /* P > 2 k 2^(WA + WB) */
  if (2 * wcap < maxsum) maxsum = 2 * wcap;
  int WA = WA_nat < 1 ? 1 : (WA_nat > WMAX ? WMAX : WA_nat), WB = WB_nat < 1 ? 1 : (WB_nat > WMAX ? WMAX : WB_nat);
  if (WA + WB > maxsum) {
    const int half = maxsum / 2;
    if (WA <= half) WB = maxsum - WA;
    else if (WB <= half) WA = maxsum - WB;
    else { WA = half; WB = maxsum - half; }
  }
  if ((WA < WA_nat && WA < 117) || (WB < WB_nat && WB < 117)) return false;   /* a cut window must still hold the largest element */
  w.WA = WA; w.WB = WB; w.truncA = WA < WA_nat; w.truncB = WB < WB_nat;
  w.N = moduli_for_bits(WA + WB + lk + 1);
  return w.N != 0;
}

} // namespace host

} // namespace crt
} // namespace qb

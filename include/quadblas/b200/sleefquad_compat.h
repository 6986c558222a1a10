/*
 * quadblas/b200/sleefquad_compat.h — the slice of <sleefquad.h> that QuadBLAS callers touch, for
 * hosts where SLEEF's quad library is not installed.
 *
 * The reference pulls <sleefquad.h> in from core/platform.hpp
 * (/root/reference/include/quadblas/core/platform.hpp:17-22) and its tests, benchmarks and README
 * examples use a handful of SLEEF scalar symbols directly on the HOST:
 * Sleef_quad, SLEEF_QUAD_C, Sleef_cast_from_doubleq1, Sleef_cast_to_doubleq1, Sleef_{add,mul,fma,
 * sqrt}q1_u05 and the 2-lane Sleef_quadx2 helpers.  This header provides exactly those for CALLER
 * code (filling inputs, printing results, the callers' own naive check loops).  No BLAS routine of
 * this library computes with them: QuadBLAS::gemm/gemv/dot/axpy and Vector/Matrix forward to the
 * CUDA library (libqblas_b200.so) and fail when it is unusable.
 *
 * The scalar ops are the library's own integer-limb binary128 arithmetic (csrc/q128.cuh compiled
 * for the host), so they are correctly rounded (= the _u05 contract) without libquadmath.
 * If real SLEEF is on the include path, core/platform.hpp uses it instead of this file.
 */
#ifndef QUADBLAS_B200_SLEEFQUAD_COMPAT_H
#define QUADBLAS_B200_SLEEFQUAD_COMPAT_H

#include <string.h>
#include "q128.cuh"

#if !defined(__SIZEOF_FLOAT128__)
#error "quadblas-b200: this compiler has no __float128; install SLEEF (sleefquad.h) to provide Sleef_quad"
#endif

typedef __float128 Sleef_quad;
typedef struct { Sleef_quad v[2]; } Sleef_quadx2;
#ifndef SLEEF_QUAD_C
#define SLEEF_QUAD_C(x) (x##Q) /* needs -std=gnu++NN (or -fext-numeric-literals), like SLEEF's own macro */
#endif

static inline q128 qb_compat_bits(Sleef_quad a) { q128 r; memcpy(&r, &a, 16); return r; }
static inline Sleef_quad qb_compat_quad(q128 a) { Sleef_quad r; memcpy(&r, &a, 16); return r; }

static inline Sleef_quad Sleef_fmaq1_u05(Sleef_quad a, Sleef_quad b, Sleef_quad c) { return qb_compat_quad(qb::q_fma(qb_compat_bits(a), qb_compat_bits(b), qb_compat_bits(c))); }
static inline Sleef_quad Sleef_mulq1_u05(Sleef_quad a, Sleef_quad b) { return qb_compat_quad(qb::q_mul(qb_compat_bits(a), qb_compat_bits(b))); }
static inline Sleef_quad Sleef_addq1_u05(Sleef_quad a, Sleef_quad b) { return qb_compat_quad(qb::q_add(qb_compat_bits(a), qb_compat_bits(b))); }
static inline Sleef_quad Sleef_subq1_u05(Sleef_quad a, Sleef_quad b) { return qb_compat_quad(qb::q_sub(qb_compat_bits(a), qb_compat_bits(b))); }
static inline Sleef_quad Sleef_sqrtq1_u05(Sleef_quad a) { return qb_compat_quad(qb::q_sqrt(qb_compat_bits(a))); }
static inline Sleef_quad Sleef_negq1(Sleef_quad a) { return qb_compat_quad(qb::q_neg(qb_compat_bits(a))); }
static inline Sleef_quad Sleef_fabsq1(Sleef_quad a) { return qb_compat_quad(qb::q_abs(qb_compat_bits(a))); }
static inline Sleef_quad Sleef_cast_from_doubleq1(double d) { uint64_t b; memcpy(&b, &d, 8); return qb_compat_quad(qb::q_from_double_bits(b)); }
static inline double Sleef_cast_to_doubleq1(Sleef_quad a) { uint64_t b = qb::q_to_double_bits(qb_compat_bits(a)); double d; memcpy(&d, &b, 8); return d; }

/* two-lane helpers (x86-64 "sse2" names; the AArch64 "advsimd" twins alias them) */
static inline Sleef_quadx2 Sleef_splatq2_sse2(Sleef_quad a) { Sleef_quadx2 r; r.v[0] = a; r.v[1] = a; return r; }
static inline Sleef_quadx2 Sleef_loadq2_sse2(Sleef_quad *p) { Sleef_quadx2 r; r.v[0] = p[0]; r.v[1] = p[1]; return r; }
static inline void Sleef_storeq2_sse2(Sleef_quad *p, Sleef_quadx2 a) { p[0] = a.v[0]; p[1] = a.v[1]; }
static inline Sleef_quad Sleef_getq2_sse2(Sleef_quadx2 a, int lane) { return a.v[lane]; }
static inline Sleef_quadx2 Sleef_setq2_sse2(Sleef_quadx2 a, int lane, Sleef_quad v) { a.v[lane] = v; return a; }
static inline Sleef_quadx2 Sleef_addq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b) { Sleef_quadx2 r; for (int l = 0; l < 2; ++l) r.v[l] = Sleef_addq1_u05(a.v[l], b.v[l]); return r; }
static inline Sleef_quadx2 Sleef_mulq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b) { Sleef_quadx2 r; for (int l = 0; l < 2; ++l) r.v[l] = Sleef_mulq1_u05(a.v[l], b.v[l]); return r; }
static inline Sleef_quadx2 Sleef_fmaq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b, Sleef_quadx2 c) { Sleef_quadx2 r; for (int l = 0; l < 2; ++l) r.v[l] = Sleef_fmaq1_u05(a.v[l], b.v[l], c.v[l]); return r; }
#define Sleef_splatq2_advsimd Sleef_splatq2_sse2
#define Sleef_loadq2_advsimd Sleef_loadq2_sse2
#define Sleef_storeq2_advsimd Sleef_storeq2_sse2
#define Sleef_getq2_advsimd Sleef_getq2_sse2
#define Sleef_setq2_advsimd Sleef_setq2_sse2
#define Sleef_addq2_u05advsimd Sleef_addq2_u05sse2
#define Sleef_mulq2_u05advsimd Sleef_mulq2_u05sse2
#define Sleef_fmaq2_u05advsimd Sleef_fmaq2_u05sse2

#endif /* QUADBLAS_B200_SLEEFQUAD_COMPAT_H */

/*
 * oracle/shim/sleefquad.h — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Stand-in for SLEEF's <sleefquad.h> (shibatch/sleef tag 3.8, libsleefquad), which is an
 * un-vendored third-party dependency of the reference (pinned only in
 * /root/reference/.github/workflows/ci.yml:26) and cannot be fetched offline.
 *
 * It declares exactly the 15 SLEEF names the reference headers touch on x86-64
 *  and maps them onto GCC's __float128 / libquadmath.  The *_u05 SLEEF
 * functions promise <= 0.5 ULP, i.e. the correctly rounded round-to-nearest-even result, which
 * is unique; libgcc's __addtf3/__multf3 and libquadmath's fmaq/sqrtq are correctly rounded too,
 * so the bits agree (NaN payloads excepted).  tests/test_oracle_scalar.py re-checks that claim
 * against exact rational arithmetic.
 */
#ifndef QB_ORACLE_SLEEFQUAD_SHIM_H
#define QB_ORACLE_SLEEFQUAD_SHIM_H

#include <quadmath.h>

typedef __float128 Sleef_quad;
typedef struct { Sleef_quad v[2]; } Sleef_quadx2;

#define SLEEF_QUAD_C(x) (x##Q)

static inline Sleef_quad Sleef_addq1_u05(Sleef_quad a, Sleef_quad b) { return a + b; }
static inline Sleef_quad Sleef_mulq1_u05(Sleef_quad a, Sleef_quad b) { return a * b; }
static inline Sleef_quad Sleef_fmaq1_u05(Sleef_quad a, Sleef_quad b, Sleef_quad c) { return fmaq(a, b, c); }
static inline Sleef_quad Sleef_sqrtq1_u05(Sleef_quad a) { return sqrtq(a); }
static inline Sleef_quad Sleef_cast_from_doubleq1(double d) { return (Sleef_quad)d; }
static inline double Sleef_cast_to_doubleq1(Sleef_quad q) { return (double)q; }

static inline Sleef_quadx2 Sleef_splatq2_sse2(Sleef_quad a) { Sleef_quadx2 r; r.v[0] = a; r.v[1] = a; return r; }
static inline Sleef_quadx2 Sleef_loadq2_sse2(Sleef_quad *p) { Sleef_quadx2 r; r.v[0] = p[0]; r.v[1] = p[1]; return r; }
static inline Sleef_quad Sleef_getq2_sse2(Sleef_quadx2 a, int i) { return a.v[i]; }
static inline Sleef_quadx2 Sleef_addq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b)
{ Sleef_quadx2 r; r.v[0] = a.v[0] + b.v[0]; r.v[1] = a.v[1] + b.v[1]; return r; }
static inline Sleef_quadx2 Sleef_mulq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b)
{ Sleef_quadx2 r; r.v[0] = a.v[0] * b.v[0]; r.v[1] = a.v[1] * b.v[1]; return r; }
static inline Sleef_quadx2 Sleef_fmaq2_u05sse2(Sleef_quadx2 a, Sleef_quadx2 b, Sleef_quadx2 c)
{ Sleef_quadx2 r; r.v[0] = fmaq(a.v[0], b.v[0], c.v[0]); r.v[1] = fmaq(a.v[1], b.v[1], c.v[1]); return r; }

#endif

/*
 * oracle/qoracle.c — TEST INFRASTRUCTURE ONLY: the CPU oracle for the binary128 hot path.
 *
 * A plain-C restatement of the reference's per-element algorithms ("reference order",
 * SURVEY.md §8a / Appendix B).  It is NOT the product and is never linked or imported by the
 * product library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it,
 * and only as the checker.
 *
 * Arithmetic: GCC __float128 (+, * from libgcc soft-fp; fmaq, sqrtq from libquadmath).  The
 * reference's arithmetic lives in SLEEF (shibatch/sleef tag 3.8, libsleefquad — an un-vendored
 * dependency pinned only by /root/reference/.github/workflows/ci.yml:26, absent offline).  SLEEF's
 * *_u05 functions are specified as <= 0.5 ULP, i.e. correctly rounded RNE, which is a unique
 * result; the GCC functions used here are correctly rounded as well (re-checked against exact
 * rational arithmetic in tests/test_oracle_scalar.py).
 *
 * Parity pinning: bitwise against the reference's own loops compiled from /root/reference
 * (oracle/_ref/libqref.so, tests/test_oracle_vs_reference.py), against golden vectors generated
 * from that build (tests/golden/, tests/golden/make_golden.py) and against every known-answer
 * value in the reference's tests (tests/test_reference_known_answers.py).  At the SLEEF boundary
 * itself the reference's tests never compare bits, so scalar-op parity rests on
 * IEEE-754 correct rounding.
 *
 * Loops are OpenMP-parallel over OUTPUT elements only; the reduction order inside an element is
 * exactly the reference's, so thread count never changes a bit.
 */
#include <math.h>
#include <quadmath.h>
#include <stddef.h>
#include <stdint.h>

typedef __float128 Q;

/* ------------------------------------------------------------------ scalar ops (a15) */
/* Sleef_fmaq1_u05 / Sleef_mulq1_u05 / Sleef_addq1_u05 / Sleef_sqrtq1_u05 call sites:
 * level1.hpp:31,70,113,119; level2.hpp:43,48,65,71,79; level3.hpp:36,40,97,107,206,210;
 * c_interface.hpp:30,42-43. */
void orc_fma(const void *a, const void *b, const void *c, void *o) { *(Q *)o = fmaq(*(const Q *)a, *(const Q *)b, *(const Q *)c); }
void orc_mul(const void *a, const void *b, void *o) { *(Q *)o = *(const Q *)a * *(const Q *)b; }
void orc_add(const void *a, const void *b, void *o) { *(Q *)o = *(const Q *)a + *(const Q *)b; }
void orc_sqrt(const void *a, void *o) { *(Q *)o = sqrtq(*(const Q *)a); }
void orc_from_double(double d, void *o) { *(Q *)o = (Q)d; }
double orc_to_double(const void *a) { return (double)*(const Q *)a; }

void orc_fma_n(long n, const void *a, const void *b, const void *c, void *o)
{
  const Q *qa = a, *qb = b, *qc = c; Q *qo = o;
#pragma omp parallel for
  for (long i = 0; i < n; ++i) qo[i] = fmaq(qa[i], qb[i], qc[i]);
}
void orc_add_n(long n, const void *a, const void *b, void *o)
{ const Q *qa = a, *qb = b; Q *qo = o; for (long i = 0; i < n; ++i) qo[i] = qa[i] + qb[i]; }
void orc_mul_n(long n, const void *a, const void *b, void *o)
{ const Q *qa = a, *qb = b; Q *qo = o; for (long i = 0; i < n; ++i) qo[i] = qa[i] * qb[i]; }
void orc_sqrt_n(long n, const void *a, void *o)
{ const Q *qa = a; Q *qo = o; for (long i = 0; i < n; ++i) qo[i] = sqrtq(qa[i]); }

/* ------------------------------------------------------------------ DOTK (a12) */
/* level1.hpp:14-35 dot_kernel_vectorized: two interleaved chains over even/odd indices from +0
 * (the two Sleef_quadx2 lanes), lane0+lane1 (quad_vector.hpp:141-150), then the odd tail folded
 * with one scalar fma (level1.hpp:29-32). */
static Q dotk(const Q *x, const Q *y, size_t n)
{
  Q l0 = 0.0Q, l1 = 0.0Q;
  size_t h = n / 2;
  for (size_t p = 0; p < h; ++p) {
    l0 = fmaq(x[2 * p], y[2 * p], l0);
    l1 = fmaq(x[2 * p + 1], y[2 * p + 1], l1);
  }
  Q r = l0 + l1;
  if (n & 1) r = fmaq(x[n - 1], y[n - 1], r);
  return r;
}

/* ------------------------------------------------------------------ DOT (a10, a11) */
/* level1.hpp:80-137 dot + :38-77 dot_parallel.  T = quadblas_get_num_threads()
 * (threading/openmp_utils.hpp:10-17) is part of the numerical contract: chunk = n/T, last chunk
 * takes the remainder, partials folded with add from +0 in tid order.  Threshold 500 =
 * PARALLEL_THRESHOLD (core/constants.hpp:15). */
void orc_dot(long n_, const void *x_, long incx, const void *y_, long incy, int T, void *out)
{
  const Q *x = x_, *y = y_;
  size_t n = (size_t)n_;
  Q r = 0.0Q;
  if (n == 0) { *(Q *)out = r; return; }
  if (T < 1) T = 1;
  if (incx == 1 && incy == 1) {
    if (n < 500) { *(Q *)out = dotk(x, y, n); return; }
    size_t chunk = n / (size_t)T;
    Q part[T];
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      size_t s = (size_t)t * chunk, e = (t == T - 1) ? n : s + chunk;
      part[t] = (s < e) ? dotk(x + s, y + s, e - s) : 0.0Q;
    }
    for (int t = 0; t < T; ++t) r = r + part[t];
    *(Q *)out = r;
    return;
  }
  if (n >= 500) { /* level1.hpp:93-120: each chunk is ONE sequential chain */
    size_t chunk = n / (size_t)T;
    Q part[T];
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < T; ++t) {
      size_t s = (size_t)t * chunk, e = (t == T - 1) ? n : s + chunk;
      Q p = 0.0Q;
      for (size_t i = s; i < e; ++i) p = fmaq(x[i * (size_t)incx], y[i * (size_t)incy], p);
      part[t] = p;
    }
    for (int t = 0; t < T; ++t) r = r + part[t];
  } else { /* level1.hpp:128-134 */
    for (size_t i = 0; i < n; ++i) r = fmaq(x[i * (size_t)incx], y[i * (size_t)incy], r);
  }
  *(Q *)out = r;
}

/* c_interface.hpp:34-44 / cpp_classes.hpp:78-81: sqrt(dot(x,x)), no scaling. */
void orc_nrm2(long n, const void *x, long incx, int T, void *out)
{
  Q d;
  orc_dot(n, x, incx, x, incx, T, &d);
  *(Q *)out = sqrtq(d);
}

/* level1.hpp:140-223: y_i = fma(alpha, x_i, y_i); order-free. n==0 -> no-op. */
void orc_axpy(long n, const void *alpha_, const void *x_, long incx, void *y_, long incy)
{
  const Q alpha = *(const Q *)alpha_; const Q *x = x_; Q *y = y_;
#pragma omp parallel for
  for (long i = 0; i < n; ++i) y[i * incy] = fmaq(alpha, x[i * incx], y[i * incy]);
}

/* ------------------------------------------------------------------ GEMV (a8, a9) */
/* layout: 'C'/'c' = ColMajor else RowMajor (c_interface.hpp:76).  This is QuadBLAS::gemv
 * (level2.hpp:85-99), i.e. AFTER the C ABI's transpose relabelling. */
void orc_gemv(char layout, long m_, long n_, const void *alpha_, const void *A_, long lda_,
              const void *x_, long incx_, const void *beta_, void *y_, long incy_)
{
  const Q alpha = *(const Q *)alpha_, beta = *(const Q *)beta_;
  const Q *A = A_, *x = x_; Q *y = y_;
  size_t m = (size_t)m_, n = (size_t)n_, lda = (size_t)lda_, incx = (size_t)incx_, incy = (size_t)incy_;
  if (m == 0 || n == 0) return; /* level2.hpp:21,59: y is NOT scaled */
  if (!(layout == 'C' || layout == 'c')) {
    /* gemv_row_major level2.hpp:15-50 */
#pragma omp parallel for
    for (size_t i = 0; i < m; ++i) {
      const Q *row = A + i * lda;
      Q s = 0.0Q;
      if (incx == 1) s = dotk(row, x, n);
      else for (size_t j = 0; j < n; ++j) s = fmaq(row[j], x[j * incx], s);
      y[i * incy] = fmaq(alpha, s, beta * y[i * incy]);
    }
  } else {
    /* gemv_col_major level2.hpp:53-82: y=beta*y; per column j: c=alpha*x_j; y_i=fma(A[j*lda+i],c,y_i) */
#pragma omp parallel for
    for (size_t i = 0; i < m; ++i) {
      Q acc = beta * y[i * incy];
      for (size_t j = 0; j < n; ++j) acc = fmaq(A[j * lda + i], alpha * x[j * incx], acc);
      y[i * incy] = acc;
    }
  }
}

/* quadblas_qgemv marshalling, c_interface.hpp:64-90: trans in {T,t,C,c} => swap(m,n), flip layout. */
void orc_c_qgemv(char layout, char trans, int m, int n, double alpha, const void *A, int lda,
                 const void *x, int incx, double beta, void *y, int incy)
{
  Q qa = (Q)alpha, qb = (Q)beta;
  int col = (layout == 'C' || layout == 'c');
  if (trans == 'T' || trans == 't' || trans == 'C' || trans == 'c') { int t = m; m = n; n = t; col = !col; }
  orc_gemv(col ? 'C' : 'R', m, n, &qa, A, lda, x, incx, &qb, y, incy);
}

/* ------------------------------------------------------------------ GEMM (a2-a6) */
/* level3.hpp:215-336 (+ gemm_simple :187-213, macro/micro kernels :20-185), per element:
 *   k-panels of KC (detail/blocking.hpp:21-66 -> 126 on x86-64 Linux, min(126,k));
 *   s_q = fma chain from +0 over the panel, ascending l (level3.hpp:77-85);
 *   C = fma(alpha, s_q, mul(q==0 ? beta : 1, C))  (level3.hpp:102-109 with :299).
 * When m,n,k <= 64 the reference takes gemm_simple = one chain over all k, which is the same
 * formula with a single panel (k <= 64 < 126).  The blocked path only exists when some dim > 64;
 * then kc = min(126,k), so the formula below with kc_param = 126 covers both.
 * ColMajor: the reference's blocked path is broken (SURVEY exec-summary bug 1); parity for
 * ColMajor is DEFINED by this layout-agnostic formula (and equals gemm_simple when dims <= 64). */
void orc_gemm(char layout, long m_, long n_, long k_, const void *alpha_, const void *A_, long lda_,
              const void *B_, long ldb_, const void *beta_, void *C_, long ldc_, long kc_)
{
  const Q alpha = *(const Q *)alpha_, beta = *(const Q *)beta_;
  const Q *A = A_, *B = B_; Q *C = C_;
  size_t m = (size_t)m_, n = (size_t)n_, k = (size_t)k_, lda = (size_t)lda_, ldb = (size_t)ldb_, ldc = (size_t)ldc_;
  size_t kc = kc_ > 0 ? (size_t)kc_ : 126;
  int col = (layout == 'C' || layout == 'c');
  if (m == 0 || n == 0 || k == 0) return; /* level3.hpp:221 */
#pragma omp parallel for collapse(2) schedule(static)
  for (size_t i = 0; i < m; ++i) {
    for (size_t j = 0; j < n; ++j) {
      size_t cidx = col ? j * ldc + i : i * ldc + j;
      Q c = C[cidx];
      for (size_t kk = 0; kk < k; kk += kc) {
        size_t ke = kk + kc < k ? kk + kc : k;
        Q s = 0.0Q;
        for (size_t l = kk; l < ke; ++l) {
          Q a = col ? A[l * lda + i] : A[i * lda + l];
          Q b = col ? B[j * ldb + l] : B[l * ldb + j];
          s = fmaq(a, b, s);
        }
        c = fmaq(alpha, s, (kk == 0 ? beta : 1.0Q) * c);
      }
      C[cidx] = c;
    }
  }
}

/* Extension (NOT reference behaviour): honour transa/transb.  The reference ignores them
 * (c_interface.hpp:109-112).  op(A) is m x k, op(B) is k x n; same panel formula. */
void orc_gemm_trans(char layout, char ta, char tb, long m_, long n_, long k_, const void *alpha_,
                    const void *A_, long lda_, const void *B_, long ldb_, const void *beta_, void *C_,
                    long ldc_, long kc_)
{
  const Q alpha = *(const Q *)alpha_, beta = *(const Q *)beta_;
  const Q *A = A_, *B = B_; Q *C = C_;
  size_t m = (size_t)m_, n = (size_t)n_, k = (size_t)k_, lda = (size_t)lda_, ldb = (size_t)ldb_, ldc = (size_t)ldc_;
  size_t kc = kc_ > 0 ? (size_t)kc_ : 126;
  int col = (layout == 'C' || layout == 'c');
  int tA = (ta == 'T' || ta == 't' || ta == 'C' || ta == 'c');
  int tB = (tb == 'T' || tb == 't' || tb == 'C' || tb == 'c');
  if (m == 0 || n == 0 || k == 0) return;
#pragma omp parallel for collapse(2) schedule(static)
  for (size_t i = 0; i < m; ++i) {
    for (size_t j = 0; j < n; ++j) {
      size_t cidx = col ? j * ldc + i : i * ldc + j;
      Q c = C[cidx];
      for (size_t kk = 0; kk < k; kk += kc) {
        size_t ke = kk + kc < k ? kk + kc : k;
        Q s = 0.0Q;
        for (size_t l = kk; l < ke; ++l) {
          /* element (i,l) of op(A); storage is (col != tA) ? column-walk : row-walk */
          Q a = (col != tA) ? A[l * lda + i] : A[i * lda + l];
          Q b = (col != tB) ? B[j * ldb + l] : B[l * ldb + j];
          s = fmaq(a, b, s);
        }
        c = fmaq(alpha, s, (kk == 0 ? beta : 1.0Q) * c);
      }
      C[cidx] = c;
    }
  }
}

/* quadblas_qgemm marshalling c_interface.hpp:95-117: double alpha/beta, trans IGNORED. */
void orc_c_qgemm(char layout, char ta, char tb, int m, int n, int k, double alpha, const void *A, int lda,
                 const void *B, int ldb, double beta, void *C, int ldc)
{
  (void)ta; (void)tb;
  Q qa = (Q)alpha, qb = (Q)beta;
  orc_gemm(layout, m, n, k, &qa, A, lda, B, ldb, &qb, C, ldc, 126);
}

/* ------------------------------------------------------------------ sampled GEMM check */
/* Recompute `ns` sampled C entries (idx[2*t], idx[2*t+1]) = (i,j) in reference order; used by the
 * full-size parity tests where recomputing all of C on the CPU would take hours. */
void orc_gemm_sample(char layout, long m_, long n_, long k_, const void *alpha_, const void *A_, long lda_,
                     const void *B_, long ldb_, const void *beta_, const void *Cin_, long ldc_, long kc_,
                     long ns, const int64_t *idx, void *out_)
{
  const Q alpha = *(const Q *)alpha_, beta = *(const Q *)beta_;
  const Q *A = A_, *B = B_, *Cin = Cin_; Q *out = out_;
  size_t k = (size_t)k_, lda = (size_t)lda_, ldb = (size_t)ldb_, ldc = (size_t)ldc_;
  size_t kc = kc_ > 0 ? (size_t)kc_ : 126;
  int col = (layout == 'C' || layout == 'c');
  (void)m_; (void)n_;
#pragma omp parallel for schedule(dynamic, 1)
  for (long t = 0; t < ns; ++t) {
    size_t i = (size_t)idx[2 * t], j = (size_t)idx[2 * t + 1];
    Q c = Cin ? Cin[col ? j * ldc + i : i * ldc + j] : 0.0Q;
    for (size_t kk = 0; kk < k; kk += kc) {
      size_t ke = kk + kc < k ? kk + kc : k;
      Q s = 0.0Q;
      for (size_t l = kk; l < ke; ++l) {
        Q a = col ? A[l * lda + i] : A[i * lda + l];
        Q b = col ? B[j * ldb + l] : B[l * ldb + j];
        s = fmaq(a, b, s);
      }
      c = fmaq(alpha, s, (kk == 0 ? beta : 1.0Q) * c);
    }
    out[t] = c;
  }
}

/* |A||B| row (i,j) in binary128, for the fast-mode forward error bound gamma_k * (|A||B|)_ij. */
void orc_absdot_sample(char layout, long k_, const void *A_, long lda_, const void *B_, long ldb_,
                       long ns, const int64_t *idx, void *out_)
{
  const Q *A = A_, *B = B_; Q *out = out_;
  size_t k = (size_t)k_, lda = (size_t)lda_, ldb = (size_t)ldb_;
  int col = (layout == 'C' || layout == 'c');
#pragma omp parallel for schedule(dynamic, 1)
  for (long t = 0; t < ns; ++t) {
    size_t i = (size_t)idx[2 * t], j = (size_t)idx[2 * t + 1];
    Q s = 0.0Q;
    for (size_t l = 0; l < k; ++l) {
      Q a = col ? A[l * lda + i] : A[i * lda + l];
      Q b = col ? B[j * ldb + l] : B[l * ldb + j];
      s += fabsq(a) * fabsq(b);
    }
    out[t] = s;
  }
}

/* ------------------------------------------------------------------ exact inner products (long accumulator) */
/* The checker of the FAST mode: fast mode may re-associate, so its reference is not the rounding
 * order of level3.hpp:77-85 but the exact inner product.  Every product of two binary128 values is added without rounding into a
 * fixed-point two's-complement accumulator that spans the whole exponent range (66 560 bits, Kulisch style), which gives
 *   exact[t]  = RNE(sum_l a_l b_l)                      (one rounding; gradual underflow and overflow as IEEE),
 *   ratio[t]  = |got[t] - sum_l a_l b_l| / (gamma_k sum_l |a_l||b_l|),  gamma_k = k u / (1 - k u), u = 2^-113
 *               (the fast-mode contract holds iff ratio <= 1; computed from the exact difference, rounded to double at the end),
 *   klass[t]  = 0 finite data; otherwise the IEEE class of the sum of products: 1 NaN, 2 +Inf, 3 -Inf (ratio is then 0 when
 *               got[t] is of that class and +Inf otherwise).
 * Plain C on 64-bit words; test infrastructure like the rest of this file. */
#define LACC_WORDS 1040
#define LACC_OFF 33024            /* bit 0 of the accumulator is 2^-33024 <= the smallest product bit 2^-32988 */
typedef struct { uint64_t w[LACC_WORDS]; } lacc;
static void lacc_add_shifted(lacc *a, const uint64_t p[5], int word, int negative)
{
  if (!negative) {
    unsigned __int128 c = 0;
    for (int i = 0; i < 5; ++i) { c += (unsigned __int128)a->w[word + i] + p[i]; a->w[word + i] = (uint64_t)c; c >>= 64; }
    for (int i = word + 5; c && i < LACC_WORDS; ++i) { c += a->w[i]; a->w[i] = (uint64_t)c; c >>= 64; }
  } else {
    uint64_t b = 0;
    for (int i = 0; i < 5; ++i) {
      const uint64_t x = a->w[word + i], y = p[i];
      const uint64_t d = x - y - b;
      b = (x < y) || (x == y && b) ? 1 : 0;
      a->w[word + i] = d;
    }
    for (int i = word + 5; b && i < LACC_WORDS; ++i) { const uint64_t x = a->w[i]; a->w[i] = x - 1; b = x == 0; }
  }
}
/* acc += (-1)^neg * ma * mb * 2^e with ma, mb < 2^113 given as (hi, lo) */
static void lacc_add_product(lacc *a, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, int e, int neg)
{
  uint64_t p[4] = {0, 0, 0, 0};
  {
    unsigned __int128 t = (unsigned __int128)al * bl; p[0] = (uint64_t)t; unsigned __int128 c = t >> 64;
    t = (unsigned __int128)al * bh; unsigned __int128 t2 = (unsigned __int128)ah * bl;
    c += (uint64_t)t; c += (uint64_t)t2; p[1] = (uint64_t)c; c >>= 64;
    c += (t >> 64); c += (t2 >> 64);
    t = (unsigned __int128)ah * bh; c += (uint64_t)t; p[2] = (uint64_t)c; c >>= 64;
    c += (t >> 64); p[3] = (uint64_t)c;
  }
  const int sh = e + LACC_OFF, word = sh >> 6, bit = sh & 63;
  uint64_t q[5];
  if (bit) { q[0] = p[0] << bit; q[1] = (p[1] << bit) | (p[0] >> (64 - bit)); q[2] = (p[2] << bit) | (p[1] >> (64 - bit)); q[3] = (p[3] << bit) | (p[2] >> (64 - bit)); q[4] = p[3] >> (64 - bit); }
  else { q[0] = p[0]; q[1] = p[1]; q[2] = p[2]; q[3] = p[3]; q[4] = 0; }
  lacc_add_shifted(a, q, word, neg);
}
/* unpack a finite binary128: value = (-1)^s (hi:lo) 2^e; returns 0 finite, 1 NaN, 2 Inf */
static int q_unpack_bits(const Q *x, uint64_t *hi, uint64_t *lo, int *e, int *s)
{
  uint64_t w[2];
  __builtin_memcpy(w, x, 16);
  const int ef = (int)((w[1] >> 48) & 0x7fff);
  *s = (int)(w[1] >> 63);
  *lo = w[0]; *hi = w[1] & 0x0000ffffffffffffULL;
  if (ef == 0x7fff) return (*hi | *lo) ? 1 : 2;
  if (ef) { *hi |= 0x0001000000000000ULL; *e = ef - 16383 - 112; } else *e = 1 - 16383 - 112;
  return 0;
}
static int lacc_negate_if_negative(lacc *a)
{
  if (!(a->w[LACC_WORDS - 1] >> 63)) return 0;
  uint64_t c = 1;
  for (int i = 0; i < LACC_WORDS; ++i) { const uint64_t v = ~a->w[i] + c; c = (c && v == 0) ? 1 : 0; a->w[i] = v; }
  return 1;
}
static int lacc_top_bit(const lacc *a) /* -1 for zero */
{
  for (int i = LACC_WORDS - 1; i >= 0; --i) if (a->w[i]) return 64 * i + 63 - __builtin_clzll(a->w[i]);
  return -1;
}
static int lacc_bit(const lacc *a, int pos) { return pos < 0 ? 0 : (int)((a->w[pos >> 6] >> (pos & 63)) & 1); }
/* RNE of a non-negative accumulator to binary128 (sign applied by the caller) */
static Q lacc_round(const lacc *a)
{
  const int top = lacc_top_bit(a);
  if (top < 0) return 0.0Q;
  int lsb = top - 112;
  const int sub = -16494 + LACC_OFF;          /* position of the least significant subnormal bit */
  if (lsb < sub) lsb = sub;
  unsigned __int128 M = 0;
  for (int pos = top; pos >= lsb; --pos) M = (M << 1) | (unsigned)lacc_bit(a, pos);
  const int guard = lacc_bit(a, lsb - 1);
  int sticky = 0;
  if (lsb - 2 >= 0) {
    const int pw = (lsb - 2) >> 6;
    for (int i = 0; i < pw; ++i) sticky |= a->w[i] != 0;
    const int nb = ((lsb - 2) & 63) + 1;
    sticky |= (a->w[pw] & (nb == 64 ? ~0ULL : ((1ULL << nb) - 1))) != 0;
  }
  if (guard && (sticky || (M & 1))) M += 1;
  const Q mq = (Q)(uint64_t)(M >> 64) * 18446744073709551616.0Q + (Q)(uint64_t)M;   /* <= 114 bits, the top one alone when 114: exact */
  return ldexpq(mq, lsb - LACC_OFF);
}
/* non-negative accumulator = m * 2^(*e) with m < 2^64 + 1 rounded UP (relative error < 2^-62); 0 -> returns 0 */
static double lacc_to_scaled(const lacc *a, int *e)
{
  const int top = lacc_top_bit(a);
  *e = 0;
  if (top < 0) return 0.0;
  uint64_t m = 0;
  for (int pos = top; pos > top - 64 && pos >= 0; --pos) m = (m << 1) | (unsigned)lacc_bit(a, pos);
  const int used = top >= 63 ? 64 : top + 1;
  *e = top + 1 - used - LACC_OFF;
  return (double)m + 1.0;
}
void orc_exact_dot_check(char layout, long k_, const void *A_, long lda_, const void *B_, long ldb_, long ns, const int64_t *idx,
                         const void *got_, void *exact_, double *ratio, int *klass)
{
  const Q *A = A_, *B = B_, *got = got_; Q *exact = exact_;
  const size_t k = (size_t)k_, lda = (size_t)lda_, ldb = (size_t)ldb_;
  const int col = (layout == 'C' || layout == 'c');
#pragma omp parallel
  {
    lacc *acc = (lacc *)__builtin_malloc(sizeof(lacc)), *ab = (lacc *)__builtin_malloc(sizeof(lacc));
#pragma omp for schedule(dynamic, 1)
    for (long t = 0; t < ns; ++t) {
      const size_t i = (size_t)idx[2 * t], j = (size_t)idx[2 * t + 1];
      __builtin_memset(acc, 0, sizeof(lacc)); __builtin_memset(ab, 0, sizeof(lacc));
      int nan = 0, pinf = 0, ninf = 0;
      for (size_t l = 0; l < k; ++l) {
        const Q *a = col ? &A[l * lda + i] : &A[i * lda + l];
        const Q *b = col ? &B[j * ldb + l] : &B[l * ldb + j];
        uint64_t ah, al, bh, bl; int ea, eb, sa, sb;
        const int ca = q_unpack_bits(a, &ah, &al, &ea, &sa), cb = q_unpack_bits(b, &bh, &bl, &eb, &sb);
        if (ca || cb) {
          const int za = !ca && !(ah | al), zb = !cb && !(bh | bl);
          if (ca == 1 || cb == 1 || za || zb) nan = 1;          /* NaN operand or Inf * 0 */
          else if (sa ^ sb) ninf = 1; else pinf = 1;
          continue;
        }
        if (!(ah | al) || !(bh | bl)) continue;
        lacc_add_product(acc, ah, al, bh, bl, ea + eb, sa ^ sb);
        lacc_add_product(ab, ah, al, bh, bl, ea + eb, 0);
      }
      const Q g = got ? got[t] : 0.0Q;
      if (nan || pinf || ninf) {
        const int kl = (nan || (pinf && ninf)) ? 1 : (pinf ? 2 : 3);
        klass[t] = kl;
        if (exact) exact[t] = kl == 1 ? nanq("") : (kl == 2 ? HUGE_VALQ : -HUGE_VALQ);
        if (ratio) ratio[t] = !got ? 0.0 : ((kl == 1 ? isnanq(g) : (isinfq(g) && ((g < 0) == (kl == 3)))) ? 0.0 : HUGE_VAL);
        continue;
      }
      klass[t] = 0;
      {
        lacc tmp = *acc;
        const int neg = lacc_negate_if_negative(&tmp);
        const Q r = lacc_round(&tmp);
        if (exact) exact[t] = neg ? -r : r;
      }
      if (ratio) {
        if (!got || isnanq(g) || isinfq(g)) { ratio[t] = got ? HUGE_VAL : 0.0; continue; }
        uint64_t gh, gl; int ge, gs;
        q_unpack_bits(&g, &gh, &gl, &ge, &gs);
        if (gh | gl) lacc_add_product(acc, gh, gl, 0, 1, ge, !gs);      /* acc -= got (exactly) */
        lacc_negate_if_negative(acc);
        /* gamma_k >= k u: the bound is k u sum|a||b|; mantissas and exponents are kept apart so that sums far outside the range of
         * double (2^-32000 .. 2^32000) still compare; the 2^-50 slack covers the two round-ups */
        int ee, eb;
        const double me = lacc_to_scaled(acc, &ee), mb = lacc_to_scaled(ab, &eb) * (1.0 - 0x1p-50);
        if (me == 0.0) ratio[t] = 0.0;
        else if (mb == 0.0) ratio[t] = HUGE_VAL;
        else {
          int d = ee - (eb - 113);
          if (d > 2000) d = 2000;
          if (d < -2000) d = -2000;
          ratio[t] = ldexp(me / (mb * (double)k), d);
        }
      }
    }
    __builtin_free(acc); __builtin_free(ab);
  }
}

const char *orc_arith(void) { return "gcc __float128: libgcc soft-fp add/mul, libquadmath fmaq/sqrtq"; }

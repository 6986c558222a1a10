/*
 * tests/host/q128_host_test.cpp — CPU bit-exactness harness for qblas_b200/csrc/q128.cuh.
 *
 * TEST INFRASTRUCTURE: compiles the product's integer-limb binary128 core with g++ (the header is
 * host/device dual) and compares every result bitwise with GCC __float128 / libquadmath, which is
 * what the oracle (oracle/qoracle.c) and the shimmed reference use.  This lets the soft-float be
 * validated on hundreds of millions of vectors without a GPU; the GPU tests then re-check the
 * same routines as compiled by nvcc.
 *
 * usage: q128_host_test [n_per_regime] [seed]     exit code = number of failing regimes
 */
#include <quadmath.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <omp.h>
#define QB_CHAIN_STATS 1
#include "../../qblas_b200/csrc/q128.cuh"
#include "../../qblas_b200/csrc/q128_chain.cuh"

typedef __float128 Q;

static inline uint64_t splitmix(uint64_t &s)
{
  uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static inline q128 mk(uint32_t sign, uint32_t e, uint64_t mh, uint64_t ml)
{ q128 r; r.hi = ((uint64_t)sign << 63) | ((uint64_t)(e & 0x7fff) << 48) | (mh & 0xffffffffffffULL); r.lo = ml; return r; }
static inline Q toQ(q128 a) { Q r; memcpy(&r, &a, 16); return r; }
static inline q128 fromQ(Q a) { q128 r; memcpy(&r, &a, 16); return r; }
static inline bool isnan_q(q128 a) { return qb::q_is_nan(a); }
static inline bool same(q128 a, q128 b) { if (isnan_q(a) && isnan_q(b)) return true; return a.lo == b.lo && a.hi == b.hi; }

/* mantissa generators */
static inline void mant(uint64_t &s, int kind, uint64_t &mh, uint64_t &ml)
{
  mh = splitmix(s); ml = splitmix(s);
  switch (kind & 7) {
  case 0: break;                                         /* full 112 random bits */
  case 1: ml &= 0xF000000000000000ULL; break;            /* a double cast to quad (52 bits) */
  case 2: ml = 0; mh &= 0xffffff000000ULL; break;        /* float-like */
  case 3: ml = ~0ULL; mh |= 0xffffffULL; break;          /* long runs of ones */
  case 4: ml = 0; mh = 0; break;                         /* powers of two */
  case 5: ml = (ml & 1) ; mh = 0; break;                 /* 1 + tiny */
  case 6: ml |= 0x7fffffffffffULL; break;                /* trailing ones */
  case 7: ml &= ~0xffffffffffULL; break;                 /* trailing zeros */
  }
}

struct Regime { const char *name; int id; };

static q128 gen(uint64_t &s, int regime, int which /*0=a,1=b,2=c*/, q128 a, q128 b)
{
  uint64_t mh, ml; mant(s, (int)(splitmix(s) & 7) < 5 ? 0 : (int)(splitmix(s) & 7), mh, ml);
  uint32_t sign = splitmix(s) & 1;
  uint32_t e;
  uint64_t r = splitmix(s);
  switch (regime) {
  default:
  case 0: e = 1 + r % 0x7ffe; break;                                      /* anything finite normal */
  case 1: e = 16383 - 40 + r % 80; break;                                 /* similar magnitudes */
  case 2: /* c close to -a*b exponent: cancellation */
    if (which < 2) e = 16383 - 20 + r % 40;
    else {
      int ea = (a.hi >> 48) & 0x7fff, eb = (b.hi >> 48) & 0x7fff;
      e = (uint32_t)(ea + eb - 16383 + (int)(r % 5) - 2);
      sign = ((a.hi ^ b.hi) >> 63) ^ 1;
    }
    break;
  case 3: /* c = -RN(a*b) or neighbours: massive cancellation */
    if (which < 2) e = 16383 - 200 + r % 400;
    else {
      Q p = toQ(a) * toQ(b);
      q128 pq = fromQ(p);
      int64_t dl = (int64_t)(r % 5) - 2;
      pq.lo += (uint64_t)dl; /* neighbours (carry ignored: fine for testing) */
      pq.hi ^= 0x8000000000000000ULL;
      return pq;
    }
    break;
  case 4: e = r % 3 ? (r >> 8) % 120 : 0; break;                          /* tiny / subnormal operands */
  case 5: /* products that land in the subnormal range */
    if (which == 0) e = 1 + (r >> 3) % 16383;
    else if (which == 1) { int ea = (a.hi >> 48) & 0x7fff; int t = 16383 - ea - 60 + (int)(r % 200); e = t < 0 ? 0 : (t > 0x7ffe ? 0x7ffe : t); }
    else e = (r % 4 == 0) ? 0 : (r >> 4) % 140;
    break;
  case 6: e = 0x7ffe - r % 60; break;                                     /* overflow region (a*b overflows) */
  case 7: /* product near overflow threshold, c huge */
    if (which == 0) e = 16383 + (r >> 3) % 16383;
    else if (which == 1) { int ea = (a.hi >> 48) & 0x7fff; int t = 0x7ffe + 16383 - ea - (int)(r % 8) + 2; e = t < 1 ? 1 : (t > 0x7ffe ? 0x7ffe : t); }
    else e = 0x7ffe - r % 130;
    break;
  case 8: { /* specials grid */
    int k = r % 12;
    switch (k) {
    case 0: return mk(sign, 0, 0, 0);
    case 1: return mk(sign, 0x7fff, 0, 0);
    case 2: return mk(sign, 0x7fff, 0x800000000000ULL, 0);
    case 3: return mk(sign, 0x7fff, 1, 5);
    case 4: return mk(sign, 0, 0, 1);
    case 5: return mk(sign, 0, 0xffffffffffffULL, ~0ULL);
    case 6: return mk(sign, 0x7ffe, 0xffffffffffffULL, ~0ULL);
    case 7: return mk(sign, 1, 0, 0);
    case 8: return mk(sign, 16383, 0, 0);
    default: e = 16383 - 3 + r % 6; break;
    }
    break; }
  case 9: /* c dominates by 100..130 bits or product dominates by 100..260 bits: sticky-only paths */
    if (which < 2) e = 16383 - 10 + r % 20;
    else {
      int ea = (a.hi >> 48) & 0x7fff, eb = (b.hi >> 48) & 0x7fff;
      int off = (r & 1) ? 100 + (int)((r >> 1) % 32) : -(100 + (int)((r >> 1) % 170));
      e = (uint32_t)(ea + eb - 16383 + off);
    }
    break;
  case 10: /* exponent gaps 0..120 both ways, same-sign and opposite */
    if (which < 2) e = 16383 - 10 + r % 20;
    else {
      int ea = (a.hi >> 48) & 0x7fff, eb = (b.hi >> 48) & 0x7fff;
      e = (uint32_t)(ea + eb - 16383 + (int)(r % 241) - 120);
    }
    break;
  }
  return mk(sign, e, mh, ml);
}

/* regime 11: exact ties and near-ties, where only the sticky/jam logic decides the rounding */
static void gen_ties(uint64_t &s, q128 &a, q128 &b, q128 &c)
{
  uint64_t r = splitmix(s);
  uint32_t sa = r & 1, sb = (r >> 1) & 1, sc = (r >> 2) & 1;
  int fam = (r >> 3) & 1;
  if (fam == 0) {
    /* c arbitrary, |a*b| = 2^(ec-112-g) * (1 + eps), g in {0,1,2}, eps in {0, tiny, ~2^-112..} */
    uint64_t mh, ml; mant(s, (int)((r >> 4) & 7), mh, ml);
    uint32_t ec = 16383 - 100 + (uint32_t)((r >> 8) % 200);
    c = mk(sc, ec, mh, ml);
    int g = (int)((r >> 20) % 3);
    int ea = 16383 - 30 + (int)((r >> 24) % 60);
    int eb = (int)ec - 112 - g - ea + 16383;
    uint64_t amh = 0, aml = 0, bmh = 0, bml = 0;
    switch ((r >> 32) & 7) {
    case 0: break;
    case 1: aml = 1; break;
    case 2: aml = 1; bml = 1; break;
    case 3: bml = 1ULL << (splitmix(s) & 63); break;
    case 4: amh = 1ULL << (splitmix(s) % 48); break;
    case 5: aml = splitmix(s); break;
    case 6: amh = 0x800000000000ULL; break;       /* 1.5 */
    case 7: amh = 0x800000000000ULL; bml = 1; break;
    }
    a = mk(sa, (uint32_t)ea, amh, aml);
    b = mk(sb, (uint32_t)eb, bmh, bml);
  } else {
    /* a = 1 + 2^-i, b = 1 + 2^-j, i + j = 113 (+-1): product has a 1 exactly at the guard bit */
    int i = 1 + (int)((r >> 4) % 111), j = 113 - i + (int)((r >> 12) % 3) - 1;
    if (j < 1) j = 1; if (j > 112) j = 112;
    auto bitpos = [](int t, uint64_t &mh, uint64_t &ml) { int pos = 112 - t; mh = pos >= 64 ? 1ULL << (pos - 64) : 0; ml = pos < 64 ? 1ULL << pos : 0; };
    uint64_t amh, aml, bmh, bml; bitpos(i, amh, aml); bitpos(j, bmh, bml);
    int ea = 16383 - 50 + (int)((r >> 16) % 100), eb = 16383 - 50 + (int)((r >> 24) % 100);
    a = mk(sa, (uint32_t)ea, amh, aml); b = mk(sb, (uint32_t)eb, bmh, bml);
    int ep = ea + eb - 16383;
    int k = (int)((r >> 32) % 4);
    if (k == 0) c = mk(sc, 0, 0, 0);
    else {
      uint64_t mh, ml; mant(s, (int)((r >> 40) & 7), mh, ml);
      int ec = ep - 114 - (int)((r >> 44) % 300);
      if (ec < 0) ec = 0;
      c = mk(sc, (uint32_t)ec, mh, ml);
    }
  }
}

/* Chains through the branch-free step the way k_gemm_nb runs them (qb_level3.cu): s starts at +0, every step that declines is redone
 * by the generic q_fma from the staged operands (gemm_redo), the accumulator carries over in its unpacked / packed forms; panels of
 * `kc` steps are folded with C = fma(alpha, s, beta * C).  Must give the bits of the same chain through libquadmath.  The inputs mix
 * similar magnitudes (long fast stretches), exact zeros of both signs (leading zeros of a triangular row), cancelling pairs and
 * a few subnormals / Inf / NaN. */
static long chain_test(long nchains, uint64_t seed, long *declined_out, long *steps_out)
{
  long bad = 0, declined = 0, steps = 0;
#pragma omp parallel reduction(+ : bad, declined, steps)
  {
    uint64_t s = seed * 7919ULL + (uint64_t)omp_get_thread_num() * 104729ULL + 1;
#pragma omp for schedule(static)
    for (long c = 0; c < nchains; ++c) {
      const int flavour = (int)(splitmix(s) % 6);
      const int len = 1 + (int)(splitmix(s) % 300), kc = (splitmix(s) & 1) ? 126 : 40;
      const int lead = (flavour == 1) ? (int)(splitmix(s) % (uint64_t)len) : 0;
      q128 alpha = gen(s, 1, 0, q128(), q128()), beta = gen(s, 1, 1, alpha, q128()), C = gen(s, 1, 2, alpha, beta), Cq = C;
      if (flavour == 5) beta = mk(splitmix(s) & 1, 0, 0, 0);
      qb::qacc2 acc = qb::qacc2_zero();
      Q sq = 0;
      int pc = 0, first = 1;
      q128 pa = q128(), pb = q128();
      for (int l = 0; l < len; ++l) {
        q128 a = gen(s, 1, 0, q128(), q128()), b = gen(s, 1, 1, a, q128());
        const uint64_t r = splitmix(s);
        if (flavour <= 2 || flavour == 5) {   /* magnitudes within 2^8 of each other: long stretches of undeclined steps */
          a.hi = (a.hi & 0x8000ffffffffffffULL) | ((uint64_t)(16383 - 4 + (r >> 8) % 8) << 48);
          b.hi = (b.hi & 0x8000ffffffffffffULL) | ((uint64_t)(16383 - 4 + (r >> 16) % 8) << 48);
        }
        if (l < lead) a = mk(r & 1, 0, 0, 0);                                   /* triangular row: +-0 products into a zero accumulator */
        else if (flavour == 2 && (l & 1)) { a = pa; b = pb; b.hi ^= 0x8000000000000000ULL; }   /* exact cancellation of the step before */
        else if (flavour == 3 && r % 16 == 0) a = gen(s, 8, 0, q128(), q128());   /* specials */
        else if (flavour == 4 && r % 8 == 0) b = mk(r >> 5 & 1, 0, 0, 0);
        else if (flavour == 4 && r % 8 == 1) a = gen(s, 0, 0, q128(), q128());    /* any exponent: products far above / below the sum */
        pa = a; pb = b;
        const qb::qstaged A = qb::qstage(a), B = qb::qstage(b);
        uint32_t col[24] = {0};
        qb::qnbctx cx; cx.col = col; cx.stride = 2; cx.zero = 0u;
        ++steps;
        if (qb::qacc_fma_nb(acc, A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, A.meta + B.meta, cx)) {
          ++declined;
          acc = qb::qacc2_from(qb::q_fma(qb::qstaged_pack(A), qb::qstaged_pack(B), qb::qacc2_pack(acc)));
        }
        sq = fmaq(toQ(a), toQ(b), sq);
        if (++pc == kc || l == len - 1) {
          const q128 sp = qb::qacc2_pack(acc);
          C = qb::q_fma(alpha, sp, first ? qb::q_mul(beta, C) : C);
          Cq = fromQ(fmaq(toQ(alpha), sq, first ? toQ(beta) * toQ(Cq) : toQ(Cq)));
          if (!same(sp, fromQ(sq)) || !same(C, Cq)) { ++bad; break; }
          acc = qb::qacc2_zero(); sq = 0; pc = 0; first = 0;
        }
      }
    }
  }
  *declined_out = declined; *steps_out = steps;
  return bad;
}

int main(int argc, char **argv)
{
  long n = argc > 1 ? atol(argv[1]) : 2000000;
  uint64_t seed0 = argc > 2 ? strtoull(argv[2], 0, 0) : 12345;
  const char *names[] = {"any-normal", "similar-mag", "cancel-exp", "massive-cancel", "tiny-operands",
                         "subnormal-results", "overflow", "near-overflow", "specials", "sticky-only", "gap-sweep", "ties"};
  int nreg = 12, failed = 0;
  long total = 0;
  for (int reg = 0; reg < nreg; ++reg) {
    long bad_fma = 0, bad_mul = 0, bad_add = 0, bad_sqrt = 0, bad_cast = 0, bad_chain = 0, bad_nb = 0, nb_total = 0, nb_declined = 0;
#pragma omp parallel reduction(+ : bad_fma, bad_mul, bad_add, bad_sqrt, bad_cast, bad_chain, bad_nb, nb_total, nb_declined)
    {
      uint64_t s = seed0 * 1000003ULL + (uint64_t)reg * 7919ULL + (uint64_t)omp_get_thread_num() * 104729ULL;
      bool printed = false;
#pragma omp for schedule(static)
      for (long i = 0; i < n; ++i) {
        q128 a, b, c;
        if (reg == 11) gen_ties(s, a, b, c);
        else { a = gen(s, reg, 0, q128(), q128()); b = gen(s, reg, 1, a, q128()); c = gen(s, reg, 2, a, b); }
        q128 r = qb::q_fma(a, b, c);
        q128 e = fromQ(fmaq(toQ(a), toQ(b), toQ(c)));
        if (!same(r, e)) {
          ++bad_fma;
          if (!printed) {
            printed = true;
#pragma omp critical
            printf("  FMA mismatch [%s] a=%016llx%016llx b=%016llx%016llx c=%016llx%016llx got=%016llx%016llx exp=%016llx%016llx\n",
                   names[reg], (unsigned long long)a.hi, (unsigned long long)a.lo, (unsigned long long)b.hi, (unsigned long long)b.lo,
                   (unsigned long long)c.hi, (unsigned long long)c.lo, (unsigned long long)r.hi, (unsigned long long)r.lo,
                   (unsigned long long)e.hi, (unsigned long long)e.lo);
          }
        }
        /* chain-form fma (the kernels' primitive) must agree with the generic one */
        {
          q128 rc = qb::q_fma_fast(a, b, c);
          uint32_t col[24] = {0};
          q128 rc2 = qb::q_fma_fast_sc(a, b, c, col, 2);
          if (!same(rc2, e)) rc = rc2;
          if (!same(rc, e)) {
            ++bad_chain;
            if (!printed) {
              printed = true;
#pragma omp critical
              printf("  CHAIN mismatch [%s] a=%016llx%016llx b=%016llx%016llx c=%016llx%016llx got=%016llx%016llx exp=%016llx%016llx\n",
                     names[reg], (unsigned long long)a.hi, (unsigned long long)a.lo, (unsigned long long)b.hi, (unsigned long long)b.lo,
                     (unsigned long long)c.hi, (unsigned long long)c.lo, (unsigned long long)rc.hi, (unsigned long long)rc.lo,
                     (unsigned long long)e.hi, (unsigned long long)e.lo);
            }
          }
        }
        /* branch-free step of the reference-order qgemm: either the same bits, or "not handled" with the accumulator untouched */
        {
          uint32_t col[24] = {0}, nb_bad = 0;
          q128 rn = qb::q_fma_fast_nb(a, b, c, col, 2, &nb_bad);
          bool negzero = (c.hi == 0x8000000000000000ULL && c.lo == 0);
          ++nb_total;
          if (nb_bad) { ++nb_declined; if (!same(rn, c)) ++bad_nb; }
          else if (!same(rn, e)) {
            ++bad_nb;
            if (!printed) {
              printed = true;
#pragma omp critical
              printf("  NB mismatch [%s] a=%016llx%016llx b=%016llx%016llx c=%016llx%016llx got=%016llx%016llx exp=%016llx%016llx\n",
                     names[reg], (unsigned long long)a.hi, (unsigned long long)a.lo, (unsigned long long)b.hi, (unsigned long long)b.lo,
                     (unsigned long long)c.hi, (unsigned long long)c.lo, (unsigned long long)rn.hi, (unsigned long long)rn.lo,
                     (unsigned long long)e.hi, (unsigned long long)e.lo);
            }
          }
          (void)negzero;
        }
        if ((i & 3) == 0) {
          q128 rm = qb::q_mul(a, b), em = fromQ(toQ(a) * toQ(b));
          if (!same(rm, em)) ++bad_mul;
          q128 ra = qb::q_add(a, c), ea = fromQ(toQ(a) + toQ(c));
          if (!same(ra, ea)) { ++bad_add; }
          q128 rs2 = qb::q_sub(a, c), es2 = fromQ(toQ(a) - toQ(c));
          if (!same(rs2, es2)) { ++bad_add; }
        }
        if ((i & 31) == 0) {
          q128 rs = qb::q_sqrt(a), es = fromQ(sqrtq(toQ(a)));
          if (!same(rs, es)) {
            ++bad_sqrt;
            if (!printed) { printed = true;
#pragma omp critical
              printf("  SQRT mismatch a=%016llx%016llx got=%016llx%016llx exp=%016llx%016llx\n", (unsigned long long)a.hi, (unsigned long long)a.lo,
                     (unsigned long long)rs.hi, (unsigned long long)rs.lo, (unsigned long long)es.hi, (unsigned long long)es.lo); }
          }
          /* casts: quad -> double (RNE) on a value squeezed into double range, and back */
          q128 t = a;
          uint32_t ee = (uint32_t)(splitmix(s) % 2300);
          t.hi = (t.hi & 0x8000ffffffffffffULL) | ((uint64_t)(16383 - 1150 + ee) << 48);
          double dd = (double)toQ(t);
          uint64_t db; memcpy(&db, &dd, 8);
          uint64_t got = qb::q_to_double_bits(t);
          bool dn = (dd != dd);
          if (!(dn ? ((got & 0x7ff0000000000000ULL) == 0x7ff0000000000000ULL && (got & 0xfffffffffffffULL)) : got == db)) ++bad_cast;
          uint64_t rb = splitmix(s);
          if ((splitmix(s) & 7) == 0) rb &= 0x800fffffffffffffULL; /* double subnormals */
          double rd; memcpy(&rd, &rb, 8);
          q128 w = qb::q_from_double_bits(rb), we = fromQ((Q)rd);
          if (!same(w, we)) ++bad_cast;
        }
      }
    }
    total += n;
    double fastfrac = (double)qb_chain_fast_hits / (double)(qb_chain_fast_hits + qb_chain_slow_hits + 1e-9);
    qb_chain_fast_hits = qb_chain_slow_hits = 0;
    bool ok = !(bad_fma | bad_mul | bad_add | bad_sqrt | bad_cast | bad_chain | bad_nb);
    printf("%-18s n=%ld fma_bad=%ld chain_bad=%ld mul_bad=%ld add_bad=%ld sqrt_bad=%ld cast_bad=%ld fast=%.3f nb_bad=%ld nb_fast=%.3f %s\n", names[reg], n, bad_fma,
           bad_chain, bad_mul, bad_add, bad_sqrt, bad_cast, fastfrac, bad_nb, 1.0 - (double)nb_declined / (double)(nb_total + 1e-9), ok ? "OK" : "FAIL");
    if (!ok) ++failed;
  }
  {
    long declined = 0, steps = 0;
    const long nch = n / 40 + 1, badc = chain_test(nch, seed0, &declined, &steps);
    printf("%-18s chains=%ld steps=%ld declined=%.4f chain_mismatches=%ld %s\n", "nb-chains", nch, steps, (double)declined / (double)(steps + 1e-9), badc,
           badc ? "FAIL" : "OK");
    if (badc) ++failed;
  }
  printf("TOTAL fma vectors: %ld, failing regimes: %d\n", total, failed);
  return failed;
}

// g++ build of qblas_b200/csrc/qb_crt.cuh (host/device dual source) for tests/test_host_crt.py: the
// residue, accumulator-reduction and reconstruction arithmetic of the tensor-path qgemm, checked
// against Python integers.  Test infrastructure only.
#define QCRT_HOST_EMULATE_PTX 1   // compile the device form of the reconstruction too, carry flag emulated
#include "../../qblas_b200/csrc/qb_crt.cuh"
#include <algorithm>
#include <vector>
using namespace qb::crt;

static Tables g_T;
static bool g_have = false;
static const Tables &tab() { if (!g_have) { host::build_tables(g_T); g_have = true; } return g_T; }

static int g_form = 0;   // 0: reconstruct (reference form), 1: reconstruct_dev (the kernel's carry-chain form)
template <int NG> static void rec(const uint32_t (&r)[NMP], const Plan &pl, uint32_t *mag, uint32_t *neg)
{
  uint32_t Y[NG + 1];
  if (g_form) reconstruct_dev<NG>(r, pl, Y, *neg);
  else reconstruct<NG>(r, pl, Y, *neg);
  for (int l = 0; l < NLMAX; ++l) mag[l] = l < NG + 1 ? Y[l] : 0;
}

// ---- the whole residue-scheme qgemm on the CPU, built from the library's own arithmetic (scan as k_oz_scan / k_oz_plan,
// element_words, residue_sym, int32 accumulation as the tensor kernel, acc_mod, reconstruct_dev, limbs_to_q, q_fma / q_mul)
template <int NW> static void residues_of(const q128 &x, int base, int N, int8_t *out, int stride)
{
  uint32_t w[NWMAX], sign;
  element_words<NW>(x, base, w, sign);
  for (int i = 0; i < N; ++i) out[(size_t)i * stride] = (int8_t)(residue_sym<NW>(w, sign, i, tab()) & 0xffu);
}
static void residues_dispatch(int nw, const q128 &x, int base, int N, int8_t *out, int stride)
{
  switch (nw) {
    case 1: residues_of<1>(x, base, N, out, stride); break;
    case 2: residues_of<2>(x, base, N, out, stride); break;
    case 3: residues_of<3>(x, base, N, out, stride); break;
    case 4: residues_of<4>(x, base, N, out, stride); break;
    case 5: residues_of<5>(x, base, N, out, stride); break;
    default: residues_of<6>(x, base, N, out, stride); break;
  }
}
static void scan_line(const q128 *x, int count, int stride, int &emax, int &lmin, int &special)
{
  emax = 0; lmin = 0x7fffffff;
  for (int l = 0; l < count; ++l) {
    const q128 a = x[(size_t)l * stride];
    const uint32_t ef = (uint32_t)(a.hi >> 48) & 0x7fffu;
    if (ef == 0x7fffu) { special = 1; continue; }
    const uint64_t mhi = (a.hi & qb::Q_MANT_HI_MASK) | (ef ? qb::Q_IMPLICIT : 0);
    if (!(a.lo | mhi)) continue;
    const int ee = ef ? (int)ef : 1;
    const int tz = a.lo ? __builtin_ctzll(a.lo) : 64 + __builtin_ctzll(mhi);
    if (ee > emax) emax = ee;
    if (ee + tz < lmin) lmin = ee + tz;
  }
}
template <int NG> static q128 fold_one(const uint32_t (&r)[NMP], const Plan &pl, int Eb, int *msb)
{
  uint32_t Y[NG + 1], neg;
  reconstruct_dev<NG>(r, pl, Y, neg);
  *msb = limbs_msb<NG + 1>(Y);
  return limbs_to_q<NG + 1>(Y, neg, Eb);
}

extern "C" {
// C (m x n) = alpha * A (m x k) * B (k x n) + beta * C, row-major, dense, as the GPU path computes it: scan, windows from
// crt::host::plan_windows (wcap = the planner's budget per operand), truncating element_words, residues, int32 accumulation,
// acc_mod, reconstruct_dev, and the acceptance test crt::accept_msb of k_crt_fold: rejected[i*n+j] = 1 marks the elements the GPU
// hands to the fix-up kernel (C is left untouched there).  returns 0, or 1 if the scheme declines (Inf/NaN are not mirrored here).
int crt_gemm_host(int m, int n, int k, const q128 *A, const q128 *B, q128 *Cm, const q128 *alpha, const q128 *beta, int wcap, int *info /* N, WA, WB, truncated */,
                  unsigned char *rejected)
{
  std::vector<int> emaxA(m), lminA(m), emaxB(n), lminB(n);
  int special = 0, WA_nat = 0, WB_nat = 0;
  for (int i = 0; i < m; ++i) { scan_line(A + (size_t)i * k, k, 1, emaxA[i], lminA[i], special); if (emaxA[i]) WA_nat = std::max(WA_nat, emaxA[i] + 113 - lminA[i]); }
  for (int j = 0; j < n; ++j) { scan_line(B + j, k, n, emaxB[j], lminB[j], special); if (emaxB[j]) WB_nat = std::max(WB_nat, emaxB[j] + 113 - lminB[j]); }
  if (special) return 1;
  host::Windows win;
  if (!host::plan_windows(WA_nat, WB_nat, k, wcap, win)) return 1;
  const int WA = win.WA, WB = win.WB, N = win.N;
  info[0] = N; info[1] = WA; info[2] = WB; info[3] = (win.truncA ? 1 : 0) | (win.truncB ? 2 : 0);
  Plan pl; host::build_plan(N, pl);
  const int nwA = std::min(NWMAX, std::max(1, (WA + 31) / 32)), nwB = std::min(NWMAX, std::max(1, (WB + 31) / 32));
  std::vector<int8_t> pA((size_t)N * m * k), pB((size_t)N * n * k);     // planes [N][rows][k]
  for (int i = 0; i < m; ++i)
    for (int l = 0; l < k; ++l) residues_dispatch(nwA, A[(size_t)i * k + l], emaxA[i] + 113 - WA, N, &pA[(size_t)i * k + l], m * k);
  for (int j = 0; j < n; ++j)
    for (int l = 0; l < k; ++l) residues_dispatch(nwB, B[(size_t)l * n + j], emaxB[j] + 113 - WB, N, &pB[(size_t)j * k + l], n * k);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      uint32_t r[NMP] = {0};
      for (int c = 0; c < N; ++c) {
        int32_t acc = 0;   // what tcgen05 kind::i8 accumulates
        const int8_t *a = &pA[((size_t)c * m + i) * k], *b = &pB[((size_t)c * n + j) * k];
        for (int l = 0; l < k; ++l) acc += (int32_t)a[l] * (int32_t)b[l];
        r[c] = acc_mod(acc, c, tab());
      }
      const int baseA = emaxA[i] + 113 - WA, baseB = emaxB[j] + 113 - WB;
      const int Eb = baseA + baseB - 2 * 16495;
      q128 s = {0, 0};
      int msb = -1;
      switch (pl.NG) {
#define CASE(g) case g: s = fold_one<g>(r, pl, Eb, &msb); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
      }
      const bool tA = emaxA[i] != 0 && lminA[i] < baseA, tB = emaxB[j] != 0 && lminB[j] < baseB;
      const bool reject = msb < accept_msb(tA, tB, WA, WB, k);
      if (rejected) rejected[(size_t)i * n + j] = reject ? 1 : 0;
      if (reject) continue;
      q128 &c = Cm[(size_t)i * n + j];
      c = qb::q_fma(*alpha, s, qb::q_mul(*beta, c));
    }
  return 0;
}
void crt_set_form(int dev_form) { g_form = dev_form ? 1 : 0; }
int crt_num_moduli(void) { return NM; }
int crt_modulus(int i) { return MODULI[i]; }
int crt_moduli_for_bits(int bits) { return host::moduli_for_bits(bits); }
double crt_plan_bits(int N) { Plan pl; host::build_plan(N, pl); return pl.bits; }
// count elements: words[e][6] little-endian |X|, sign[e]; out[e][N] int8 residues
void crt_residues(int count, const uint32_t *words, const uint32_t *sign, int nw, int N, int8_t *out)
{
  for (int e = 0; e < count; ++e) {
    uint32_t w[NWMAX];
    for (int j = 0; j < NWMAX; ++j) w[j] = words[e * NWMAX + j];
    for (int i = 0; i < N; ++i) {
      uint32_t b = 0;
      switch (nw) {
        case 1: b = residue_byte<1>(w, sign[e], i, tab()); break;
        case 2: b = residue_byte<2>(w, sign[e], i, tab()); break;
        case 3: b = residue_byte<3>(w, sign[e], i, tab()); break;
        case 4: b = residue_byte<4>(w, sign[e], i, tab()); break;
        case 5: b = residue_byte<5>(w, sign[e], i, tab()); break;
        default: b = residue_byte<6>(w, sign[e], i, tab()); break;
      }
      out[e * N + i] = (int8_t)b;
    }
  }
}
// magnitudes mag[e][14] (nl limbs used), neg[e], Eb[e] -> out[e] = round(+-mag * 2^Eb) as binary128 (lo, hi)
void crt_round(int count, int nl, const uint32_t *mag, const uint32_t *neg, const int32_t *Eb, uint64_t *out)
{
  for (int e = 0; e < count; ++e) {
    q128 r = {0, 0};
    switch (nl) {
#define CASE(n) case n: { uint32_t L[n]; for (int l = 0; l < n; ++l) L[l] = mag[e * NLMAX + l]; r = limbs_to_q<n>(L, neg[e], Eb[e]); } break;
      CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13) CASE(14)
#undef CASE
    }
    out[2 * e] = r.lo; out[2 * e + 1] = r.hi;
  }
}
// int32 accumulators acc[e][N] -> residues in [0,p) -> reconstruction: mag[e][14], neg[e]
void crt_fold(int count, const int32_t *acc, int N, uint32_t *mag, uint32_t *neg)
{
  Plan pl; host::build_plan(N, pl);
  for (int e = 0; e < count; ++e) {
    uint32_t r[NMP] = {0};
    for (int i = 0; i < N; ++i) r[i] = acc_mod(acc[e * N + i], i, tab());
    switch (pl.NG) {
#define CASE(g) case g: rec<g>(r, pl, mag + e * NLMAX, neg + e); break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12) CASE(13)
#undef CASE
    }
  }
}
}

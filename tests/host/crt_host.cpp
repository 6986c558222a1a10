// g++ build of qblas_b200/csrc/qb_crt.cuh (host/device dual source) for tests/test_host_crt.py: the
// residue, accumulator-reduction and reconstruction arithmetic of the tensor-path qgemm, checked
// against Python integers.  Test infrastructure only.
#define QCRT_HOST_EMULATE_PTX 1   // compile the device form of the reconstruction too, carry flag emulated
#include "../../qblas_b200/csrc/qb_crt.cuh"
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

extern "C" {
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

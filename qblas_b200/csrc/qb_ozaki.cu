/*
 * qb_ozaki.cu — fast-mode binary128 GEMM on the 5th-generation tensor cores.
 *
 * Two schemes share the scan / plan step and the tcgen05 kernel of this file:
 *   residues (default, qb_set_tensor_scheme(1)): one int8 GEMM per modulus + Chinese-remainder fold, always exact; the
 *            arithmetic is in qb_crt.cuh, the kernels (k_crt_residues*, k_oz_mma<1>, k_crt_fold) and the stream-pipelined
 *            driver (launch_gemm_crt) are in the section "residue scheme" below;
 *   digit diagonals (qb_set_tensor_scheme(0), and the fallback when the moduli cannot cover the operands' bit span):
 *            described next.
 *
 * What it replaces: the arithmetic of QuadBLAS::gemm (/root/reference/include/quadblas/algorithms/
 * level3.hpp:215-336; hot loop :77-85 = one Sleef_fmaq2_u05 per two products) when the library is
 * in QB_MODE_FAST.  Fast mode is free to re-associate (last paragraph); this
 * path goes further and computes every dot product EXACTLY, then rounds once:
 *
 *   scan   : per row of op(A) / column of op(B): largest exponent and lowest set mantissa bit
 *   plan   : W = widest row span in bits  ->  S = ceil((W + 2) / 8) signed 8-bit slices, so that
 *            every element is represented EXACTLY as  x = 2^base(row) * sum_s d_s 256^(S-1-s),
 *            d_s in [-128, 127]  (row-wise block fixed point, no truncation at all)
 *   slice  : write the S int8 digit planes, K-major, for A ([S_A][m][Kp]) and B ([S_B][n][Kp])
 *   mma    : for every diagonal d = s + t:  D_d = sum_{s+t=d} A_s * B_t^T  as ONE K-concatenated
 *            int8 x int8 -> int32 GEMM on tcgen05 (kind::i8), operands staged by TMA into 128-byte
 *            swizzled shared memory, accumulators in TMEM, warp-specialised persistent kernel.
 *            Exact: |d| <= 128 and Kc * min(S_A, S_B) * 2^14 < 2^31 (K is cut into chunks of Kc).
 *   fold   : I = sum_d D_d 256^(ndiag-1-d) as a 448-bit two's complement integer per C element
 *            (accumulated across K chunks in a workspace), one correctly rounded conversion to
 *            binary128, then the reference epilogue  C = fma(alpha, s, mul(beta, C))
 *            (level3.hpp:102-109).
 *
 * Accuracy, exact setting (qb_set_tensor_keep(0)): the inner product is exact, so |c^ - c| <= u |A B|_ij
 * (+ the two epilogue roundings), far inside the fast-mode contract  gamma_k (|A||B|)_ij.
 *
 * Bounded setting (default, qb_set_tensor_keep(16)): only the `keep` most significant diagonals
 * d = 0 .. keep-1 are multiplied (136 digit-plane products instead of 18 x 18 = 324 for full 113-bit
 * mantissas).  With J = sum_{d<keep} D_d 256^(keep-1-d), the dropped diagonals amount to less than
 * min(S_A,S_B) * k * 64.3 units of J, so whenever |J| >= 2^125 the truncation is below
 * (k-1) u |c_ij| and the rounded J meets  |c^ - c| <= k u |c| <= gamma_k (|A||B|)_ij  (k >= 2,
 * S <= 24; derivation in DESIGN.md §4.1).  The fold checks that per element; an element that fails
 * it (cancellation: |c_ij| more than ~2^10 below the typical size; 0.02 % of the entries of a random
 * 8192^3 product with 16 diagonals, 1e-6 with 17) is NOT written by the fold
 * and is recomputed by k_oz_fixup, one warp per element, in the unrounded window accumulator of
 * qwide.cuh (error < k 2^-133 max|a b|, also inside the contract).  If more than 1/64 of a row pass
 * fails (structured cancellation), the pass is redone with all diagonals for the failed elements.  Inputs this cannot represent (Inf/NaN,
 * or a row whose exponent span needs more than QB_OZ_MAX_SLICES digits) make the planner decline
 * and the caller runs the integer-limb kernel (qb_level3.cu) instead; that is a different CUDA
 * kernel of this library, not a CPU fallback.
 */
#include "qb_internal.h"
#include "qb_tc.cuh"
#include "q128_chain.cuh"
#include "qwide.cuh"
#include "qb_crt.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

namespace qb {

/* ------------------------------------------------------------------ geometry */
static constexpr int OZ_BM = 128;            /* C tile rows   = UMMA M */
static constexpr int OZ_BN = 256;            /* C tile cols   = UMMA N */
static constexpr int OZ_BK = 128;            /* bytes (= int8 elements) of K per pipeline stage = one 128 B swizzle row */
static constexpr int OZ_UK = 32;             /* K per tcgen05.mma kind::i8 */
static constexpr int OZ_STAGES = 4;
static constexpr int OZ_A_BYTES = OZ_BM * OZ_BK; /* 16 KiB */
static constexpr int OZ_B_BYTES = OZ_BN * OZ_BK; /* 32 KiB */
static constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;
static constexpr int OZ_SMEM = OZ_STAGES * OZ_STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
static constexpr int OZ_THREADS = 192;       /* warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue */
static constexpr int OZ_MAX_S = QB_OZ_MAX_SLICES;
static constexpr int OZ_MAX_DIAG = 2 * OZ_MAX_S - 1;
static constexpr int OZ_NL = 14;             /* 32-bit limbs of the exact integer (448 bits) */
static constexpr int OZ_SCAN_CH = 256;       /* k per scan work item: 8192^2 gives 262k warps / threads, enough loads in flight for HBM */

struct OzMmaArgs {
  int32_t *D;                 /* [ndiag][Mp][Np] int32 */
  int64_t Mp, Np;             /* padded to tile multiples */
  int SA, SB, ndiag;
  int m_tiles, n_tiles;
  int kb_begin, nkb;          /* k-blocks of OZ_BK for this K chunk */
  uint8_t order[OZ_MAX_DIAG + 1]; /* diagonals, heaviest first */
  uint8_t *R;                 /* residue scheme: [N][Mp][Np] bytes, R_i = (A_i B_i^T) mod p_i in [0, p_i); ndiag = N */
};

/* per-modulus constants of the residue scheme (qb_crt.cuh); uploaded once per device by crt_upload_tables() */
__constant__ crt::Tables c_crt;

/* Tile order inside one diagonal: bands of OZ_GM m-tiles, n-tile outer / m-tile inner inside a band, so
 * that the ~148 tiles in flight form a compact 16 x 9 block of C: every A row panel is shared by ~9
 * CTAs and every B column panel by 16 through L2 (profiles/r1c: the n-fastest order re-read the
 * planes from HBM 3.4x more often than this needs at 8192^3). */
static constexpr int OZ_GM = 16;
__device__ __forceinline__ void oz_tile_decode(int rem, int m_tiles, int n_tiles, int &mt, int &nt)
{
  const int band_sz = OZ_GM * n_tiles;
  const int band = rem / band_sz, r = rem % band_sz;
  const int rows = min(OZ_GM, m_tiles - band * OZ_GM);
  nt = r / rows;
  mt = band * OZ_GM + r % rows;
}

/* ------------------------------------------------------------------ the tensor-core kernel */
/* CRT = 0: digit diagonals, D_d = sum_{s+t=d} A_s B_t^T as int32.  CRT = 1: residue planes, one product A_i B_i^T per
 * modulus, reduced mod p_i in the epilogue and stored as bytes. */
template <int CRT>
__global__ void __launch_bounds__(OZ_THREADS, 1)
k_oz_mma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const OzMmaArgs g)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + OZ_STAGES * OZ_STAGE_BYTES);
  uint64_t *full = bars;                       /* [OZ_STAGES] TMA -> MMA */
  uint64_t *empty = bars + OZ_STAGES;          /* [OZ_STAGES] MMA -> TMA */
  uint64_t *tfull = bars + 2 * OZ_STAGES;      /* [2] MMA -> epilogue */
  uint64_t *tempty = bars + 2 * OZ_STAGES + 2; /* [2] epilogue -> MMA */
  uint32_t *tmem_slot = (uint32_t *)(bars + 2 * OZ_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    for (int s = 0; s < OZ_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512); /* 2 accumulators x 256 columns */
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_diag = g.m_tiles * g.n_tiles;
  const int total = g.ndiag * tiles_per_diag;

  if (warp == 0) {
    /* ===================== TMA producer (one thread) ===================== */
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int d = CRT ? tile / tiles_per_diag : g.order[tile / tiles_per_diag];
        int mt, nt;
        oz_tile_decode(tile % tiles_per_diag, g.m_tiles, g.n_tiles, mt, nt);
        const int s_lo = CRT ? d : (d - (g.SB - 1) > 0 ? d - (g.SB - 1) : 0);
        const int s_hi = CRT ? d : (d < g.SA - 1 ? d : g.SA - 1);
        for (int s = s_lo; s <= s_hi; ++s) {
          const int t = CRT ? s : d - s;
          for (int kb = 0; kb < g.nkb; ++kb) {
            tc::mbar_wait(&empty[stage], phase ^ 1);
            uint8_t *sa = smem + stage * OZ_STAGE_BYTES;
            uint8_t *sb = sa + OZ_A_BYTES;
            tc::mbar_expect_tx(&full[stage], OZ_STAGE_BYTES);
            const int kc = (g.kb_begin + kb) * OZ_BK;
            tc::tma_load_3d(sa, &tmA, &full[stage], kc, mt * OZ_BM, s);
            tc::tma_load_3d(sb, &tmB, &full[stage], kc, nt * OZ_BN, t);
            if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    /* ===================== MMA issuer (one thread) ===================== */
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_i8(OZ_BM, OZ_BN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int d = CRT ? tile / tiles_per_diag : g.order[tile / tiles_per_diag];
        const int s_lo = CRT ? d : (d - (g.SB - 1) > 0 ? d - (g.SB - 1) : 0);
        const int s_hi = CRT ? d : (d < g.SA - 1 ? d : g.SA - 1);
        const int iters = (s_hi - s_lo + 1) * g.nkb;
        tc::mbar_wait(&tempty[acc], acc_phase ^ 1); /* epilogue has drained this accumulator */
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * OZ_BN;
        for (int it = 0; it < iters; ++it) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * OZ_STAGE_BYTES);
          const uint64_t adesc = tc::make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = tc::make_kmajor_sw128_desc(sa + OZ_A_BYTES);
#pragma unroll
          for (int k = 0; k < OZ_BK / OZ_UK; ++k)
            tc::mma_i8_ss(tmem_d, adesc + (uint64_t)(k * OZ_UK >> 4), bdesc + (uint64_t)(k * OZ_UK >> 4), idesc,
                          (it | k) != 0 ? 1u : 0u);
          tc::tc_commit(&empty[stage]); /* frees the smem slot when these MMAs have read it */
          if (++stage == OZ_STAGES) { stage = 0; phase ^= 1; }
        }
        tc::tc_commit(&tfull[acc]);     /* accumulator complete */
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    /* ===================== epilogue: TMEM -> registers -> D (int32) ===================== */
    const int quarter = warp & 3;       /* TMEM lanes 32*quarter .. +31 are the ones this warp may read */
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int d = CRT ? tile / tiles_per_diag : g.order[tile / tiles_per_diag];
      int mt, nt;
      oz_tile_decode(tile % tiles_per_diag, g.m_tiles, g.n_tiles, mt, nt);
      tc::mbar_wait(&tfull[acc], acc_phase);
      tc::tc_fence_after();
      const int64_t row = (int64_t)mt * OZ_BM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * OZ_BN;
      if (CRT) {
        /* accumulator mod p_d (crt::acc_mod), 32 residues = 32 bytes per thread and 32-column step */
        uint8_t *dst = g.R + ((int64_t)d * g.Mp + row) * g.Np + (int64_t)nt * OZ_BN;
        const uint32_t p = c_crt.p[d], off = c_crt.off[d], finv = c_crt.finv[d];
#pragma unroll 1
        for (int c = 0; c < OZ_BN / 32; ++c) {
          uint32_t v[32];
          tc::tmem_ld_32x32(taddr + c * 32, v);
          tc::tmem_ld_wait();
          uint32_t w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t u = v[4 * q + b] + off;
              uint32_t r = u - __umulhi(u, finv) * p;
              r = r >= p ? r - p : r;
              word |= r << (8 * b);
            }
            w[q] = word;
          }
          *reinterpret_cast<uint4 *>(dst + c * 32) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4 *>(dst + c * 32 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      } else {
      int32_t *dst = g.D + ((int64_t)d * g.Mp + row) * g.Np + (int64_t)nt * OZ_BN;
#pragma unroll 1
      for (int c = 0; c < OZ_BN / 32; ++c) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4 *>(dst + c * 32 + q * 4) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

/* ------------------------------------------------------------------ scan: row spans */
/* element fields: value = (-1)^s * M * 2^(ee - 16495), M < 2^113 */
struct OzElem { uint64_t lo, hi; int ee; uint32_t sign; int special; };
__device__ __forceinline__ OzElem oz_unpack(q128 a)
{
  OzElem o;
  const uint32_t ef = (uint32_t)(a.hi >> 48) & 0x7fffu;
  o.sign = (uint32_t)(a.hi >> 63);
  o.special = ef == 0x7fffu;
  o.lo = a.lo;
  o.hi = (a.hi & Q_MANT_HI_MASK) | (ef ? Q_IMPLICIT : 0);
  o.ee = ef ? (int)ef : 1;
  return o;
}
__device__ __forceinline__ int oz_tz(uint64_t lo, uint64_t hi) { return lo ? __ffsll((long long)lo) - 1 : 64 + __ffsll((long long)hi) - 1; }

/* rows x K view: X[r * sr + k * sk].  emax[r] = max ee over non-zero elements (0 if none),
 * lmin[r] = min (ee + tz(M)); flags |= 1 on Inf/NaN.  One warp per (row, OZ_SCAN_CH-wide k chunk) when k
 * is the contiguous direction, otherwise one thread per (row, chunk) with lanes along rows. */
__global__ void k_oz_scan(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, int *emax, int *lmin, int *flags)
{
  constexpr int CH = OZ_SCAN_CH;
  const int64_t nchunk = (K + CH - 1) / CH;
  int em = 0, lm = 0x7fffffff, sp = 0;
  int64_t r;
  if (sk == 1 || sr != 1) { /* warp per (row, chunk), lanes along k */
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= rows * nchunk) return;
    r = w / nchunk;
    const int64_t k0 = (w % nchunk) * CH, k1 = k0 + CH < K ? k0 + CH : K;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
      const OzElem e = oz_unpack(X[r * sr + k * sk]);
      sp |= e.special;
      if (e.lo | e.hi) { em = max(em, e.ee); lm = min(lm, e.ee + oz_tz(e.lo, e.hi)); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      em = max(em, __shfl_xor_sync(0xffffffffu, em, o));
      lm = min(lm, __shfl_xor_sync(0xffffffffu, lm, o));
      sp |= __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if (lane != 0) return;
  } else { /* rows contiguous: thread per (row, chunk), lanes along rows */
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * nchunk) return;
    r = t % rows;
    const int64_t k0 = (t / rows) * CH, k1 = k0 + CH < K ? k0 + CH : K;
    for (int64_t k = k0; k < k1; ++k) {
      const OzElem e = oz_unpack(X[r * sr + k * sk]);
      sp |= e.special;
      if (e.lo | e.hi) { em = max(em, e.ee); lm = min(lm, e.ee + oz_tz(e.lo, e.hi)); }
    }
  }
  if (em) { atomicMax(&emax[r], em); atomicMin(&lmin[r], lm); }
  if (sp) atomicOr(flags, 1);
}

/* out[0] = widest span of A rows, out[1] = widest span of B columns, out[2] = flags */
__global__ void k_oz_plan(const int *emaxA, const int *lminA, int64_t m, const int *emaxB, const int *lminB, int64_t n, const int *flags, int *out)
{
  __shared__ int sw[2];
  if (threadIdx.x == 0) { sw[0] = 0; sw[1] = 0; }
  __syncthreads();
  int wa = 0, wb = 0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) if (emaxA[i]) wa = max(wa, emaxA[i] + 113 - lminA[i]);
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) if (emaxB[j]) wb = max(wb, emaxB[j] + 113 - lminB[j]);
  atomicMax(&sw[0], wa);
  atomicMax(&sw[1], wb);
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = sw[0]; out[1] = sw[1]; out[2] = *flags; }
}

/* ------------------------------------------------------------------ slice: exact signed 8-bit digits */
/* thread = (row r, 4 consecutive k); plane s (0 = most significant digit) is [rows][Kp] int8.
 * x = X_int * 2^(base - 16495), base = emax[r] + 115 - 8 S, |X_int| < 2^(8S-2); digits are the
 * balanced base-256 expansion of X_int (so the sign needs no separate plane). */
/* digits of the 4 elements (row r, k = 4 g4 .. 4 g4 + 3): word[j] = byte q is digit j (0 = least significant) of element q */
template <int MAXS>
__device__ __forceinline__ void oz_slice4(const q128 *__restrict__ X, int64_t r, int64_t g4, int64_t K, int64_t sr, int64_t sk, int base, int S,
                                          uint32_t (&word)[MAXS])
{
#pragma unroll
  for (int s = 0; s < MAXS; ++s) word[s] = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t k = g4 * 4 + q;
    if (k >= K) continue;
    const OzElem e = oz_unpack(X[r * sr + k * sk]);
    if (!(e.lo | e.hi) || e.special) continue;
    u256 v; v.w0 = e.lo; v.w1 = e.hi; v.w2 = 0; v.w3 = 0;
    const int sh = e.ee - base;
    v = sh >= 0 ? u256_shl(v, (uint32_t)sh) : u256_shr_jam(v, (uint32_t)(-sh)); /* exact by construction of S */
    const uint64_t w[4] = {v.w0, v.w1, v.w2, v.w3};
    /* balanced digits, least significant first; a negative number takes digits of its magnitude
     * in [-127, 128] and negates them, so both signs land in [-128, 127] */
    uint32_t carry = 0;
    const uint32_t thr = 128u + e.sign;
#pragma unroll
    for (int j = 0; j < MAXS; ++j) {
      if (j < S) {
        uint32_t t = (uint32_t)((w[j >> 3] >> (8 * (j & 7))) & 0xffu) + carry;
        carry = t >= thr ? 1u : 0u;                 /* digit = t - 256 */
        const uint32_t dig = e.sign ? (0u - t) : t; /* low 8 bits are the int8 digit either way */
        word[j] |= (dig & 0xffu) << (8 * q);
      }
    }
  }
}

/* k contiguous in memory (or fully strided): thread = (row, 4 consecutive k), lanes along k; every thread
 * stores one 32-bit word per plane and a warp's stores are 128 contiguous bytes of one plane row. */
template <int MAXS>
__global__ void k_oz_slice(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax, int S, int64_t Kp,
                           int8_t *__restrict__ planes)
{
  const int64_t groups = Kp >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= rows * groups) return;
  const int64_t r = tid / groups, g4 = tid % groups;
  uint32_t word[MAXS];
  oz_slice4<MAXS>(X, r, g4, K, sr, sk, emax[r] + 115 - 8 * S, S, word);
#pragma unroll
  for (int j = 0; j < MAXS; ++j)
    if (j < S) *reinterpret_cast<uint32_t *>(planes + ((int64_t)(S - 1 - j) * rows + r) * Kp + g4 * 4) = word[j];
}

/* rows contiguous in memory (B of a row-major product, A of a col-major one): a CTA of 32 x 8 threads takes a tile of
 * 32 rows x 32 k with lanes along the rows, so the 16-byte loads coalesce; the plane words are transposed through
 * shared memory and leave as 32-byte runs along K (the direct store would scatter 4-byte words Kp bytes apart). */
template <int MAXS>
__global__ void __launch_bounds__(256) k_oz_slice_t(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax,
                                                   int S, int64_t Kp, int8_t *__restrict__ planes)
{
  __shared__ uint32_t sm[MAXS][32][9];            /* [digit][row][k-group], padded against bank conflicts */
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * 32, g0 = (int64_t)blockIdx.y * 8;
  const int64_t r = r0 + lane, g4 = g0 + wg;
  if (r < rows) {
    uint32_t word[MAXS];
    oz_slice4<MAXS>(X, r, g4, K, sr, sk, emax[r] + 115 - 8 * S, S, word);
#pragma unroll
    for (int j = 0; j < MAXS; ++j)
      if (j < S) sm[j][lane][wg] = word[j];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < S * 64; q += 256) {
    const int j = q >> 6, row = (q & 63) >> 1, half = q & 1;
    if (r0 + row >= rows) continue;
    const uint32_t *src = &sm[j][row][half * 4];
    *reinterpret_cast<uint4 *>(planes + ((int64_t)(S - 1 - j) * rows + r0 + row) * Kp + (g0 + half * 4) * 4) = make_uint4(src[0], src[1], src[2], src[3]);
  }
}

/* ------------------------------------------------------------------ fold: diagonals -> binary128 */
struct OzFoldArgs {
  const int32_t *D; int64_t Mp, Np; int ndiag;
  int nsets; int64_t set_stride;     /* K chunks kept as separate int32 diagonal sets (summed here in 64 bits) */
  int64_t m, n, row0;                /* C rows [row0, row0 + m) of the full problem are this pass */
  const int *emaxA, *emaxB; int SA, SB;
  uint32_t *W; int w_in, w_out;      /* 448-bit running sums across K chunks: [OZ_NL][Mp*Np] */
  q128 alpha, beta; q128 *C; int64_t sci, scj;
  /* bounded setting: ndiag = kept diagonals, the integer is J (units of 256^exp8 above the full one) */
  int exp8;                          /* dropped diagonals (0 in the exact setting) */
  int check;                         /* 1: elements with |J| < 2^OZ_JMIN_BIT are flagged and left unwritten */
  int only_flagged;                  /* 1: (exact redo) write only the elements flagged earlier */
  uint8_t *flag;                     /* [Mp*Np] */
  int2 *list; int list_cap; int *counter;
};
static constexpr int OZ_JMIN_BIT = 125;

/* NSETS > 0: that many diagonal sets, unrolled; NSETS = 0: g.nsets of them (any count) */
template <int NSETS>
__global__ void __launch_bounds__(256) k_oz_fold(const OzFoldArgs g)
{
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = idx / g.n, j = idx % g.n;
  if (i >= g.m) return;
  const int64_t plane = g.Mp * g.Np, off = i * g.Np + j;
  uint32_t L[OZ_NL];
  long long carry = 0;
#pragma unroll
  for (int l = 0; l < OZ_NL; ++l) {
    long long acc = carry;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int d = g.ndiag - 1 - (4 * l + b);
      if (d >= 0) {
        const int32_t *dp = g.D + (int64_t)d * plane + off;
        long long v = dp[0];
        if (NSETS > 0) {
#pragma unroll
          for (int c = 1; c < NSETS; ++c) v += dp[c * g.set_stride];
        } else {
          for (int c = 1; c < g.nsets; ++c) v += dp[c * g.set_stride];
        }
        acc += v << (8 * b);
      }
    }
    L[l] = (uint32_t)acc;
    carry = acc >> 32;
  }
  if (g.w_in) {
    uint32_t c = 0;
#pragma unroll
    for (int l = 0; l < OZ_NL; ++l) {
      const uint64_t t = (uint64_t)L[l] + g.W[(int64_t)l * plane + off] + c;
      L[l] = (uint32_t)t; c = (uint32_t)(t >> 32);
    }
  }
  if (g.w_out) {
#pragma unroll
    for (int l = 0; l < OZ_NL; ++l) g.W[(int64_t)l * plane + off] = L[l];
    return;
  }
  if (g.only_flagged && !g.flag[off]) return;
  /* ---- sign / magnitude ---- */
  const uint32_t neg = L[OZ_NL - 1] >> 31;
  if (neg) {
    uint32_t c = 1;
#pragma unroll
    for (int l = 0; l < OZ_NL; ++l) { const uint64_t t = (uint64_t)(~L[l]) + c; L[l] = (uint32_t)t; c = (uint32_t)(t >> 32); }
  }
  int top = -1;
  uint32_t topv = 0;
#pragma unroll
  for (int l = 0; l < OZ_NL; ++l) if (L[l]) { top = l; topv = L[l]; }
  if (g.check) { /* |J| >= 2^125 or the element goes to the fix-up (header comment) */
    const int msb = top < 0 ? -1 : 32 * top + 31 - __clz((int)topv);
    const bool weak = msb < OZ_JMIN_BIT;
    g.flag[off] = weak ? 1 : 0;
    if (weak) {
      const int slot = atomicAdd(g.counter, 1);
      if (slot < g.list_cap) g.list[slot] = make_int2((int)(g.row0 + i), (int)j);
      return;
    }
  }
  q128 sum;
  if (top < 0) {
    sum = q_zero(0); /* an exact zero sum is +0 (a +0-seeded chain never yields -0) */
  } else {
    /* 9-limb window ending at the top limb, through a local buffer (dynamic index) */
    uint32_t buf[OZ_NL + 8];
#pragma unroll
    for (int l = 0; l < 8; ++l) buf[l] = 0;
#pragma unroll
    for (int l = 0; l < OZ_NL; ++l) buf[8 + l] = L[l];
    uint32_t w[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) w[k] = buf[top + k];
    uint32_t sticky = 0;
#pragma unroll
    for (int l = 0; l < OZ_NL; ++l) if (l < top - 8) sticky |= L[l];
    const int lz = __clz((int)w[8]);
    uint32_t R[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) R[k] = __funnelshift_l(w[k], w[k + 1], lz);
    sticky |= w[0] << lz; /* lz == 0: w[0] is entirely below the window */
    if (lz == 0) sticky |= w[0];
    u256 Rq;
    Rq.w0 = ((uint64_t)R[1] << 32) | R[0] | (sticky != 0);
    Rq.w1 = ((uint64_t)R[3] << 32) | R[2];
    Rq.w2 = ((uint64_t)R[5] << 32) | R[4];
    Rq.w3 = ((uint64_t)R[7] << 32) | R[6];
    /* I * 2^Eb, Eb = baseA + baseB - 2 * 16495; MSB of I at bit p = 32 top + 31 - lz  ->  er = p + Eb + QBIAS */
    const int baseA = g.emaxA[g.row0 + i] + 115 - 8 * g.SA, baseB = g.emaxB[j] + 115 - 8 * g.SB;
    const int p = 32 * top + 31 - lz;
    const int er = p + 8 * g.exp8 + baseA + baseB - 2 * 16495 + QBIAS;
    sum = q_round_pack(neg, er, Rq);
  }
  q128 *c = g.C + i * g.sci + j * g.scj;
  *c = q_fma(g.alpha, sum, q_mul(g.beta, *c)); /* level3.hpp:102-109: beta*C is always evaluated */
}

/* ------------------------------------------------------------------ fix-up of the flagged elements (bounded setting) */
struct OzFixArgs {
  const int2 *list; const int *counter; int list_cap;
  const q128 *A; int64_t sai, sal; const q128 *B; int64_t sbl, sbj; int64_t k;
  q128 alpha, beta; q128 *C; int64_t sci, scj;
};
__device__ __noinline__ qwide oz_merge(qwide a, qwide b) { qw_merge(a, b); return a; }
/* one warp per flagged C element: the k products go into the unrounded window accumulator (lanes
 * stride over k), a shuffle tree merges the 32 windows, lane 0 rounds once and applies the epilogue */
__global__ void __launch_bounds__(128) k_oz_fixup(const OzFixArgs g)
{
  __shared__ uint32_t scr[QWA_COL_WORDS * 128];
  const int lane = threadIdx.x & 31;
  uint32_t *col = scr + threadIdx.x;
  qwa_col_init(col, 128);
  int count = *g.counter;
  if (count > g.list_cap) count = g.list_cap;
  const int nwarps = gridDim.x * 4;
  for (int e = blockIdx.x * 4 + (threadIdx.x >> 5); e < count; e += nwarps) {
    const int2 ij = g.list[e];
    const q128 *ap = g.A + (int64_t)ij.x * g.sai, *bp = g.B + (int64_t)ij.y * g.sbj;
    qwacc acc = qwa_zero();
    uint32_t bad = 0;
    for (int64_t l = lane; l < g.k; l += 32) {
      const q128 a = ap[l * g.sal], b = bp[l * g.sbl];
      if (qwa_fma(acc, qop_load_n(a), qop_load_n(b), col, 128)) qwa_fma_rare(acc, a, b, bad);
    }
    qwide v = qwa_fold(acc);
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      qwide t;
      t.w0 = __shfl_down_sync(0xffffffffu, v.w0, o); t.w1 = __shfl_down_sync(0xffffffffu, v.w1, o);
      t.w2 = __shfl_down_sync(0xffffffffu, v.w2, o); t.w3 = __shfl_down_sync(0xffffffffu, v.w3, o);
      t.w4 = __shfl_down_sync(0xffffffffu, v.w4, o); t.w5 = __shfl_down_sync(0xffffffffu, v.w5, o);
      t.E = __shfl_down_sync(0xffffffffu, v.E, o);
      v = oz_merge(v, t);
    }
    if (lane == 0) {
      q128 *c = g.C + (int64_t)ij.x * g.sci + (int64_t)ij.y * g.scj;
      *c = q_fma(g.alpha, qw_finish(v, bad), q_mul(g.beta, *c));
    }
  }
}


/* ================================================================== residue scheme (qb_crt.cuh) */
/* The 4 elements (row r, k = 4 g4 .. 4 g4 + 3) as words of |X| and signs, X = x / 2^(base - 16495) an exact integer below 2^W. */
struct Crt4 { uint32_t w[4][crt::NWMAX]; uint32_t sign[4]; };
template <int NW>
__device__ __forceinline__ void crt_load4(const q128 *__restrict__ X, int64_t r, int64_t g4, int64_t K, int64_t sr, int64_t sk, int base, Crt4 &c)
{
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int j = 0; j < crt::NWMAX; ++j) c.w[q][j] = 0;
    c.sign[q] = 0;
    const int64_t k = g4 * 4 + q;
    if (k >= K) continue;
    const OzElem e = oz_unpack(X[r * sr + k * sk]);
    if (!(e.lo | e.hi) || e.special) continue;
    u256 v; v.w0 = e.lo; v.w1 = e.hi; v.w2 = 0; v.w3 = 0;
    const int sh = e.ee - base;
    v = sh >= 0 ? u256_shl(v, (uint32_t)sh) : u256_shr_jam(v, (uint32_t)(-sh)); /* exact: base <= lowest set bit of the row */
    c.w[q][0] = (uint32_t)v.w0; c.w[q][1] = (uint32_t)(v.w0 >> 32);
    c.w[q][2] = (uint32_t)v.w1; c.w[q][3] = (uint32_t)(v.w1 >> 32);
    c.w[q][4] = (uint32_t)v.w2; c.w[q][5] = (uint32_t)(v.w2 >> 32);
    c.sign[q] = e.sign;
    if (e.sign) crt::negate_words<NW>(c.w[q]);   /* non-zero here: 2^(32 NW) - |X|, the seed of residue_sym carries the rest */
  }
}
/* residues of the 4 elements modulo p_i packed as 4 int8 (byte q = element q) */
template <int NW>
__device__ __forceinline__ uint32_t crt_word(const Crt4 &c, int i)
{
  /* the elements arrive with negative ones already in two's complement (crt_load4<NW>) */
  const uint32_t r0 = crt::residue_sym<NW>(c.w[0], c.sign[0], i, c_crt), r1 = crt::residue_sym<NW>(c.w[1], c.sign[1], i, c_crt);
  const uint32_t r2 = crt::residue_sym<NW>(c.w[2], c.sign[2], i, c_crt), r3 = crt::residue_sym<NW>(c.w[3], c.sign[3], i, c_crt);
  return __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);   /* low bytes of r0..r3 */
}

/* k contiguous (or fully strided): thread = (row, 4 consecutive k); plane i is [rows][Kp] int8 */
template <int NW>
__global__ void __launch_bounds__(256) k_crt_residues(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax,
                                                      int W, int N, int64_t Kp, int8_t *__restrict__ planes)
{
  const int64_t groups = Kp >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= rows * groups) return;
  const int64_t r = tid / groups, g4 = tid % groups;
  Crt4 c;
  crt_load4<NW>(X, r, g4, K, sr, sk, emax[r] + 113 - W, c);
#pragma unroll
  for (int g = 0; g < crt::NGMAX; ++g) {
    if (4 * g < N) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = 4 * g + b;
        if (i < N) *reinterpret_cast<uint32_t *>(planes + ((int64_t)i * rows + r) * Kp + g4 * 4) = crt_word<NW>(c, i);
      }
    }
  }
}

/* rows contiguous in memory: CTA = 32 rows x 32 k with lanes along the rows (coalesced 16-byte loads); the plane words are
 * transposed through shared memory and leave as 32-byte runs along K (cf. k_oz_slice_t) */
static constexpr int CRT_T_SMEM = crt::NMP * 32 * 9 * 4;
template <int NW>
__global__ void __launch_bounds__(256) k_crt_residues_t(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax,
                                                        int W, int N, int64_t Kp, int8_t *__restrict__ planes)
{
  extern __shared__ uint32_t crt_sm[];            /* [plane][row][k-group], padded: 9 words per row */
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const int64_t r0 = (int64_t)blockIdx.x * 32, g0 = (int64_t)blockIdx.y * 8;
  const int64_t r = r0 + lane, g4 = g0 + wg;
  if (r < rows) {
    Crt4 c;
    crt_load4<NW>(X, r, g4, K, sr, sk, emax[r] + 113 - W, c);
#pragma unroll
    for (int g = 0; g < crt::NGMAX; ++g) {
      if (4 * g < N) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = 4 * g + b;
          if (i < N) crt_sm[(i * 32 + lane) * 9 + wg] = crt_word<NW>(c, i);
        }
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < N * 64; q += 256) {
    const int i = q >> 6, row = (q & 63) >> 1, half = q & 1;
    if (r0 + row >= rows) continue;
    const uint32_t *src = &crt_sm[(i * 32 + row) * 9 + half * 4];
    *reinterpret_cast<uint4 *>(planes + ((int64_t)i * rows + r0 + row) * Kp + (g0 + half * 4) * 4) = make_uint4(src[0], src[1], src[2], src[3]);
  }
}

struct CrtFoldArgs {
  const uint8_t *R; int64_t Mp, Np;
  int64_t m, n, row0;                /* C rows [row0, row0 + m) of the full problem are this pass */
  const int *emaxA, *emaxB; int WA, WB;
  q128 alpha, beta; q128 *C; int64_t sci, scj;
  int simple;                        /* alpha == 1 and beta == +-0 */
  int64_t n4p;                       /* threads per row: ceil(n / 4) rounded up to 32 */
  int npeer; q128 *peer[QB_MAX_PEERS]; /* fused gather: the same C block inside each peer GPU's buffer (NVLink peer stores) */
};
/* thread = 4 consecutive columns of one C row: one 32-bit load per residue plane, then per element the reconstruction
 * (crt::reconstruct_dev), ONE rounding to binary128 and the reference epilogue C = fma(alpha, s, mul(beta, C)) (level3.hpp:102-109).
 * Threads per row are padded to a multiple of 32 (n4p), so a warp never straddles two rows.
 * Fused gather (npeer > 0): the finished elements also go into every peer GPU's copy of C.  NVLink wants whole 128-byte lines, so
 * the 4 x 32 elements of a warp are staged in shared memory (2 KB per warp, dynamic: none without peers) and leave as four
 * 512-byte contiguous warp stores per peer instead of 16-byte pieces 64 bytes apart. */
template <int NG>
__global__ void __launch_bounds__(128) k_crt_fold(const CrtFoldArgs g, const __grid_constant__ crt::Plan pl)
{
  extern __shared__ uint4 fold_sm[];            /* [4 warps][128 elements] when npeer > 0 and C is row-major */
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = idx / g.n4p, jg = idx % g.n4p, j0 = jg * 4;
  const bool valid = i < g.m && j0 < g.n;
  const bool stage = g.npeer > 0 && g.scj == 1;
  if (!valid && !stage) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (valid) {
    const int64_t plane = g.Mp * g.Np, off = i * g.Np + j0;
    uint32_t rw[4 * NG];
#pragma unroll
    for (int c = 0; c < 4 * NG; ++c) rw[c] = c < pl.N ? *reinterpret_cast<const uint32_t *>(g.R + (int64_t)c * plane + off) : 0u;
    const int baseA = g.emaxA[g.row0 + i] + 113 - g.WA;
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
      const int64_t j = j0 + e;
      if (j >= g.n) break;
      uint32_t r[crt::NMP];
#pragma unroll
      for (int c = 0; c < 4 * NG; ++c) r[c] = (rw[c] >> (8 * e)) & 0xffu;
      uint32_t Y[NG + 1], neg;
      crt::reconstruct_dev<NG>(r, pl, Y, neg);
      const int baseB = g.emaxB[j] + 113 - g.WB;
      const q128 sum = crt::limbs_to_q<NG + 1>(Y, neg, baseA + baseB - 2 * 16495);
      q128 *c = g.C + i * g.sci + j * g.scj;
      const q128 cin = *c;
      /* alpha = 1, beta = +-0, finite C: mul(beta, C) = +-0 and fma(1, s, +-0) = s (s is never -0) - the same bits as the
       * general line below, without the two software roundings */
      q128 out;
      if (g.simple && ((cin.hi >> 48) & 0x7fffu) != 0x7fffu) out = sum;
      else out = q_fma(g.alpha, sum, q_mul(g.beta, cin)); /* beta*C is always evaluated */
      *c = out;
      if (stage) fold_sm[warp * 128 + 4 * lane + e] = make_uint4((uint32_t)out.lo, (uint32_t)(out.lo >> 32), (uint32_t)out.hi, (uint32_t)(out.hi >> 32));
      else
        for (int q = 0; q < g.npeer; ++q) g.peer[q][i * g.sci + j * g.scj] = out;   /* col-major C: element-wise peer stores */
    }
  }
  if (stage) {
    __syncwarp();
    if (i < g.m) {
      const int64_t jw = (jg - lane) * 4;       /* first column of the warp's 128 */
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int64_t col = jw + 32 * s4 + lane;
        if (col < g.n) {
          const uint4 v = fold_sm[warp * 128 + 32 * s4 + lane];
          for (int q = 0; q < g.npeer; ++q) *reinterpret_cast<uint4 *>(g.peer[q] + i * g.sci + col) = v;
        }
      }
    }
  }
}
template <int NG>
static void launch_crt_fold_ng(const CrtFoldArgs &f, const crt::Plan &pl, cudaStream_t st)
{
  const int64_t threads = f.m * f.n4p;
  const size_t smem = (f.npeer > 0 && f.scj == 1) ? 4 * 128 * sizeof(uint4) : 0;
  k_crt_fold<NG><<<(unsigned)((threads + 127) / 128), 128, smem, st>>>(f, pl);
}
static void launch_crt_fold(const CrtFoldArgs &f, const crt::Plan &pl, cudaStream_t st)
{
  switch (pl.NG) {
#define QB_CRT_CASE(G) case G: launch_crt_fold_ng<G>(f, pl, st); break;
    QB_CRT_CASE(1) QB_CRT_CASE(2) QB_CRT_CASE(3) QB_CRT_CASE(4) QB_CRT_CASE(5) QB_CRT_CASE(6) QB_CRT_CASE(7)
    QB_CRT_CASE(8) QB_CRT_CASE(9) QB_CRT_CASE(10) QB_CRT_CASE(11) QB_CRT_CASE(12) QB_CRT_CASE(13)
#undef QB_CRT_CASE
  }
  count_launch();
}

/* ------------------------------------------------------------------ host side */
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode()
{
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

/* planes [S][rows][Kp] int8 -> 3-D map {Kp, rows, S}, box {128, box_rows, 1}, 128-byte swizzle; rows
 * beyond `rows` are zero-filled by the TMA unit */
static bool make_plane_map(CUtensorMap *tm, const int8_t *planes, int S, int64_t rows, int64_t Kp, int box_rows)
{
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)Kp, (cuuint64_t)Kp * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_sm_count = 0;
static int sm_count()
{
  if (!g_sm_count) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

/* D[d] = sum_{s+t=d} A_s B_t^T over k-blocks [kb_begin, kb_begin + nkb) */
cudaError_t launch_oz_mma(const int8_t *pA, const int8_t *pB, int SA, int SB, int64_t m, int64_t n, int64_t Kp, int kb_begin, int nkb,
                          int32_t *D, int64_t Mp, int64_t Np, cudaStream_t st, int keep)
{
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_oz_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_oz_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (!make_plane_map(&tmA, pA, SA, m, Kp, OZ_BM) || !make_plane_map(&tmB, pB, SB, n, Kp, OZ_BN)) return cudaErrorInvalidValue;
  OzMmaArgs g;
  g.D = D; g.R = nullptr; g.Mp = Mp; g.Np = Np; g.SA = SA; g.SB = SB; g.ndiag = SA + SB - 1;
  if (keep > 0 && keep < g.ndiag) g.ndiag = keep;   /* diagonals 0 .. keep-1 only */
  g.m_tiles = (int)(Mp / OZ_BM); g.n_tiles = (int)(Np / OZ_BN);
  g.kb_begin = kb_begin; g.nkb = nkb;
  /* heaviest diagonals first so the static round-robin over CTAs stays balanced */
  int idx[OZ_MAX_DIAG + 1];
  for (int d = 0; d < g.ndiag; ++d) idx[d] = d;
  auto pairs = [&](int d) { return std::min(d, SA - 1) - std::max(0, d - (SB - 1)) + 1; };
  std::stable_sort(idx, idx + g.ndiag, [&](int a, int b) { return pairs(a) > pairs(b); });
  for (int d = 0; d < g.ndiag; ++d) g.order[d] = (uint8_t)idx[d];
  const int total = g.ndiag * g.m_tiles * g.n_tiles;
  const int grid = std::min(total, sm_count());
  k_oz_mma<0><<<grid, OZ_THREADS, OZ_SMEM, st>>>(tmA, tmB, g);
  count_launch();
  return cudaGetLastError();
}

/* residue scheme: R_i = (A_i B_i^T) mod p_i for the N residue planes, all k-blocks [0, nkb) in one accumulation
 * (|acc| <= 128*128*Kp <= 2^30: Kp <= 65536) */
static cudaError_t launch_crt_mma(const int8_t *pA, const int8_t *pB, int N, int64_t m, int64_t n, int64_t Kp, uint8_t *R, int64_t Mp, int64_t Np,
                                  cudaStream_t st)
{
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_oz_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (!make_plane_map(&tmA, pA, N, m, Kp, OZ_BM) || !make_plane_map(&tmB, pB, N, n, Kp, OZ_BN)) return cudaErrorInvalidValue;
  OzMmaArgs g;
  memset(&g, 0, sizeof(g));
  g.R = R; g.Mp = Mp; g.Np = Np; g.SA = N; g.SB = N; g.ndiag = N;
  g.m_tiles = (int)(Mp / OZ_BM); g.n_tiles = (int)(Np / OZ_BN);
  g.kb_begin = 0; g.nkb = (int)(Kp / OZ_BK);
  const int64_t total = (int64_t)N * g.m_tiles * g.n_tiles;
  const int grid = (int)std::min<int64_t>(total, sm_count());
  k_oz_mma<1><<<grid, OZ_THREADS, OZ_SMEM, st>>>(tmA, tmB, g);
  count_launch();
  return cudaGetLastError();
}

/* ---- workspaces (grow-only, per process; callers hold the library mutex) ----
 * meta: per-row exponent data + plan (small, survives a regrowth of the big buffer)
 * buf : digit planes, diagonals, wide accumulators */
struct OzWork {
  void *buf = nullptr; size_t bytes = 0;
  void *meta = nullptr; size_t meta_bytes = 0;
  int device = -1;
  int *h_plan = nullptr; /* pinned, 16 ints */
};
static OzWork g_oz;
static cudaError_t oz_grow(void **p, size_t *have, size_t want)
{
  if (want <= *have) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) return e;
  *have = want;
  return cudaSuccess;
}
static cudaError_t oz_reserve(size_t meta_bytes, size_t bytes)
{
  int dev = 0; cudaGetDevice(&dev);
  if (g_oz.device != dev) { /* buffers of another device are left to that context */
    g_oz.buf = nullptr; g_oz.bytes = 0; g_oz.meta = nullptr; g_oz.meta_bytes = 0; g_oz.device = dev;
  }
  if (!g_oz.h_plan) { cudaError_t e = cudaMallocHost((void **)&g_oz.h_plan, 64); if (e != cudaSuccess) return e; }
  cudaError_t e = oz_grow(&g_oz.meta, &g_oz.meta_bytes, meta_bytes);
  if (e != cudaSuccess) return e;
  return oz_grow(&g_oz.buf, &g_oz.bytes, bytes);
}
void oz_release()
{
  if (g_oz.buf) cudaFree(g_oz.buf);
  if (g_oz.meta) cudaFree(g_oz.meta);
  g_oz.buf = nullptr; g_oz.bytes = 0; g_oz.meta = nullptr; g_oz.meta_bytes = 0;
}

static inline int64_t rup(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static OzStats g_last_stats;
OzStats oz_last_stats() { return g_last_stats; }
static int g_oz_keep = 16;      /* leading diagonals multiplied in the bounded setting; 0 = all (exact) */
void oz_set_keep(int keep) { g_oz_keep = keep < 0 ? 0 : keep; }
int oz_get_keep() { return g_oz_keep; }

/* CUDA events around every k_oz_mma launch of the last qgemm, on the launching stream (bench.py's
 * roofline needs the kernel's own duration, not the whole call's) */
static constexpr int OZ_MAX_EV = 64;
static cudaEvent_t g_ev[2 * OZ_MAX_EV];
static int g_ev_made = 0, g_ev_used = 0;
static void oz_ev_record(int which, cudaStream_t st)
{
  if (g_ev_used >= OZ_MAX_EV) return;
  if (!g_ev_made) { for (int i = 0; i < 2 * OZ_MAX_EV; ++i) cudaEventCreate(&g_ev[i]); g_ev_made = 1; }
  cudaEventRecord(g_ev[2 * g_ev_used + which], st);
  if (which == 1) ++g_ev_used;
}
/* blocks until the last recorded launch has finished; returns the summed duration in ms */
double oz_last_mma_ms(int *launches)
{
  double tot = 0;
  for (int i = 0; i < g_ev_used; ++i) {
    float ms = 0;
    if (cudaEventSynchronize(g_ev[2 * i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, g_ev[2 * i], g_ev[2 * i + 1]) == cudaSuccess) tot += ms;
  }
  if (launches) *launches = g_ev_used;
  return tot;
}

static void launch_oz_slice(const q128 *X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *emax, int S, int64_t Kp, int8_t *planes, cudaStream_t st)
{
  if (sk == 1 || sr != 1) {
    const int64_t threads = rows * (Kp / 4);
    k_oz_slice<OZ_MAX_S><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, rows, K, sr, sk, emax, S, Kp, planes);
  } else {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)(Kp / 32));
    k_oz_slice_t<OZ_MAX_S><<<grid, 256, 0, st>>>(X, rows, K, sr, sk, emax, S, Kp, planes);
  }
  count_launch();
}

/* capacity of the fix-up list of one row pass: above 1/64 of the pass the exact redo is cheaper */
static inline int64_t oz_list_cap(int64_t mb, int64_t n) { return std::max<int64_t>(1024, std::min<int64_t>((mb * n) / 64, (int64_t)1 << 22)); }


/* ---- residue scheme, host side ---- */
static int g_oz_scheme = 1;     /* 1 = residues (qb_crt.cuh), 0 = digit diagonals */
void oz_set_scheme(int v) { g_oz_scheme = v ? 1 : 0; }
int oz_get_scheme() { return g_oz_scheme; }

static cudaError_t crt_upload_tables()
{
  static int uploaded_dev = -1;
  int dev = 0; cudaGetDevice(&dev);
  if (uploaded_dev == dev) return cudaSuccess;
  static crt::Tables T;
  crt::host::build_tables(T);
  cudaError_t e = cudaMemcpyToSymbol(c_crt, &T, sizeof(T));
  if (e != cudaSuccess) return e;
  e = cudaDeviceSynchronize();   /* once per device: the copy from pageable memory must have landed before any stream reads it */
  if (e != cudaSuccess) return e;
#define QB_CRT_ATTR(NW) e = cudaFuncSetAttribute(k_crt_residues_t<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, CRT_T_SMEM); if (e != cudaSuccess) return e;
  QB_CRT_ATTR(1) QB_CRT_ATTR(2) QB_CRT_ATTR(3) QB_CRT_ATTR(4) QB_CRT_ATTR(5) QB_CRT_ATTR(6)
#undef QB_CRT_ATTR
  uploaded_dev = dev;
  return cudaSuccess;
}
static const crt::Plan &crt_plan(int N)
{
  static crt::Plan plans[crt::NM + 1];
  static bool have[crt::NM + 1] = {false};
  if (!have[N]) { crt::host::build_plan(N, plans[N]); have[N] = true; }
  return plans[N];
}
static void launch_crt_residues(const q128 *X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *emax, int W, int N, int64_t Kp, int8_t *planes,
                                cudaStream_t st)
{
  const int nw = std::min(crt::NWMAX, std::max(1, (W + 31) / 32));
  const bool direct = (sk == 1 || sr != 1);
  const int64_t threads = rows * (Kp / 4);
  const unsigned g1 = (unsigned)((threads + 255) / 256);
  const dim3 g2((unsigned)((rows + 31) / 32), (unsigned)(Kp / 32));
  switch (nw) {
#define QB_CRT_CASE(NW)                                                                                     \
  case NW:                                                                                                  \
    if (direct) k_crt_residues<NW><<<g1, 256, 0, st>>>(X, rows, K, sr, sk, emax, W, N, Kp, planes);          \
    else k_crt_residues_t<NW><<<g2, 256, CRT_T_SMEM, st>>>(X, rows, K, sr, sk, emax, W, N, Kp, planes);      \
    break;
    QB_CRT_CASE(1) QB_CRT_CASE(2) QB_CRT_CASE(3) QB_CRT_CASE(4) QB_CRT_CASE(5) QB_CRT_CASE(6)
#undef QB_CRT_CASE
  }
  count_launch();
}

/* Rows of each pass of the residue scheme: the sizes sum to m, every size but the last is a multiple of OZ_BM, none exceeds `cap`.
 * shape 0 (default, the measured one): equal passes of `cap` rows.  shape 1 (experimental, qb_set_tensor_pass_shape): a short
 * first pass (its A residues cannot overlap a tensor pass) and a short last pass (neither can its fold, nor its peer stores),
 * equal passes in between. */
static int g_oz_pass_shape = 0;
void oz_set_pass_shape(int v) { g_oz_pass_shape = v == 1 ? 1 : 0; }
int oz_get_pass_shape() { return g_oz_pass_shape; }
std::vector<int64_t> oz_crt_pass_rows(int64_t m, int64_t cap, int shape)
{
  std::vector<int64_t> rows;
  if (m <= 0 || cap <= 0) return rows;
  const int64_t small = std::max<int64_t>(OZ_BM, rup(cap / 4, OZ_BM));
  const bool shaped = shape == 1 && cap >= 4 * OZ_BM && m >= 2 * small + cap;
  int64_t left = m, last = 0;
  if (shaped) {   /* the ragged remainder of m stays in the LAST pass, so every pass starts on a tile boundary */
    rows.push_back(small);
    last = small + (m - 2 * small) % OZ_BM;
    left = m - small - last;
  }
  while (left > 0) { const int64_t r = std::min(cap, left); rows.push_back(r); left -= r; }
  if (shaped) rows.push_back(last);
  return rows;
}

/* Internal streams of the residue scheme: the tensor kernel of row pass p (sM, highest priority) runs while the residues of
 * the A rows of pass p+1 (sA) and the reconstruction of pass p-1 (sF) use the integer pipes of the same SMs (the persistent
 * tensor kernel holds one 192-thread CTA per SM; the other two kernels need no shared memory and fit beside it). */
struct CrtStreams {
  cudaStream_t sM = nullptr, sA = nullptr, sF = nullptr;
  cudaEvent_t in = nullptr, resA[2] = {nullptr, nullptr}, mma[2] = {nullptr, nullptr}, fold[2] = {nullptr, nullptr};
  int device = -1;
};
static CrtStreams g_cs;
static cudaError_t crt_streams()
{
  int dev = 0; cudaGetDevice(&dev);
  if (g_cs.device == dev) return cudaSuccess;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaError_t e = cudaStreamCreateWithPriority(&g_cs.sM, cudaStreamNonBlocking, hi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&g_cs.sA, cudaStreamNonBlocking, lo);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&g_cs.sF, cudaStreamNonBlocking, lo);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_cs.in, cudaEventDisableTiming);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = cudaEventCreateWithFlags(&g_cs.resA[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_cs.mma[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_cs.fold[b], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) g_cs.device = dev;
  return e;
}

/* The residue-scheme GEMM after scan + plan (WA, WB = widest row / column span in bits, no Inf/NaN).  *used = 0: declined.
 * The C rows are produced in row passes, software-pipelined over three internal streams (CrtStreams); `st` waits for the
 * reconstruction of every pass before the pass callback runs and before this function returns, so for the caller all the
 * work is ordered on `st` as usual. */
static cudaError_t launch_gemm_crt(const GemmArgs &a, cudaStream_t st, int *used, size_t ws_budget, oz_pass_cb cb, void *cb_user, int min_passes,
                                   int WA, int WB, size_t meta_ints)
{
  const int64_t m = a.m, n = a.n, k = a.k;
  const int64_t Kp = rup(k, OZ_BK);
  WA = std::max(WA, 1); WB = std::max(WB, 1);
  if (WA > crt::WMAX || WB > crt::WMAX || Kp > 65536) return cudaSuccess;
  int lk = 0;
  while (((int64_t)1 << lk) < k) ++lk;
  const int need_bits = WA + WB + lk + 1;             /* P > 2 * k * 2^(WA + WB) >= 2 |I| */
  const int N = crt::host::moduli_for_bits(need_bits);
  if (N == 0) return cudaSuccess;
  cudaError_t e = crt_upload_tables();
  if (e != cudaSuccess) return e;
  e = crt_streams();
  if (e != cudaSuccess) return e;
  const crt::Plan &pl = crt_plan(N);
  /* ---- workspace: residue planes of B | 2 x (residue planes of the A rows of a pass | R of the pass) ---- */
  const int64_t Np = rup(n, OZ_BN);
  const size_t pb_b = rup((int64_t)N * n * Kp, 1024);
  auto pa_bytes = [&](int64_t mb) -> size_t { return (size_t)rup((int64_t)N * mb * Kp, 1024); };
  auto pass_bytes = [&](int64_t mb) -> size_t { return pa_bytes(mb) + (size_t)rup((int64_t)N * rup(mb, OZ_BM) * Np, 1024); };
  int want_passes = m >= 4096 ? 4 : (m >= 2048 ? 2 : 1);
  if (cb && min_passes > want_passes) want_passes = min_passes;
  int64_t mb = std::max<int64_t>(OZ_BM, rup((m + want_passes - 1) / want_passes, OZ_BM));
  auto bufs = [&](int64_t mb_) -> int { return mb_ < m ? 2 : 1; };
  while (mb > OZ_BM && pb_b + bufs(mb) * pass_bytes(mb) > ws_budget) mb = rup((mb + 1) / 2, OZ_BM);
  if (pb_b + bufs(mb) * pass_bytes(mb) > ws_budget) return cudaSuccess;
  const int nb = bufs(mb);
  e = oz_reserve(meta_ints * 4, pb_b + nb * pass_bytes(mb));
  if (e != cudaSuccess) return e;
  int *meta = (int *)g_oz.meta;
  const int *emaxA = meta, *emaxB = meta + 2 * m;
  int8_t *pB = (int8_t *)g_oz.buf;
  int8_t *pA[2]; uint8_t *R[2];
  for (int b = 0; b < 2; ++b) {
    pA[b] = pB + pb_b + (size_t)(b % nb) * pass_bytes(mb);
    R[b] = (uint8_t *)(pA[b] + pa_bytes(mb));
  }
  CrtStreams &cs = g_cs;
#define QB_CRT_TRY(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)
  QB_CRT_TRY(cudaEventRecord(cs.in, st));
  QB_CRT_TRY(cudaStreamWaitEvent(cs.sM, cs.in, 0)); QB_CRT_TRY(cudaStreamWaitEvent(cs.sA, cs.in, 0)); QB_CRT_TRY(cudaStreamWaitEvent(cs.sF, cs.in, 0));
  launch_crt_residues(a.B, n, k, a.sbj, a.sbl, emaxB, WB, N, Kp, pB, cs.sM);
  const std::vector<int64_t> pass_rows = oz_crt_pass_rows(m, mb, (cb || mb >= m) ? 0 : g_oz_pass_shape);
  int pass = 0;
  int64_t r0 = 0;
  for (; pass < (int)pass_rows.size(); r0 += pass_rows[pass], ++pass) {
    const int64_t mr = pass_rows[pass];
    const int64_t Mp = rup(mr, OZ_BM);
    const int b = pass & 1;
    /* residues of this pass's A rows: its buffer was last read by the tensor kernel of pass - 2 */
    if (pass >= 2) QB_CRT_TRY(cudaStreamWaitEvent(cs.sA, cs.mma[b], 0));
    launch_crt_residues(a.A + r0 * a.sai, mr, k, a.sai, a.sal, emaxA + r0, WA, N, Kp, pA[b], cs.sA);
    QB_CRT_TRY(cudaEventRecord(cs.resA[b], cs.sA));
    /* tensor kernel: needs the residues, and R[b] drained by the reconstruction of pass - 2 */
    QB_CRT_TRY(cudaStreamWaitEvent(cs.sM, cs.resA[b], 0));
    if (pass >= 2) QB_CRT_TRY(cudaStreamWaitEvent(cs.sM, cs.fold[b], 0));
    oz_ev_record(0, cs.sM);
    e = launch_crt_mma(pA[b], pB, N, mr, n, Kp, R[b], Mp, Np, cs.sM);
    oz_ev_record(1, cs.sM);
    if (e != cudaSuccess) return e;
    QB_CRT_TRY(cudaEventRecord(cs.mma[b], cs.sM));
    /* reconstruction + epilogue */
    QB_CRT_TRY(cudaStreamWaitEvent(cs.sF, cs.mma[b], 0));
    CrtFoldArgs f;
    f.R = R[b]; f.Mp = Mp; f.Np = Np; f.m = mr; f.n = n; f.row0 = r0;
    f.emaxA = emaxA; f.emaxB = emaxB; f.WA = WA; f.WB = WB;
    f.alpha = a.alpha; f.beta = a.beta; f.C = a.C + r0 * a.sci; f.sci = a.sci; f.scj = a.scj;
    f.simple = (a.alpha.hi == 0x3fff000000000000ULL && a.alpha.lo == 0 && (a.beta.hi & 0x7fffffffffffffffULL) == 0 && a.beta.lo == 0) ? 1 : 0;
    f.n4p = rup((n + 3) / 4, 32);
    f.npeer = a.npeer;
    for (int q = 0; q < QB_MAX_PEERS; ++q) f.peer[q] = q < a.npeer ? a.peerC[q] + r0 * a.sci : nullptr;
    launch_crt_fold(f, pl, cs.sF);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    QB_CRT_TRY(cudaEventRecord(cs.fold[b], cs.sF));
    e = cudaStreamWaitEvent(st, cs.fold[b], 0);   /* the caller's stream sees the rows of this pass */
    if (e != cudaSuccess) return e;
    if (cb) cb(r0, mr, cb_user);
  }
  g_last_stats.SA = (WA + 7) / 8; g_last_stats.SB = (WB + 7) / 8; g_last_stats.ndiag = N; g_last_stats.nchunks = 1;
  g_last_stats.row_passes = pass;
  g_last_stats.pairs = N; g_last_stats.keep = 0; g_last_stats.flagged = 0; g_last_stats.redo_passes = 0;
  g_last_stats.ws_bytes = (int64_t)g_oz.bytes; g_last_stats.Kp = Kp;
  g_last_stats.scheme = 1; g_last_stats.WA = WA; g_last_stats.WB = WB;
  g_last_stats.peer_written = a.npeer;
  *used = 1;
  return cudaSuccess;
#undef QB_CRT_TRY
}

/* The whole fast-mode GEMM.  *used = 0 means the planner declined (caller runs the integer kernel). */
cudaError_t launch_gemm_ozaki(const GemmArgs &a, cudaStream_t st, int *used, size_t ws_budget, oz_pass_cb cb, void *cb_user, int min_passes)
{
  *used = 0;
  g_ev_used = 0;
  { const int64_t keep = g_last_stats.ws_bytes; g_last_stats = OzStats(); g_last_stats.ws_bytes = keep; }
  if (!get_encode()) return cudaSuccess;
  const int64_t m = a.m, n = a.n, k = a.k;
  const int64_t Kp = rup(k, OZ_BK);
  /* ---- scan + plan ---- */
  const size_t meta_ints = (size_t)(2 * m + 2 * n + 16);
  cudaError_t e = oz_reserve(meta_ints * 4, 0);
  if (e != cudaSuccess) return e;
  int *meta = (int *)g_oz.meta;
  int *emaxA = meta, *lminA = meta + m, *emaxB = meta + 2 * m, *lminB = meta + 2 * m + n, *flags = meta + 2 * m + 2 * n, *plan = flags + 4;
  {
    cudaMemsetAsync(emaxA, 0, (size_t)m * 4, st); cudaMemsetAsync(lminA, 0x7f, (size_t)m * 4, st);
    cudaMemsetAsync(emaxB, 0, (size_t)n * 4, st); cudaMemsetAsync(lminB, 0x7f, (size_t)n * 4, st);
    cudaMemsetAsync(flags, 0, 64, st);
    const int64_t nch = (k + OZ_SCAN_CH - 1) / OZ_SCAN_CH;
    {
      const bool warp_mode = (a.sal == 1 || a.sai != 1);
      const int64_t threads = warp_mode ? m * nch * 32 : m * nch;
      k_oz_scan<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a.A, m, k, a.sai, a.sal, emaxA, lminA, flags);
    }
    {
      const bool warp_mode = (a.sbl == 1 || a.sbj != 1);
      const int64_t threads = warp_mode ? n * nch * 32 : n * nch;
      k_oz_scan<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a.B, n, k, a.sbj, a.sbl, emaxB, lminB, flags);
    }
    k_oz_plan<<<1, 1024, 0, st>>>(emaxA, lminA, m, emaxB, lminB, n, flags, plan);
    count_launch(3);
    e = cudaMemcpyAsync(g_oz.h_plan, plan, 16, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
  }
  const int WA = g_oz.h_plan[0], WB = g_oz.h_plan[1], fl = g_oz.h_plan[2];
  if (g_oz_scheme == 1 && fl == 0) { /* residue scheme: one int8 GEMM per modulus; declines (-> digit diagonals) when the moduli cannot cover the span */
    e = launch_gemm_crt(a, st, used, ws_budget, cb, cb_user, min_passes, WA, WB, meta_ints);
    if (e != cudaSuccess || *used) return e;
  }
  const int SA = std::max(1, (WA + 2 + 7) / 8), SB = std::max(1, (WB + 2 + 7) / 8);
  if (fl != 0 || SA > OZ_MAX_S || SB > OZ_MAX_S) return cudaSuccess; /* decline: Inf/NaN or span too wide */
  const int ndiag = SA + SB - 1;
  /* bounded setting: multiply only the leading diagonals (needs k >= 2 for the error budget, header comment) */
  const int keep = (g_oz_keep > 0 && g_oz_keep < ndiag && k >= 2) ? g_oz_keep : ndiag;
  const bool bounded = keep < ndiag;
  /* int32 exactness: Kc * min(SA, SB) * 2^14 <= 2^31 - 1 */
  int64_t kc_blocks = ((((int64_t)1 << 17) - 1) / std::min(SA, SB)) / OZ_BK;
  if (kc_blocks < 1) return cudaSuccess;
  const int64_t nkb_total = Kp / OZ_BK;
  const int nchunks = (int)((nkb_total + kc_blocks - 1) / kc_blocks);
  kc_blocks = (nkb_total + nchunks - 1) / nchunks;   /* equal chunks: a short last chunk runs the tensor kernel at a worse duty cycle */
  /* ---- workspace: planes B | planes A (per row pass) | D | W ---- */
  const int64_t Np = rup(n, OZ_BN);
  const size_t pb_b = rup((int64_t)SB * n * Kp, 1024);
  /* K chunks: either every chunk keeps its own int32 diagonal set and ONE fold sums them (`sets` = nchunks, no wide
   * workspace, least traffic), or - when that does not fit next to the full-diagonal redo buffer - one set is
   * folded after every chunk into the 448-bit running sums W */
  const int nd_main = keep;                      /* diagonals of the main sweep (== ndiag in the exact setting) */
  auto d_planes = [&](bool multi) -> int64_t { return multi ? std::max<int64_t>((int64_t)nchunks * nd_main, ndiag) : ndiag; };
  auto pass_bytes_v = [&](int64_t mb, bool multi) -> size_t {
    const int64_t Mp = rup(mb, OZ_BM);
    size_t b = rup((int64_t)SA * mb * Kp, 1024) + (size_t)d_planes(multi) * Mp * Np * 4;
    if (nchunks > 1) b += (size_t)OZ_NL * Mp * Np * 4;      /* W: always reserved (the redo sweep folds per chunk) */
    if (bounded) b += (size_t)rup(Mp * Np, 1024) + (size_t)oz_list_cap(mb, n) * sizeof(int2);
    return b;
  };
  const int64_t mb_hint = (cb && min_passes > 1) ? std::max<int64_t>(OZ_BM, rup((m + min_passes - 1) / min_passes, OZ_BM)) : m;
  const bool multi = nchunks > 1 && pb_b + pass_bytes_v(mb_hint, true) <= ws_budget;
  auto pass_bytes = [&](int64_t mb) -> size_t { return pass_bytes_v(mb, multi); };
  int64_t mb = m;
  if (cb && min_passes > 1) mb = std::max<int64_t>(OZ_BM, rup((m + min_passes - 1) / min_passes, OZ_BM));
  while (mb > OZ_BM && pb_b + pass_bytes(mb) > ws_budget) mb = rup((mb + 1) / 2, OZ_BM);
  if (pb_b + pass_bytes(mb) > ws_budget) return cudaSuccess; /* does not fit: decline */
  e = oz_reserve(meta_ints * 4, pb_b + pass_bytes(mb));
  if (e != cudaSuccess) return e;
  int8_t *pB = (int8_t *)g_oz.buf;
  int8_t *pA = pB + pb_b;
  /* ---- slice B once ---- */
  launch_oz_slice(a.B, n, k, a.sbj, a.sbl, emaxB, SB, Kp, pB, st);
  int64_t flagged_total = 0, pairs_done = 0;
  int redo_passes = 0;
  auto npairs = [&](int nd) { int64_t p = 0; for (int d = 0; d < nd; ++d) p += std::min(d, SA - 1) - std::max(0, d - (SB - 1)) + 1; return p; };
  int *counter = flags + 8;   /* meta: one int, zeroed per row pass */
  for (int64_t r0 = 0; r0 < m; r0 += mb) {
    const int64_t mr = std::min(mb, m - r0);
    const int64_t Mp = rup(mr, OZ_BM);
    int32_t *D = (int32_t *)(pA + rup((int64_t)SA * mb * Kp, 1024));
    uint32_t *W = (uint32_t *)(D + (size_t)d_planes(multi) * rup(mb, OZ_BM) * Np);
    uint8_t *flagp = (uint8_t *)(W + (nchunks > 1 ? (size_t)OZ_NL * rup(mb, OZ_BM) * Np : 0));
    int2 *list = (int2 *)(flagp + rup(rup(mb, OZ_BM) * Np, 1024));
    const int list_cap = (int)oz_list_cap(mb, n);
    launch_oz_slice(a.A + r0 * a.sai, mr, k, a.sai, a.sal, emaxA + r0, SA, Kp, pA, st);
    /* one sweep over the K chunks with `nd` leading diagonals; `redo` = exact redo of the flagged elements */
    auto sweep = [&](int nd, bool check, bool redo) -> cudaError_t {
      const bool sets = multi && !redo;          /* every chunk into its own diagonal set, one fold at the end */
      const int64_t set_stride = (int64_t)nd * Mp * Np;
      for (int c = 0; c < nchunks; ++c) {
        const int kb0 = (int)(c * kc_blocks), nkb = (int)std::min<int64_t>(kc_blocks, nkb_total - kb0);
        oz_ev_record(0, st);
        cudaError_t e2 = launch_oz_mma(pA, pB, SA, SB, mr, n, Kp, kb0, nkb, D + (sets ? c * set_stride : 0), Mp, Np, st, nd);
        oz_ev_record(1, st);
        if (e2 != cudaSuccess) return e2;
        if (sets && c + 1 < nchunks) continue;
        OzFoldArgs f;
        f.D = D; f.Mp = Mp; f.Np = Np; f.ndiag = nd; f.m = mr; f.n = n; f.row0 = r0;
        f.nsets = sets ? nchunks : 1; f.set_stride = set_stride;
        f.emaxA = emaxA; f.emaxB = emaxB; f.SA = SA; f.SB = SB;
        f.W = W; f.w_in = !sets && c > 0; f.w_out = !sets && c + 1 < nchunks;
        f.alpha = a.alpha; f.beta = a.beta; f.C = a.C + r0 * a.sci; f.sci = a.sci; f.scj = a.scj;
        f.exp8 = ndiag - nd; f.check = check ? 1 : 0; f.only_flagged = redo ? 1 : 0;
        f.flag = flagp; f.list = list; f.list_cap = list_cap; f.counter = counter;
        const int64_t elems = mr * n;
        const unsigned fg = (unsigned)((elems + 255) / 256);
        switch (f.nsets) {
        case 1: k_oz_fold<1><<<fg, 256, 0, st>>>(f); break;
        case 2: k_oz_fold<2><<<fg, 256, 0, st>>>(f); break;
        case 3: k_oz_fold<3><<<fg, 256, 0, st>>>(f); break;
        case 4: k_oz_fold<4><<<fg, 256, 0, st>>>(f); break;
        case 5: k_oz_fold<5><<<fg, 256, 0, st>>>(f); break;
        default: k_oz_fold<0><<<fg, 256, 0, st>>>(f); break;
        }
        count_launch();
        e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return e2;
      }
      pairs_done += npairs(nd);
      return cudaSuccess;
    };
    if (!bounded) {
      e = sweep(ndiag, false, false);
      if (e != cudaSuccess) return e;
      if (cb) cb(r0, mr, cb_user);
      continue;
    }
    cudaMemsetAsync(counter, 0, 4, st);
    e = sweep(keep, true, false);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(g_oz.h_plan + 8, counter, 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    const int nflag = g_oz.h_plan[8];
    flagged_total += nflag;
    if (nflag == 0) { if (cb) cb(r0, mr, cb_user); continue; }
    if (nflag <= list_cap) {
      OzFixArgs x;
      x.list = list; x.counter = counter; x.list_cap = list_cap;
      x.A = a.A; x.sai = a.sai; x.sal = a.sal; x.B = a.B; x.sbl = a.sbl; x.sbj = a.sbj; x.k = k;
      x.alpha = a.alpha; x.beta = a.beta; x.C = a.C; x.sci = a.sci; x.scj = a.scj;
      const int blocks = (int)std::min<int64_t>((nflag + 3) / 4, (int64_t)sm_count() * 8);
      k_oz_fixup<<<blocks, 128, 0, st>>>(x);
      count_launch();
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    } else { /* structured cancellation: all diagonals, written for the flagged elements only */
      ++redo_passes;
      e = sweep(ndiag, false, true);
      if (e != cudaSuccess) return e;
    }
    if (cb) cb(r0, mr, cb_user);
  }
  g_last_stats.SA = SA; g_last_stats.SB = SB; g_last_stats.ndiag = ndiag; g_last_stats.nchunks = nchunks;
  g_last_stats.row_passes = (int)((m + mb - 1) / mb);
  g_last_stats.pairs = pairs_done / g_last_stats.row_passes;   /* digit-plane products per row pass (bounded + any redo) */
  g_last_stats.keep = keep; g_last_stats.flagged = flagged_total; g_last_stats.redo_passes = redo_passes;
  g_last_stats.ws_bytes = (int64_t)g_oz.bytes; g_last_stats.Kp = Kp;
  *used = 1;
  return cudaSuccess;
}

} // namespace qb

/*
 * qb_ozaki.cu — fast-mode binary128 GEMM on the 5th-generation tensor cores (residue scheme).
 *
 * What it replaces: the arithmetic of QuadBLAS::gemm (/root/reference/include/quadblas/algorithms/
 * level3.hpp:215-336; hot loop :77-85 = one Sleef_fmaq2_u05 per two products) when the library is in
 * QB_MODE_FAST.  Fast mode is free to re-associate (last paragraph); this path
 * computes every inner product as an exact integer and rounds once:
 *
 *   scan     : per row of op(A) / column of op(B): largest exponent, lowest set mantissa bit, Inf/NaN flag   (k_oz_scan)
 *   plan     : widest bit spans -> windows W_A, W_B and the number of moduli N (crt::host::plan_windows).  Rows / columns
 *              become block fixed point, x = X 2^(base - 16495), |X| < 2^W; the windows normally cover the spans (X exact);
 *              spans wider than the moduli can cover (exponent spreads of +-40 binades and more) are cut: the low bits of the
 *              small elements of a row are dropped and every C element is tested against the dropped mass (below)
 *   residues : a_i = X mod p_i as symmetric int8, one byte plane per modulus, K-major                      (k_crt_residues*)
 *   mma      : R_i = (A_i B_i^T) mod p_i — ONE int8 x int8 -> int32 GEMM per modulus on tcgen05 (kind::i8), operands staged
 *              by TMA into 128-byte swizzled shared memory, accumulators in TMEM, warp-specialised persistent kernel,
 *              the accumulator reduced mod p_i in the epilogue and stored as one byte                      (k_oz_mma<1>)
 *   fold     : Chinese-remainder reconstruction of the exact integer (crt::reconstruct_dev), ONE correctly rounded conversion
 *              to binary128, then the reference epilogue C = fma(alpha, s, mul(beta, C)) (level3.hpp:102-109) (k_crt_fold)
 *   fix-up   : C elements the fold does not accept — the reconstructed integer is too small against what a capped window
 *              dropped (crt::accept_msb; cancellation), or the row / column holds Inf / NaN — are recomputed from the
 *              original operands in the unrounded window accumulator of qwide.cuh, one warp per element      (k_crt_fixup)
 *
 * The work is cut into units (row pass of A) x (column panel of B) and software-pipelined over four internal streams:
 * tensor kernel of unit u || residues of the next A pass / B panel || reconstruction of unit u-1 (launch_gemm_crt).
 *
 * Accuracy: windows that cover the spans (every D53 / D113 input of BASELINE configs 1-4): the inner product is exact and
 * rounded once, |c^ - c| <= u |A B|_ij (+ the two epilogue roundings).  Capped windows: every accepted element satisfies
 * |c^ - c| <= k u (|A||B|)_ij by the test in crt::accept_msb, every other element comes from the window accumulator
 * (error < k 2^-133 max|a b| + one rounding); both are inside the fast-mode contract gamma_k (|A||B|)_ij.
 */
#include "qb_internal.h"
#include "qb_tc.cuh"
#include "q128_chain.cuh"
#include "qwide.cuh"
#include "qb_crt.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
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
static constexpr int OZ_SCAN_CH = 256;       /* k per scan work item: 8192^2 gives 262k warps / threads, enough loads in flight for HBM */
static constexpr int OZ_KCHUNK = 65536;      /* K per int32 accumulation: 128 * 128 * 65536 = 2^30 */

struct OzMmaArgs {
  int32_t *D;                 /* plain mode (CRT = 0): [ndiag][Mp][Np] int32 */
  int64_t Mp, Np;             /* padded to tile multiples */
  int SA, SB, ndiag;
  int m_tiles, n_tiles;
  int kb_begin, nkb;          /* k-blocks of OZ_BK of this launch */
  uint8_t order[OZ_MAX_DIAG + 1];
  uint8_t *R;                 /* residue mode (CRT = 1): N planes of [Mp][Np] bytes `rplane` bytes apart, R_i = (A_i B_i^T) mod p_i in [0, p_i); ndiag = N */
  int64_t rplane;
  int accum;                  /* residue mode: add to the residues already in R (K chunks beyond OZ_KCHUNK) */
};

/* per-modulus constants of the residue scheme (qb_crt.cuh); uploaded once per device */
__constant__ crt::Tables c_crt;

/* Tile order inside one modulus: bands of OZ_GM m-tiles, n-tile outer / m-tile inner inside a band, so that the ~148 tiles
 * in flight form a compact block of C: every A row panel is shared by ~9 CTAs and every B column panel by 16 through L2
 * (profiles/r1c: the n-fastest order re-read the planes from HBM 3.4x more often than this needs at 8192^3). */
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
/* CRT = 1: residue planes, one product A_i B_i^T per modulus, reduced mod p_i in the epilogue and stored as bytes.
 * CRT = 0: the same pipeline with int32 output, D_d = sum_{s+t=d} A_s B_t^T over plane pairs (qb_oz_i8gemm_dev: the int8 peak
 * microbenchmark and the standalone test of the tcgen05 path). */
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
    /* ===================== epilogue: TMEM -> registers -> global ===================== */
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
        uint8_t *dst = g.R + (int64_t)d * g.rplane + row * g.Np + (int64_t)nt * OZ_BN;
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
          if (g.accum) { /* a later K chunk: (old + new) mod p, byte by byte */
            const uint4 o0 = *reinterpret_cast<const uint4 *>(dst + c * 32), o1 = *reinterpret_cast<const uint4 *>(dst + c * 32 + 16);
            const uint32_t o[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              uint32_t word = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                uint32_t r = ((w[q] >> (8 * b)) & 0xffu) + ((o[q] >> (8 * b)) & 0xffu);
                r = r >= p ? r - p : r;
                word |= r << (8 * b);
              }
              w[q] = word;
            }
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

/* rows x K view: X[r * sr + k * sk].  Over the finite non-zero elements of row r: emax[r] = max ee (0 if none),
 * lmin[r] = min (ee + tz(M)); sp[r] = 1 when the row holds an Inf or NaN (those elements count as zero for the tensor kernel,
 * every C element of the row goes to the fix-up).  One warp per (row, OZ_SCAN_CH-wide k chunk) when k is the contiguous
 * direction, otherwise one thread per (row, chunk) with lanes along rows.  emax / sp start at 0, lmin at a large value. */
__global__ void k_oz_scan(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, int *emax, int *lmin, int *sp)
{
  constexpr int CH = OZ_SCAN_CH;
  const int64_t nchunk = (K + CH - 1) / CH;
  int em = 0, lm = 0x7fffffff, spc = 0;
  int64_t r;
  if (sk == 1 || sr != 1) { /* warp per (row, chunk), lanes along k */
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= rows * nchunk) return;
    r = w / nchunk;
    const int64_t k0 = (w % nchunk) * CH, k1 = k0 + CH < K ? k0 + CH : K;
    for (int64_t k = k0 + lane; k < k1; k += 32) {
      const OzElem e = oz_unpack(X[r * sr + k * sk]);
      spc |= e.special;
      if (!e.special && (e.lo | e.hi)) { em = max(em, e.ee); lm = min(lm, e.ee + oz_tz(e.lo, e.hi)); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      em = max(em, __shfl_xor_sync(0xffffffffu, em, o));
      lm = min(lm, __shfl_xor_sync(0xffffffffu, lm, o));
      spc |= __shfl_xor_sync(0xffffffffu, spc, o);
    }
    if (lane != 0) return;
  } else { /* rows contiguous: thread per (row, chunk), lanes along rows */
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * nchunk) return;
    r = t % rows;
    const int64_t k0 = (t / rows) * CH, k1 = k0 + CH < K ? k0 + CH : K;
    for (int64_t k = k0; k < k1; ++k) {
      const OzElem e = oz_unpack(X[r * sr + k * sk]);
      spc |= e.special;
      if (!e.special && (e.lo | e.hi)) { em = max(em, e.ee); lm = min(lm, e.ee + oz_tz(e.lo, e.hi)); }
    }
  }
  if (em) { atomicMax(&emax[r], em); atomicMin(&lmin[r], lm); }
  if (spc) sp[r] = 1;
}

/* out[0] = widest span of A rows, out[1] = widest span of B columns, out[2] = 1 when any row / column holds Inf or NaN */
__global__ void k_oz_plan(const int *emaxA, const int *lminA, const int *spA, int64_t m, const int *emaxB, const int *lminB, const int *spB, int64_t n, int *out)
{
  __shared__ int sw[3];
  if (threadIdx.x == 0) { sw[0] = 0; sw[1] = 0; sw[2] = 0; }
  __syncthreads();
  int wa = 0, wb = 0, fl = 0;
  for (int64_t i = threadIdx.x; i < m; i += blockDim.x) { if (emaxA[i]) wa = max(wa, emaxA[i] + 113 - lminA[i]); fl |= spA[i]; }
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) { if (emaxB[j]) wb = max(wb, emaxB[j] + 113 - lminB[j]); fl |= spB[j]; }
  atomicMax(&sw[0], wa);
  atomicMax(&sw[1], wb);
  if (fl) atomicOr(&sw[2], 1);
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = sw[0]; out[1] = sw[1]; out[2] = sw[2]; }
}

/* ------------------------------------------------------------------ residues */
/* The 4 elements (row r, k = 4 g4 .. 4 g4 + 3) as the words residue_sym consumes (crt::element_words: |X| or its two's
 * complement, X = trunc(x / 2^(base - 16495)) < 2^W). */
struct Crt4 { uint32_t w[4][crt::NWMAX]; uint32_t sign[4]; };
template <int NW>
__device__ __forceinline__ void crt_load4(const q128 *__restrict__ X, int64_t r, int64_t g4, int64_t K, int64_t sr, int64_t sk, int base, Crt4 &c)
{
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t k = g4 * 4 + q;
    if (k >= K) {
#pragma unroll
      for (int j = 0; j < crt::NWMAX; ++j) c.w[q][j] = 0;
      c.sign[q] = 0;
      continue;
    }
    crt::element_words<NW>(X[r * sr + k * sk], base, c.w[q], c.sign[q]);
  }
}
/* residues of the 4 elements modulo p_i packed as 4 int8 (byte q = element q) */
template <int NW>
__device__ __forceinline__ uint32_t crt_word(const Crt4 &c, int i)
{
  const uint32_t r0 = crt::residue_sym<NW>(c.w[0], c.sign[0], i, c_crt), r1 = crt::residue_sym<NW>(c.w[1], c.sign[1], i, c_crt);
  const uint32_t r2 = crt::residue_sym<NW>(c.w[2], c.sign[2], i, c_crt), r3 = crt::residue_sym<NW>(c.w[3], c.sign[3], i, c_crt);
  return __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);   /* low bytes of r0..r3 */
}

/* k contiguous (or fully strided): thread = (row, 4 consecutive k); plane i is [rows][Kp] int8 */
template <int NW>
__global__ void __launch_bounds__(256) k_crt_residues(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax,
                                                      int W, int N, int64_t Kp, int8_t *__restrict__ planes, int64_t pstride)
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
        if (i < N) *reinterpret_cast<uint32_t *>(planes + (int64_t)i * pstride + r * Kp + g4 * 4) = crt_word<NW>(c, i);
      }
    }
  }
}

/* rows contiguous in memory (B of a row-major product, A of a col-major one): CTA = 16 rows x 32 k, 128 threads with the 16 lanes of
 * a half-warp along the rows (256-byte contiguous loads); the plane words are transposed through shared memory and leave as
 * 32-byte runs along K (the direct store would scatter 4-byte words Kp bytes apart).  The tile is this small on purpose:
 * N * 16 * 9 words (23.6 KB at 41 moduli, 28.2 KB at 49) fit NEXT TO the persistent tensor kernel's 193 KB of shared memory, so
 * the residues of the next B panel are computed on the integer pipes while the tensor kernel of the current panel runs. */
static inline int crt_t_smem(int N) { return N * 16 * 9 * 4; }
template <int NW>
__global__ void __launch_bounds__(128) k_crt_residues_t(const q128 *__restrict__ X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *__restrict__ emax,
                                                        int W, int N, int64_t Kp, int8_t *__restrict__ planes, int64_t pstride)
{
  extern __shared__ uint32_t crt_sm[];            /* [plane][row][k-group], padded: 9 words per row */
  const int row_l = threadIdx.x & 15, kg = threadIdx.x >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * 16, g0 = (int64_t)blockIdx.y * 8;
  const int64_t r = r0 + row_l, g4 = g0 + kg;
  if (r < rows) {
    Crt4 c;
    crt_load4<NW>(X, r, g4, K, sr, sk, emax[r] + 113 - W, c);
#pragma unroll
    for (int g = 0; g < crt::NGMAX; ++g) {
      if (4 * g < N) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = 4 * g + b;
          if (i < N) crt_sm[(i * 16 + row_l) * 9 + kg] = crt_word<NW>(c, i);
        }
      }
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < N * 32; q += 128) {
    const int i = q >> 5, row = (q & 31) >> 1, half = q & 1;
    if (r0 + row >= rows) continue;
    const uint32_t *src = &crt_sm[(i * 16 + row) * 9 + half * 4];
    *reinterpret_cast<uint4 *>(planes + (int64_t)i * pstride + (r0 + row) * Kp + (g0 + half * 4) * 4) = make_uint4(src[0], src[1], src[2], src[3]);
  }
}

/* ------------------------------------------------------------------ fold: residues -> binary128 */
struct CrtFoldArgs {
  const uint8_t *R; int64_t Mp, Np, rplane;  /* residues of this unit: N planes of [Mp][Np] bytes, rplane bytes apart */
  int64_t m, n, row0, col0;          /* the unit is C rows [row0, row0 + m) x columns [col0, col0 + n) of the call's C */
  const int *emaxA, *lminA, *spA, *emaxB, *lminB, *spB; int WA, WB;   /* per row / column of the whole call */
  int64_t k;
  q128 alpha, beta; q128 *C; int64_t sci, scj;
  int simple;                        /* alpha == 1 and beta == +-0 */
  int64_t n4p;                       /* threads per row: ceil(n / 4) rounded up to 32 */
  int check;                         /* capped windows or Inf / NaN present: test every element (crt::accept_msb), rejected ones are listed */
  int2 *list; int list_cap; int *counter; uint8_t *flag;   /* rejected elements: (row, column) of the call's C; beyond list_cap: flag[i * Np + j] = 1 */
  int npeer; q128 *peer[QB_MAX_PEERS]; /* fused gather: the call's C inside each peer GPU's buffer (or one NVSwitch multicast address) */
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
    const int64_t plane = g.rplane, off = i * g.Np + j0;
    uint32_t rw[4 * NG];
#pragma unroll
    for (int c = 0; c < 4 * NG; ++c) rw[c] = c < pl.N ? *reinterpret_cast<const uint32_t *>(g.R + (int64_t)c * plane + off) : 0u;
    const int64_t gi = g.row0 + i;
    const int eA = g.emaxA[gi];
    const int baseA = eA + 113 - g.WA;
    bool tA = false, sA = false;
    if (g.check) { tA = eA != 0 && g.lminA[gi] < baseA; sA = g.spA[gi] != 0; }
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
      const int64_t j = j0 + e;
      if (j >= g.n) break;
      const int64_t gj = g.col0 + j;
      uint32_t r[crt::NMP];
#pragma unroll
      for (int c = 0; c < 4 * NG; ++c) r[c] = (rw[c] >> (8 * e)) & 0xffu;
      uint32_t Y[NG + 1], neg;
      crt::reconstruct_dev<NG>(r, pl, Y, neg);
      const int eB = g.emaxB[gj];
      const int baseB = eB + 113 - g.WB;
      q128 *c = g.C + gi * g.sci + gj * g.scj;
      const q128 cin = *c;
      q128 out;
      bool reject = false;
      if (g.check) {
        const bool tB = eB != 0 && g.lminB[gj] < baseB;
        reject = sA || g.spB[gj] != 0 || crt::limbs_msb<NG + 1>(Y) < crt::accept_msb(tA, tB, g.WA, g.WB, g.k);
      }
      if (reject) { /* left to k_crt_fixup (C_in stays in place for its epilogue) */
        const int slot = atomicAdd(g.counter, 1);
        if (slot < g.list_cap) g.list[slot] = make_int2((int)gi, (int)gj);
        else g.flag[i * g.Np + j] = 1;
        out = cin;
      } else {
        const q128 sum = crt::limbs_to_q<NG + 1>(Y, neg, baseA + baseB - 2 * 16495);
        /* alpha = 1, beta = +-0, finite C: mul(beta, C) = +-0 and fma(1, s, +-0) = s (s is never -0) - the same bits as the
         * general line below, without the two software roundings */
        if (g.simple && ((cin.hi >> 48) & 0x7fffu) != 0x7fffu) out = sum;
        else out = q_fma(g.alpha, sum, q_mul(g.beta, cin)); /* beta*C is always evaluated */
        *c = out;
      }
      if (stage) fold_sm[warp * 128 + 4 * lane + e] = make_uint4((uint32_t)out.lo, (uint32_t)(out.lo >> 32), (uint32_t)out.hi, (uint32_t)(out.hi >> 32));
      else if (!reject)
        for (int q = 0; q < g.npeer; ++q) g.peer[q][gi * g.sci + gj * g.scj] = out;   /* col-major C: element-wise peer stores */
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
          for (int q = 0; q < g.npeer; ++q) *reinterpret_cast<uint4 *>(g.peer[q] + (g.row0 + i) * g.sci + g.col0 + col) = v;
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

/* ------------------------------------------------------------------ fix-up of the rejected elements */
struct CrtFixArgs {
  const int2 *list; const int *counter; int list_cap; const uint8_t *flag;
  int64_t um, un, uNp, row0, col0;   /* the unit (for the flag scan when the list overflowed) */
  const q128 *A; int64_t sai, sal; const q128 *B; int64_t sbl, sbj; int64_t k;
  q128 alpha, beta; q128 *C; int64_t sci, scj;
  int *total;                        /* running count of fixed elements of the call (stats) */
  int npeer; q128 *peer[QB_MAX_PEERS];
};
__device__ __noinline__ qwide crt_merge(qwide a, qwide b) { qw_merge(a, b); return a; }
/* One warp per rejected C element: the k products go into the unrounded window accumulator of qwide.cuh (lanes stride over k),
 * a shuffle tree merges the 32 windows and the Inf / NaN classes, lane 0 rounds once and applies the reference epilogue.
 * Elements come from the list; when more were rejected than it holds, the rest is found by scanning the unit's flag bytes. */
__global__ void __launch_bounds__(128) k_crt_fixup(const CrtFixArgs g)
{
  __shared__ uint32_t scr[QWA_COL_WORDS * 128];
  const int lane = threadIdx.x & 31;
  uint32_t *col = scr + threadIdx.x;
  qwa_col_init(col, 128);
  const int count = *g.counter;
  if (count == 0) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(g.total, count);
  const int64_t nwarps = (int64_t)gridDim.x * 4, wid = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t listed = count < g.list_cap ? count : g.list_cap;
  const int64_t scan = count > g.list_cap ? g.um * g.uNp : 0;
  for (int64_t e = wid; e < listed + scan; e += nwarps) {
    int64_t gi, gj;
    if (e < listed) { const int2 ij = g.list[e]; gi = ij.x; gj = ij.y; }
    else {
      const int64_t f = e - listed;
      if (!g.flag[f]) continue;
      gi = g.row0 + f / g.uNp; gj = g.col0 + f % g.uNp;
    }
    const q128 *ap = g.A + gi * g.sai, *bp = g.B + gj * g.sbj;
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
      bad |= __shfl_down_sync(0xffffffffu, bad, o);
      v = crt_merge(v, t);
    }
    if (lane == 0) {
      const int64_t off = gi * g.sci + gj * g.scj;
      const q128 out = q_fma(g.alpha, qw_finish(v, bad), q_mul(g.beta, g.C[off]));
      g.C[off] = out;
      for (int q = 0; q < g.npeer; ++q) g.peer[q][off] = out;
    }
  }
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

/* 2-D map over a matrix of 16-byte elements (qb_level2.cu: tiles of A for the sliced qgemv).  swizzle128: dims {inner * 4 uint32,
 * outer}, box {box_inner * 4, box_outer} with box_inner * 16 = 128 bytes; else dims {inner * 2 uint64, outer}, box {box_inner * 2,
 * box_outer}.  stride = bytes between consecutive outer indices.  Out-of-range elements read as zero. */
bool make_quad_map(CUtensorMap *tm, const void *base, int64_t inner, int64_t outer, int64_t stride_bytes, int box_inner, int box_outer, bool swizzle128)
{
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  const int per = swizzle128 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)inner * per, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)stride_bytes};
  cuuint32_t box[2] = {(cuuint32_t)(box_inner * per), (cuuint32_t)box_outer};
  cuuint32_t es[2] = {1, 1};
  return enc(tm, swizzle128 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void *>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* planes [S][rows][Kp] int8 -> 3-D map {Kp, rows, S}, box {128, box_rows, 1}, 128-byte swizzle; rows
 * beyond `rows` are zero-filled by the TMA unit */
static bool make_plane_map(CUtensorMap *tm, const int8_t *planes, int S, int64_t rows, int64_t Kp, int box_rows, int64_t pstride = 0)
{
  PFN_encodeTiled enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)Kp, pstride > 0 ? (cuuint64_t)pstride : (cuuint64_t)Kp * (cuuint64_t)rows};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)planes, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* ---- per-device state (callers hold the library mutex): streams, events, workspaces, one-time kernel attributes.
 * A process that switches the current device gets an independent set for each device. ---- */
static constexpr int OZ_MAX_EV = 1024;       /* timed tensor-kernel launches per qgemm */
enum { EV_IN = 0, EV_SCAN_A, EV_SCAN_B, EV_DONE, EV_MMA0, EV_MMA1, EV_FOLD0, EV_FOLD1, EV_SLOT0 };
struct OzDev {
  bool ready = false;
  int sm_count = 148;
  void *buf = nullptr; size_t bytes = 0;          /* grow-only: residue planes, residues, flags, list */
  void *meta = nullptr; size_t meta_bytes = 0;    /* per-row exponent data + plan words (small, survives a regrowth of buf) */
  int *h_plan = nullptr;                          /* pinned, 16 ints */
  cudaStream_t sM = nullptr, sA = nullptr, sB = nullptr, sF = nullptr;
  std::vector<cudaEvent_t> ev;                    /* ordering events (no timing) */
  std::vector<cudaEvent_t> tev;                   /* timing events around the tensor kernel */
  int tev_used = 0;
  bool stats_pending = false;
};
static OzDev g_dev[QB_MAX_DEVICES];

static cudaError_t oz_dev(OzDev **out)
{
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= QB_MAX_DEVICES) return cudaErrorInvalidDevice;
  OzDev &D = g_dev[dev];
  *out = &D;
  if (D.ready) return cudaSuccess;
  cudaDeviceGetAttribute(&D.sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (D.sm_count <= 0) D.sm_count = 148;
  e = cudaFuncSetAttribute(k_oz_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_oz_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM);
  if (e != cudaSuccess) return e;
  static crt::Tables T;
  crt::host::build_tables(T);
  e = cudaMemcpyToSymbol(c_crt, &T, sizeof(T));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();   /* once per device: the copy from pageable memory must have landed before any stream reads it */
  if (e != cudaSuccess) return e;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t s[4] = {nullptr, nullptr, nullptr, nullptr};
  /* priorities: the tensor kernel first; then the reconstruction (it frees the residue buffer the tensor kernel of unit + 2 needs) and
   * the residues of the A passes (the tensor kernel of the NEXT unit needs them during the first panel); the residues of the B panels
   * (needed a whole panel ahead) last.  Blocks of a higher-priority kernel are scheduled before pending blocks of a lower one.
   * QBLAS_STREAM_PRIO=abcd overrides (digits = priority levels below the highest for sM, sA, sB, sF; experiments). */
  int off[4] = {0, 1, 2, 1};
  if (const char *v = getenv("QBLAS_STREAM_PRIO")) for (int i = 0; i < 4 && v[i] >= '0' && v[i] <= '9'; ++i) off[i] = v[i] - '0';
  for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaStreamCreateWithPriority(&s[i], cudaStreamNonBlocking, hi < lo ? std::min(lo, hi + off[i]) : hi);
  if (e == cudaSuccess) e = cudaMallocHost((void **)&D.h_plan, 64);
  if (e != cudaSuccess) { for (int i = 0; i < 4; ++i) if (s[i]) cudaStreamDestroy(s[i]); return e; }
  D.sM = s[0]; D.sA = s[1]; D.sB = s[2]; D.sF = s[3];
  D.ready = true;
  return cudaSuccess;
}
/* one-time setup of the current device (constant tables, kernel attributes, streams) for entry points that launch kernels directly */
cudaError_t oz_prepare_device() { OzDev *D; return oz_dev(&D); }
static cudaError_t oz_event(OzDev &D, int idx, cudaEvent_t *out)
{
  while ((int)D.ev.size() <= idx) {
    cudaEvent_t e;
    const cudaError_t r = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (r != cudaSuccess) return r;
    D.ev.push_back(e);
  }
  *out = D.ev[idx];
  return cudaSuccess;
}
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
void oz_release()
{
  OzDev *D;
  if (oz_dev(&D) != cudaSuccess) return;
  if (D->buf) cudaFree(D->buf);
  if (D->meta) cudaFree(D->meta);
  D->buf = nullptr; D->bytes = 0; D->meta = nullptr; D->meta_bytes = 0;
}

static inline int64_t rup(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static OzStats g_last_stats;
OzStats oz_last_stats(bool wait)
{
  OzDev *D;
  if (wait && g_last_stats.pairs > 0 && oz_dev(&D) == cudaSuccess && D->stats_pending) { /* the fix-up count arrives with the call's last work */
    cudaEvent_t done;
    if (oz_event(*D, EV_DONE, &done) == cudaSuccess && cudaEventSynchronize(done) == cudaSuccess) g_last_stats.flagged = D->h_plan[8];
    D->stats_pending = false;
  }
  return g_last_stats;
}

/* CUDA events around every k_oz_mma launch of the last qgemm, on the launching stream (bench.py's
 * roofline needs the kernel's own duration, not the whole call's) */
static void oz_ev_record(OzDev &D, int which, cudaStream_t st)
{
  if (D.tev_used >= OZ_MAX_EV) return;
  while ((int)D.tev.size() < 2 * (D.tev_used + 1)) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; D.tev.push_back(e); }
  cudaEventRecord(D.tev[2 * D.tev_used + which], st);
  if (which == 1) ++D.tev_used;
}
/* per-launch (start relative to the first launch, duration) in ms of the tensor-kernel launches of the last qgemm (profiling aid) */
int oz_last_mma_timeline(double *out, int max_pairs)
{
  OzDev *D;
  if (oz_dev(&D) != cudaSuccess) return 0;
  int n = 0;
  for (int i = 0; i < D->tev_used && n < max_pairs; ++i, ++n) {
    float t0 = 0, dt = 0;
    if (cudaEventSynchronize(D->tev[2 * i + 1]) != cudaSuccess) break;
    cudaEventElapsedTime(&t0, D->tev[0], D->tev[2 * i]);
    cudaEventElapsedTime(&dt, D->tev[2 * i], D->tev[2 * i + 1]);
    out[2 * n] = t0; out[2 * n + 1] = dt;
  }
  return n;
}
/* blocks until the last recorded launch has finished; returns the summed duration in ms */
double oz_last_mma_ms(int *launches)
{
  OzDev *D;
  double tot = 0;
  if (launches) *launches = 0;
  if (oz_dev(&D) != cudaSuccess) return 0;
  for (int i = 0; i < D->tev_used; ++i) {
    float ms = 0;
    if (cudaEventSynchronize(D->tev[2 * i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, D->tev[2 * i], D->tev[2 * i + 1]) == cudaSuccess) tot += ms;
  }
  if (launches) *launches = D->tev_used;
  return tot;
}

/* plain int8 GEMM on the tcgen05 pipeline: D[d] = sum_{s+t=d} A_s B_t^T over k-blocks [kb_begin, kb_begin + nkb) */
cudaError_t launch_oz_mma(const int8_t *pA, const int8_t *pB, int SA, int SB, int64_t m, int64_t n, int64_t Kp, int kb_begin, int nkb,
                          int32_t *Dout, int64_t Mp, int64_t Np, cudaStream_t st, int keep)
{
  OzDev *D;
  cudaError_t e = oz_dev(&D);
  if (e != cudaSuccess) return e;
  CUtensorMap tmA, tmB;
  if (!make_plane_map(&tmA, pA, SA, m, Kp, OZ_BM) || !make_plane_map(&tmB, pB, SB, n, Kp, OZ_BN)) return cudaErrorInvalidValue;
  OzMmaArgs g;
  memset(&g, 0, sizeof(g));
  g.D = Dout; g.R = nullptr; g.Mp = Mp; g.Np = Np; g.SA = SA; g.SB = SB; g.ndiag = SA + SB - 1;
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
  const int grid = std::min(total, D->sm_count);
  k_oz_mma<0><<<grid, OZ_THREADS, OZ_SMEM, st>>>(tmA, tmB, g);
  count_launch();
  return cudaGetLastError();
}

/* residue scheme: R_i (+)= (A_i B_i^T) mod p_i for the N residue planes over k-blocks [kb_begin, kb_begin + nkb)
 * (|acc| <= 128 * 128 * OZ_KCHUNK = 2^30) */
static cudaError_t launch_crt_mma(OzDev &D, const int8_t *pA, const int8_t *pB, int N, int64_t m, int64_t n, int64_t Kp, int kb_begin, int nkb, int accum,
                                  uint8_t *R, int64_t Mp, int64_t Np, int64_t rplane, cudaStream_t st, int64_t pstrideB = 0)
{
  CUtensorMap tmA, tmB;
  if (!make_plane_map(&tmA, pA, N, m, Kp, OZ_BM) || !make_plane_map(&tmB, pB, N, n, Kp, OZ_BN, pstrideB)) return cudaErrorInvalidValue;
  OzMmaArgs g;
  memset(&g, 0, sizeof(g));
  g.R = R; g.rplane = rplane; g.Mp = Mp; g.Np = Np; g.SA = N; g.SB = N; g.ndiag = N;
  g.m_tiles = (int)(Mp / OZ_BM); g.n_tiles = (int)(Np / OZ_BN);
  g.kb_begin = kb_begin; g.nkb = nkb; g.accum = accum;
  const int64_t total = (int64_t)N * g.m_tiles * g.n_tiles;
  const int grid = (int)std::min<int64_t>(total, D.sm_count);
  k_oz_mma<1><<<grid, OZ_THREADS, OZ_SMEM, st>>>(tmA, tmB, g);
  count_launch();
  return cudaGetLastError();
}

static const crt::Plan &crt_plan(int N)
{
  static crt::Plan plans[crt::NM + 1];
  static bool have[crt::NM + 1] = {false};
  if (!have[N]) { crt::host::build_plan(N, plans[N]); have[N] = true; }
  return plans[N];
}
/* planes: plane i holds rows x Kp bytes at planes + i * pstride (pstride = 0: packed, rows * Kp) */
void launch_crt_residues(const q128 *X, int64_t rows, int64_t K, int64_t sr, int64_t sk, const int *emax, int W, int N, int64_t Kp, int8_t *planes,
                         cudaStream_t st, int64_t pstride)
{
  if (pstride <= 0) pstride = rows * Kp;
  const int nw = std::min(crt::NWMAX, std::max(1, (W + 31) / 32));
  const bool direct = (sk == 1 || sr != 1);
  const int64_t threads = rows * (Kp / 4);
  const unsigned g1 = (unsigned)((threads + 255) / 256);
  const dim3 g2((unsigned)((rows + 15) / 16), (unsigned)(Kp / 32));
  switch (nw) {
#define QB_CRT_CASE(NW)                                                                                     \
  case NW:                                                                                                  \
    if (direct) k_crt_residues<NW><<<g1, 256, 0, st>>>(X, rows, K, sr, sk, emax, W, N, Kp, planes, pstride);  \
    else k_crt_residues_t<NW><<<g2, 128, crt_t_smem(N), st>>>(X, rows, K, sr, sk, emax, W, N, Kp, planes, pstride); \
    break;
    QB_CRT_CASE(1) QB_CRT_CASE(2) QB_CRT_CASE(3) QB_CRT_CASE(4) QB_CRT_CASE(5) QB_CRT_CASE(6)
#undef QB_CRT_CASE
  }
  count_launch();
}

/* scan of `rows` rows of one operand on `st`: per-row statistics into emax / lmin / sp */
static void launch_scan(const q128 *X, int64_t rows, int64_t K, int64_t sr, int64_t sk, int *emax, int *lmin, int *sp, cudaStream_t st)
{
  cudaMemsetAsync(emax, 0, (size_t)rows * 4, st); cudaMemsetAsync(lmin, 0x7f, (size_t)rows * 4, st); cudaMemsetAsync(sp, 0, (size_t)rows * 4, st);
  const int64_t nch = (K + OZ_SCAN_CH - 1) / OZ_SCAN_CH;
  const bool warp_mode = (sk == 1 || sr != 1);
  const int64_t threads = warp_mode ? rows * nch * 32 : rows * nch;
  k_oz_scan<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(X, rows, K, sr, sk, emax, lmin, sp);
  count_launch();
}
/* column statistics of op(B) for a caller that streams B in panels (qb_gemm_colstats_dev): stats = [emax[n] | lmin[n] | sp[n]] */
cudaError_t launch_colstats(const q128 *B, int64_t n, int64_t k, int64_t sbj, int64_t sbl, int *stats, cudaStream_t st)
{
  launch_scan(B, n, k, sbj, sbl, stats, stats + n, stats + 2 * n, st);
  return cudaGetLastError();
}

/* ---- tuning knobs (qb_set_tensor_window / qb_set_tensor_unit) ---- */
static int g_window = 144;            /* bits per window the planner grants when the spans do not fit (2 x 144 + log2 k + 1 -> 42 moduli at k = 8192) */
static int64_t g_unit_rows = 2048, g_unit_cols = 4096;
static int64_t g_ramp_rows = 0, g_ramp_cols = 0;   /* short first pass / panel (0 = like the others) */
static size_t g_ws_limit = 0;                       /* bytes the workspace may take (0 = 85 % of the free device memory) */
void oz_set_ws_limit(size_t bytes) { g_ws_limit = bytes; }
size_t oz_get_ws_limit() { return g_ws_limit; }
void oz_set_ramp(int64_t rows, int64_t cols) { g_ramp_rows = rows < 0 ? 0 : rows; g_ramp_cols = cols < 0 ? 0 : cols; }
void oz_get_ramp(int64_t *rows, int64_t *cols) { *rows = g_ramp_rows; *cols = g_ramp_cols; }
void oz_set_window(int bits) { g_window = bits < 120 ? 120 : (bits > crt::WMAX ? crt::WMAX : bits); }
int oz_get_window() { return g_window; }
void oz_set_unit(int64_t rows, int64_t cols)
{
  g_unit_rows = rows <= 0 ? 2048 : std::max<int64_t>(OZ_BM, rup(rows, OZ_BM));
  g_unit_cols = cols <= 0 ? 4096 : std::max<int64_t>(OZ_BN, rup(cols, OZ_BN));
}
void oz_get_unit(int64_t *rows, int64_t *cols) { *rows = g_unit_rows; *cols = g_unit_cols; }
/* legacy knobs of the row-pass partition (kept for the tests of the partition helper) */
static int g_oz_pass_shape = 0;
void oz_set_pass_shape(int v) { g_oz_pass_shape = v == 1 ? 1 : 0; }
int oz_get_pass_shape() { return g_oz_pass_shape; }
/* Rows of each pass: the sizes sum to m, every size but the last is a multiple of OZ_BM, none exceeds `cap`.  shape 0: equal
 * passes of `cap` rows.  shape 1: a short first pass (its residues cannot overlap a tensor pass) and a short last pass (neither
 * can its fold, nor its peer stores), equal passes in between. */
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

/* The whole fast-mode GEMM.  *used = 0 means the planner declined (no TMA entry point, or not even one unit fits the
 * workspace budget) and nothing was written: the caller runs the integer-limb kernel. */
cudaError_t launch_gemm_ozaki(const GemmArgs &a, cudaStream_t st, int *used, const OzHooks &h)
{
  *used = 0;
  { const int64_t keep = g_last_stats.ws_bytes; g_last_stats = OzStats(); g_last_stats.ws_bytes = keep; }
  if (!get_encode()) return cudaSuccess;
  OzDev *Dp;
  cudaError_t e = oz_dev(&Dp);
  if (e != cudaSuccess) return e;
  OzDev &D = *Dp;
  D.tev_used = 0;
  D.stats_pending = false;
  const int64_t m = a.m, n = a.n, k = a.k;
  const int64_t Kp = rup(k, OZ_BK);
#define QB_TRY(call) do { const cudaError_t e_ = (call); if (e_ != cudaSuccess) return e_; } while (0)
  cudaEvent_t evIn, evScanA, evScanB, evDone;
  QB_TRY(oz_event(D, EV_IN, &evIn)); QB_TRY(oz_event(D, EV_SCAN_A, &evScanA)); QB_TRY(oz_event(D, EV_SCAN_B, &evScanB)); QB_TRY(oz_event(D, EV_DONE, &evDone));
  /* ---- scan (A on one stream, B on another) + plan; ONE host sync, for the 3-int plan ---- */
  const size_t meta_ints = (size_t)(3 * m + 3 * n + 32);
  QB_TRY(oz_grow(&D.meta, &D.meta_bytes, meta_ints * 4));
  int *meta = (int *)D.meta;
  int *statA = meta, *statB = meta + 3 * m, *misc = meta + 3 * m + 3 * n;
  const int *emaxA = statA, *lminA = statA + m, *spA = statA + 2 * m, *emaxB = statB, *lminB = statB + n, *spB = statB + 2 * n;
  int *plan = misc, *total = misc + 8, *counter = misc + 9;
  QB_TRY(cudaMemsetAsync(misc, 0, 32 * 4, st));
  QB_TRY(cudaEventRecord(evIn, st));
  QB_TRY(cudaStreamWaitEvent(D.sA, evIn, 0)); QB_TRY(cudaStreamWaitEvent(D.sB, evIn, 0));
  /* streamed rows (h.rows_in): A is not on the device yet; its rows are scanned pass by pass right before their residues, the
   * window of A is the planner's budget and every element is tested (check) */
  if (h.rows_in) { QB_TRY(cudaMemsetAsync(statA, 0, (size_t)m * 4, D.sA)); QB_TRY(cudaMemsetAsync(statA + m, 0x7f, (size_t)m * 4, D.sA)); QB_TRY(cudaMemsetAsync(statA + 2 * m, 0, (size_t)m * 4, D.sA)); }
  else launch_scan(a.A, m, k, a.sai, a.sal, statA, statA + m, statA + 2 * m, D.sA);
  QB_TRY(cudaEventRecord(evScanA, D.sA));
  if (h.bstats) QB_TRY(cudaMemcpyAsync(statB, h.bstats, (size_t)n * 12, cudaMemcpyDeviceToDevice, D.sB));
  else launch_scan(a.B, n, k, a.sbj, a.sbl, statB, statB + n, statB + 2 * n, D.sB);
  QB_TRY(cudaEventRecord(evScanB, D.sB));
  QB_TRY(cudaStreamWaitEvent(st, evScanA, 0)); QB_TRY(cudaStreamWaitEvent(st, evScanB, 0));
  k_oz_plan<<<1, 1024, 0, st>>>(emaxA, lminA, spA, m, emaxB, lminB, spB, n, plan);
  count_launch();
  QB_TRY(cudaMemcpyAsync(D.h_plan, plan, 16, cudaMemcpyDeviceToHost, st));
  QB_TRY(cudaStreamSynchronize(st));
  /* streamed rows: the spans of A are not known yet; assume the widest, so that the planner gives A whatever the budget leaves
   * after B - never less than the resident call gets (same bits whenever that one is exact) */
  const int WA_nat = h.rows_in ? crt::WMAX : D.h_plan[0], WB_nat = h.bplanes_N ? h.bplanes_W : D.h_plan[1], fl = D.h_plan[2];
  crt::host::Windows win;
  if (!crt::host::plan_windows(WA_nat, WB_nat, k, g_window, win)) return cudaSuccess;
  const int WA = win.WA, WB = win.WB, N = win.N;
  const int check = (win.truncA || win.truncB || fl || h.rows_in) ? 1 : 0;
  /* B handed over as residue planes (computed elsewhere with window bplanes_W and bplanes_N moduli, e.g. cooperatively by the ranks of
   * a row-sharded product): this call must arrive at the same window, need no more moduli, and have no element to fix up (the fix-up
   * reads the original B) */
  if (h.bplanes_N && (!h.bp || WB != h.bplanes_W || N > h.bplanes_N || check)) return cudaErrorNotSupported;
  const crt::Plan &pl = crt_plan(N);
  /* ---- units: row passes of `ur` rows x column panels of `uc` columns ----
   * Loop order 0 (default): panels outer, passes inner — the residue planes of the A rows stay resident, the B panels are double
   * buffered (what a device-resident or a panel-streamed B wants).  Order 1: passes outer — the B planes stay resident and the A
   * passes are double buffered, so the rows of C complete pass by pass (what the all-host path wants: rows stream in and out). */
  int64_t ur = std::min(rup(m, OZ_BM), g_unit_rows), uc = std::min(rup(n, OZ_BN), h.bp ? rup(h.bp_cols, OZ_BN) : g_unit_cols);
  if (h.cb && h.min_passes > 1) ur = std::min(ur, std::max<int64_t>(OZ_BM, rup((m + h.min_passes - 1) / h.min_passes, OZ_BM)));
  /* the N residue planes of a unit are padded apart: with power-of-two plane sizes the N words a reconstruction thread gathers (one
   * per plane) would all fall into the same DRAM channel / L2 slice */
  auto rplane_of = [](int64_t Mp_, int64_t Np_) -> int64_t { return Mp_ * Np_ + 4352; };
  auto planes_bytes = [&](int64_t rows) -> size_t { return (size_t)rup((int64_t)N * rows * Kp, 1024); };
  auto rest_bytes = [&](int64_t ur_, int64_t uc_) -> size_t {   /* 2 residue buffers, flags + list */
    size_t b = 2 * (size_t)rup((int64_t)N * rplane_of(rup(ur_, OZ_BM), rup(uc_, OZ_BN)), 1024);
    if (check) b += (size_t)rup(rup(ur_, OZ_BM) * rup(uc_, OZ_BN), 1024) + (size_t)std::max<int64_t>(4096, ur_ * uc_ / 64) * sizeof(int2);
    return b;
  };
  auto need_bytes = [&](int64_t ur_, int64_t uc_, int64_t nA_, int64_t nB_) -> size_t { return nA_ * planes_bytes(ur_) + nB_ * planes_bytes(uc_) + rest_bytes(ur_, uc_); };
  auto npass_of = [&](int64_t ur_) { return (m + ur_ - 1) / ur_; };
  auto npanel_of = [&](int64_t uc_) { return (n + uc_ - 1) / uc_; };
  auto min_need = [&](int64_t ur_, int64_t uc_) { return need_bytes(ur_, uc_, std::min<int64_t>(npass_of(ur_) + 1, 2), std::min<int64_t>(npanel_of(uc_) + 1, 2)); };
  /* what fits: anything inside the workspace the library already holds, else 85 % of the free device memory on top of it (queried only
   * when the workspace would have to grow: cudaMemGetInfo is a driver call that the steady state does not need) */
  size_t budget = 0; bool have_budget = false;
  auto fits = [&](size_t need) -> bool {
    if (g_ws_limit) return need <= g_ws_limit;     /* qb_set_tensor_workspace_limit: an explicit cap (tests of the block loops, shared GPUs) */
    if (need <= D.bytes) return true;
    if (!have_budget) { size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); budget = (size_t)((double)fr * 0.85) + D.bytes; have_budget = true; }
    return need <= budget;
  };
  while (!fits(min_need(ur, uc)) && !h.bp && uc > OZ_BN) uc = rup(uc / 2, OZ_BN);
  while (!fits(min_need(ur, uc)) && ur > OZ_BM) ur = rup(ur / 2, OZ_BM);
  if (!fits(min_need(ur, uc))) return cudaSuccess;
  /* pass / panel boundaries.  The first pass and the first panel are short (g_ramp_rows x g_ramp_cols): their residues are the only
   * ones nothing can hide, and a short first unit lets the tensor kernel start early; streamed panels keep the caller's width. */
  std::vector<int64_t> prow, pcol;
  {
    const int64_t fr = (g_ramp_rows > 0 && !h.rows_in) ? std::min(ur, rup(g_ramp_rows, OZ_BM)) : ur;
    const int64_t fc = (g_ramp_cols > 0 && !h.bp) ? std::min(uc, rup(g_ramp_cols, OZ_BN)) : uc;
    for (int64_t r = 0; r < m; r += (r == 0 && m > fr + OZ_BM ? fr : ur)) prow.push_back(r);
    prow.push_back(m);
    for (int64_t c = 0; c < n; c += (c == 0 && n > fc + OZ_BN ? fc : uc)) pcol.push_back(c);
    pcol.push_back(n);
  }
  const int64_t nP = (int64_t)prow.size() - 1, nJ = (int64_t)pcol.size() - 1;
  const int order = h.order == 1 ? 1 : 0;
  int64_t nA = std::min<int64_t>(nP, 2), nB = std::min<int64_t>(nJ, 2);
  if (order == 0) { while (nA < nP && fits(need_bytes(ur, uc, nA + 1, nB))) ++nA; }
  else { while (nB < nJ && fits(need_bytes(ur, uc, nA, nB + 1))) ++nB; }
  if (h.bp && nA < nP) return cudaSuccess;                   /* a streamed B passes once */
  const int64_t Mu = rup(ur, OZ_BM), Nu = rup(uc, OZ_BN);
  const size_t pa_bytes = planes_bytes(ur), pb_bytes = planes_bytes(uc), r_bytes = (size_t)rup((int64_t)N * rplane_of(Mu, Nu), 1024);
  const int list_cap = (int)std::max<int64_t>(4096, ur * uc / 64);
  const size_t flag_bytes = check ? (size_t)rup(Mu * Nu, 1024) : 0, list_bytes = check ? (size_t)list_cap * sizeof(int2) : 0;
  QB_TRY(oz_grow(&D.buf, &D.bytes, nA * pa_bytes + nB * pb_bytes + 2 * r_bytes + flag_bytes + list_bytes));
  int8_t *pA0 = (int8_t *)D.buf, *pB0 = pA0 + nA * pa_bytes;
  uint8_t *R[2] = {(uint8_t *)(pB0 + nB * pb_bytes), (uint8_t *)(pB0 + nB * pb_bytes + r_bytes)};
  uint8_t *flagp = R[1] + r_bytes;
  int2 *list = (int2 *)(flagp + flag_bytes);
  /* K chunks of the int32 accumulation */
  const int64_t nkb_total = Kp / OZ_BK;
  const int nkc = (int)((Kp + OZ_KCHUNK - 1) / OZ_KCHUNK);
  const int64_t kcb = (nkb_total + nkc - 1) / nkc;
  const int simple = (a.alpha.hi == 0x3fff000000000000ULL && a.alpha.lo == 0 && (a.beta.hi & 0x7fffffffffffffffULL) == 0 && a.beta.lo == 0) ? 1 : 0;
  /* ---- the pipeline: tensor kernel of unit u (sM) || residues for the coming units (sA, sB) || reconstruction of unit u-1 (sF) ---- */
  cudaEvent_t evMma[2], evFold[2];
  for (int b = 0; b < 2; ++b) { QB_TRY(oz_event(D, EV_MMA0 + b, &evMma[b])); QB_TRY(oz_event(D, EV_FOLD0 + b, &evFold[b])); }
  struct Slot { int64_t who = -1; cudaEvent_t ready = nullptr, read = nullptr; bool read_rec = false; const q128 *panel = nullptr; int64_t ld = 0; };
  std::vector<Slot> slotA((size_t)nA), slotB((size_t)nB);
  for (int64_t i = 0; i < nA + nB; ++i) {
    Slot &sl = i < nA ? slotA[(size_t)i] : slotB[(size_t)(i - nA)];
    QB_TRY(oz_event(D, EV_SLOT0 + 2 * (int)i, &sl.ready)); QB_TRY(oz_event(D, EV_SLOT0 + 2 * (int)i + 1, &sl.read));
  }
  std::vector<int> rows_in_done((size_t)nP, 0), panels_left((size_t)nP, (int)nJ);
  int64_t unit = 0;
  int passes_done = 0;
  auto run_unit = [&](int64_t p, int64_t j) -> cudaError_t {
    const int64_t r0 = prow[(size_t)p], mr = prow[(size_t)p + 1] - r0, c0 = pcol[(size_t)j], w = pcol[(size_t)j + 1] - c0;
    const int64_t Mp = rup(mr, OZ_BM), Np = rup(w, OZ_BN);
    Slot &sa = slotA[(size_t)(p % nA)], &sb = slotB[(size_t)(j % nB)];
    int8_t *pAp = pA0 + (size_t)(p % nA) * pa_bytes, *pBj = pB0 + (size_t)(j % nB) * pb_bytes;
    if (h.rows_in && !rows_in_done[(size_t)p]) {   /* streamed rows: the caller orders the arrival of A's rows on sA, of C's on sF */
      if (h.rows_in(r0, mr, (void *)D.sA, (void *)D.sF, h.rows_user) != 0) return cudaErrorInvalidValue;
      rows_in_done[(size_t)p] = 1;
    }
    if (sa.who != p) {   /* residues of this pass's A rows; the slot was last read by the tensor kernel that recorded sa.read */
      if (sa.read_rec) QB_TRY(cudaStreamWaitEvent(D.sA, sa.read, 0));
      if (h.rows_in && rows_in_done[(size_t)p] == 1) { launch_scan(a.A + r0 * a.sai, mr, k, a.sai, a.sal, statA + r0, statA + m + r0, statA + 2 * m + r0, D.sA); rows_in_done[(size_t)p] = 2; }
      launch_crt_residues(a.A + r0 * a.sai, mr, k, a.sai, a.sal, emaxA + r0, WA, N, Kp, pAp, D.sA, 0);
      QB_TRY(cudaEventRecord(sa.ready, D.sA));
      QB_TRY(cudaStreamWaitEvent(D.sM, sa.ready, 0));
      sa.who = p;
    }
    const q128 *Bp = a.B + c0 * a.sbj;
    int64_t sbl = a.sbl, sbj = a.sbj;
    if (sb.who != j) {
      if (sb.read_rec) QB_TRY(cudaStreamWaitEvent(D.sB, sb.read, 0));
      if (h.bp) {   /* streamed B: the caller hands over the panel and orders its arrival on sB */
        const void *ptr = nullptr; int64_t ldp = 0;
        if (h.bp(c0, w, (void *)D.sB, &ptr, &ldp, h.bp_user) != 0 || !ptr) return cudaErrorInvalidValue;
        sb.panel = (const q128 *)ptr; sb.ld = ldp;
      }
      if (h.bp) { Bp = sb.panel; if (a.sbj == 1) sbl = sb.ld; else sbj = sb.ld; }
      if (!h.bplanes_N) launch_crt_residues(Bp, w, k, sbj, sbl, emaxB + c0, WB, N, Kp, pBj, D.sB, 0);
      QB_TRY(cudaEventRecord(sb.ready, D.sB));
      QB_TRY(cudaStreamWaitEvent(D.sM, sb.ready, 0));
      sb.who = j;
    } else if (h.bp) { Bp = sb.panel; if (a.sbj == 1) sbl = sb.ld; else sbj = sb.ld; }
    const int rb = (int)(unit & 1);
    if (unit >= 2) QB_TRY(cudaStreamWaitEvent(D.sM, evFold[rb], 0));   /* R[rb] drained by the reconstruction of unit - 2 */
    for (int c = 0; c < nkc; ++c) {
      const int kb0 = (int)(c * kcb), nkb = (int)std::min<int64_t>(kcb, nkb_total - kb0);
      oz_ev_record(D, 0, D.sM);
      const cudaError_t e2 = h.bplanes_N ? launch_crt_mma(D, pAp, (const int8_t *)sb.panel, N, mr, w, Kp, kb0, nkb, c > 0, R[rb], Mp, Np, rplane_of(Mp, Np), D.sM, sb.ld)
                                         : launch_crt_mma(D, pAp, pBj, N, mr, w, Kp, kb0, nkb, c > 0, R[rb], Mp, Np, rplane_of(Mp, Np), D.sM);
      oz_ev_record(D, 1, D.sM);
      if (e2 != cudaSuccess) return e2;
    }
    QB_TRY(cudaEventRecord(evMma[rb], D.sM));
    if (nA < nP) { QB_TRY(cudaEventRecord(sa.read, D.sM)); sa.read_rec = true; }
    if (nB < nJ) { QB_TRY(cudaEventRecord(sb.read, D.sM)); sb.read_rec = true; }
    /* reconstruction + epilogue (+ fix-up of the elements it rejects) */
    QB_TRY(cudaStreamWaitEvent(D.sF, evMma[rb], 0));
    CrtFoldArgs f;
    f.R = R[rb]; f.Mp = Mp; f.Np = Np; f.rplane = rplane_of(Mp, Np); f.m = mr; f.n = w; f.row0 = r0; f.col0 = c0;
    f.emaxA = emaxA; f.lminA = lminA; f.spA = spA; f.emaxB = emaxB; f.lminB = lminB; f.spB = spB; f.WA = WA; f.WB = WB; f.k = k;
    f.alpha = a.alpha; f.beta = a.beta; f.C = a.C; f.sci = a.sci; f.scj = a.scj;
    f.simple = simple;
    f.n4p = rup((w + 3) / 4, 32);
    f.check = check; f.list = list; f.list_cap = list_cap; f.counter = counter; f.flag = flagp;
    f.npeer = a.npeer;
    for (int q = 0; q < QB_MAX_PEERS; ++q) f.peer[q] = q < a.npeer ? a.peerC[q] : nullptr;
    if (check) { QB_TRY(cudaMemsetAsync(counter, 0, 4, D.sF)); QB_TRY(cudaMemsetAsync(flagp, 0, (size_t)(Mp * Np), D.sF)); }
    launch_crt_fold(f, pl, D.sF);
    QB_TRY(cudaGetLastError());
    if (check) {
      CrtFixArgs x;
      x.list = list; x.counter = counter; x.list_cap = list_cap; x.flag = flagp;
      x.um = mr; x.un = w; x.uNp = Np; x.row0 = r0; x.col0 = c0;
      x.A = a.A; x.sai = a.sai; x.sal = a.sal; x.k = k;
      x.B = Bp - c0 * sbj; x.sbl = sbl; x.sbj = sbj;   /* column gj of op(B) lives at Bp + (gj - c0) * sbj (the panel as handed over, or B itself) */
      x.alpha = a.alpha; x.beta = a.beta; x.C = a.C; x.sci = a.sci; x.scj = a.scj; x.total = total;
      x.npeer = a.npeer;
      for (int q = 0; q < QB_MAX_PEERS; ++q) x.peer[q] = q < a.npeer ? a.peerC[q] : nullptr;
      k_crt_fixup<<<D.sm_count * 4, 128, 0, D.sF>>>(x);
      count_launch();
      QB_TRY(cudaGetLastError());
    }
    QB_TRY(cudaEventRecord(evFold[rb], D.sF));
    if (--panels_left[(size_t)p] == 0) {   /* the rows of this pass are complete */
      ++passes_done;
      if (h.cb) { QB_TRY(cudaStreamWaitEvent(st, evFold[rb], 0)); h.cb(r0, mr, h.cb_user); }
    }
    ++unit;
    return cudaSuccess;
  };
  if (order == 0) {
    for (int64_t pb = 0; pb < nP; pb += nA)                 /* blocks of resident A passes (one block unless the workspace is short) */
      for (int64_t j = 0; j < nJ; ++j)
        for (int64_t p = pb; p < std::min(nP, pb + nA); ++p) QB_TRY(run_unit(p, j));
  } else {
    for (int64_t jb = 0; jb < nJ; jb += nB)
      for (int64_t p = 0; p < nP; ++p)
        for (int64_t j = jb; j < std::min(nJ, jb + nB); ++j) QB_TRY(run_unit(p, j));
  }
  if (check) { QB_TRY(cudaMemcpyAsync(D.h_plan + 8, total, 4, cudaMemcpyDeviceToHost, D.sF)); D.stats_pending = true; }
  QB_TRY(cudaEventRecord(evDone, D.sF));
  QB_TRY(cudaStreamWaitEvent(st, evDone, 0));   /* sF is the last stage of every unit: the caller's stream sees all of C */
  g_last_stats.SA = (WA + 7) / 8; g_last_stats.SB = (WB + 7) / 8; g_last_stats.ndiag = N; g_last_stats.nchunks = nkc;
  g_last_stats.row_passes = passes_done; g_last_stats.panels = (int)nJ; g_last_stats.units = unit;
  g_last_stats.pairs = N; g_last_stats.keep = 0; g_last_stats.flagged = 0; g_last_stats.redo_passes = 0;
  g_last_stats.ws_bytes = (int64_t)D.bytes; g_last_stats.Kp = Kp;
  g_last_stats.scheme = 1; g_last_stats.WA = WA; g_last_stats.WB = WB; g_last_stats.WA_nat = WA_nat; g_last_stats.WB_nat = WB_nat;
  g_last_stats.truncated = (win.truncA ? 1 : 0) | (win.truncB ? 2 : 0) | (fl ? 4 : 0);
  g_last_stats.peer_written = a.npeer;
  *used = 1;
  return cudaSuccess;
#undef QB_TRY
}

} // namespace qb

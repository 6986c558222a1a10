/*
 * qb_level2.cu — binary128 GEMV, A streamed once from HBM at 16 B/element.
 *
 * Replaces QuadBLAS::gemv (/root/reference/include/quadblas/algorithms/level2.hpp:85-99):
 *   row-major  (gemv_row_major :15-50):  y_i = fma(alpha, S_i, mul(beta, y_i)),
 *       S_i = dot_kernel_vectorized(row_i, x, n) when incx == 1 (level1.hpp:14-35: even-index
 *       chain + odd-index chain from +0, add(lane0, lane1), odd tail folded with one more fma),
 *       else one ascending chain (:39-45).
 *   col-major  (gemv_col_major :53-82):  y_i = mul(beta, y_i); for j ascending:
 *       y_i = fma(A[j*lda+i], mul(alpha, x_j), y_i).
 * Both orders are thread-count independent in the reference, so they are reproduced exactly.
 *
 * row-major kernel: a CTA owns ROWS rows; A tiles are loaded with fully coalesced 128-bit loads
 * (a warp reads 32 consecutive quads of one row) into padded shared memory and each chain thread
 * then walks its own row.  col-major kernel: thread per row, consecutive threads read consecutive
 * quads of a column, so the loads are coalesced without staging.
 *
 * FAST mode (k_gemv_row_wide / k_gemv_col_wide): S_i = sum_j A_ij x_j is accumulated in the unrounded
 * 192-bit window of qwide.cuh (~95 integer instructions per element instead of ~280 for the rounded
 * two-lane chain) and rounded once, then y_i = fma(alpha, S_i, mul(beta, y_i)) for both layouts.
 * Contract: |y^_i - y_i| <= gamma_n (|alpha| |A||x| + |beta y|)_i; for fixed shapes
 * the merge tree is fixed, so results are reproducible.
 */
#include "qb_internal.h"
#include "q128_chain.cuh"
#include "qwide.cuh"

namespace qb {

__device__ __forceinline__ q128 ldg128(const q128 *p)
{
  uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  q128 r;
  r.lo = ((uint64_t)v.y << 32) | v.x;
  r.hi = ((uint64_t)v.w << 32) | v.z;
  return r;
}

/* epilogue y = fma(alpha, s, mul(beta, y)) (level2.hpp:48) */
__device__ __noinline__ q128 gemv_epilogue(q128 alpha, q128 s, q128 beta, q128 y) { return q_fma(alpha, s, q_mul(beta, y)); }
__device__ __noinline__ q128 dev_add(q128 a, q128 b) { return q_add(a, b); }
__device__ __noinline__ q128 dev_mul(q128 a, q128 b) { return q_mul(a, b); }

/* LANES = 2: the reference's two-lane dot kernel (incx == 1); LANES = 1: single chain (incx != 1) */
template <int LANES, int ROWS, int TW>
__global__ void __launch_bounds__(ROWS * LANES)
k_gemv_row(GemvArgs g)
{
  constexpr int NT = ROWS * LANES;
  constexpr int PAD = 2;
  __shared__ q128 sA[ROWS][TW + PAD];
  __shared__ q128 sx[TW];
  const int tid = threadIdx.x;
  const int lane = tid % LANES;         /* which of the two interleaved chains */
  const int r = tid / LANES;            /* row inside the CTA */
  const int64_t row0 = (int64_t)blockIdx.x * ROWS;
  const int64_t nchain = (LANES == 2) ? (g.n / 2) * 2 : g.n; /* elements covered by the lane chains */

  qacc acc = qacc_zero();
  for (int64_t c0 = 0; c0 < nchain; c0 += TW) {
    /* coalesced tile load: consecutive threads -> consecutive columns of one row */
    for (int idx = tid; idx < ROWS * TW; idx += NT) {
      const int cc = idx % TW, rr = idx / TW;
      const int64_t gr = row0 + rr, gc = c0 + cc;
      sA[rr][cc] = (gr < g.m && gc < g.n) ? ldg128(g.A + gr * g.lda + gc) : q_one();
    }
    for (int idx = tid; idx < TW; idx += NT) {
      const int64_t gc = c0 + idx;
      sx[idx] = (gc < g.n) ? g.x[gc * g.incx] : q_one();
    }
    __syncthreads();
    const int lim = (int)((nchain - c0) < TW ? (nchain - c0) : TW);
    for (int c = lane; c < lim; c += LANES) qacc_fma(acc, qop_load(sA[r][c]), qop_load(sx[c]));
    __syncthreads();
  }

  q128 s = qacc_pack(acc);
  if (LANES == 2) {
    /* horizontal_sum = lane0 + lane1 (quad_vector.hpp:141-150), then the odd tail (level1.hpp:29-32) */
    q128 o;
    o.lo = __shfl_down_sync(0xffffffffu, s.lo, 1);
    o.hi = __shfl_down_sync(0xffffffffu, s.hi, 1);
    if (lane == 0) {
      s = dev_add(s, o);
      const int64_t gr = row0 + r;
      if ((g.n & 1) && gr < g.m) s = q_fma_slow_packed(ldg128(g.A + gr * g.lda + (g.n - 1)), g.x[(g.n - 1) * g.incx], s);
    }
  }
  const int64_t gr = row0 + r;
  if (lane == 0 && gr < g.m) {
    q128 *yp = g.y + gr * g.incy;
    *yp = gemv_epilogue(g.alpha, s, g.beta, *yp);
  }
}

template <int ROWS, int TW>
__global__ void __launch_bounds__(ROWS)
k_gemv_col(GemvArgs g)
{
  __shared__ q128 sc[TW];               /* c_j = mul(alpha, x_j) (level2.hpp:71) */
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * ROWS + tid;
  const bool live = i < g.m;
  qacc acc = qacc_from(live ? dev_mul(g.beta, g.y[i * g.incy]) : q_one()); /* level2.hpp:63-66 */
  for (int64_t j0 = 0; j0 < g.n; j0 += TW) {
    for (int idx = tid; idx < TW; idx += ROWS) {
      const int64_t j = j0 + idx;
      sc[idx] = (j < g.n) ? dev_mul(g.alpha, g.x[j * g.incx]) : q_one();
    }
    __syncthreads();
    const int lim = (int)((g.n - j0) < TW ? (g.n - j0) : TW);
    const q128 *col = g.A + j0 * g.lda + (live ? i : 0);
#pragma unroll 4
    for (int j = 0; j < lim; ++j) {
      const q128 a = ldg128(col + (int64_t)j * g.lda);
      qacc_fma(acc, qop_load(a), qop_load(sc[j]));
    }
    __syncthreads();
  }
  if (live) g.y[i * g.incy] = qacc_pack(acc);
}


/* ------------------------------------------------------------------ fast mode (window accumulator) */
__device__ __noinline__ qwide gv_merge(qwide a, qwide b) { qw_merge(a, b); return a; }
__device__ __forceinline__ qwide gv_shfl_down(const qwide &s, int off)
{
  qwide t;
  t.w0 = __shfl_down_sync(0xffffffffu, s.w0, off); t.w1 = __shfl_down_sync(0xffffffffu, s.w1, off);
  t.w2 = __shfl_down_sync(0xffffffffu, s.w2, off); t.w3 = __shfl_down_sync(0xffffffffu, s.w3, off);
  t.w4 = __shfl_down_sync(0xffffffffu, s.w4, off); t.w5 = __shfl_down_sync(0xffffffffu, s.w5, off);
  t.E = __shfl_down_sync(0xffffffffu, s.E, off);
  return t;
}
__device__ __forceinline__ void gv_store(uint32_t *dst, const qwide &s, uint32_t bad)
{
  reinterpret_cast<uint4 *>(dst)[0] = make_uint4(s.w0, s.w1, s.w2, s.w3);
  reinterpret_cast<uint4 *>(dst)[1] = make_uint4(s.w4, s.w5, (uint32_t)s.E, bad);
}
__device__ __forceinline__ qwide gv_load(const uint32_t *src, uint32_t &bad)
{
  const uint4 a = reinterpret_cast<const uint4 *>(src)[0], b = reinterpret_cast<const uint4 *>(src)[1];
  qwide s;
  s.w0 = a.x; s.w1 = a.y; s.w2 = a.z; s.w3 = a.w; s.w4 = b.x; s.w5 = b.y; s.E = (int32_t)b.z;
  bad |= b.w;
  return s;
}

/* Row-major: a CTA of NT threads owns R consecutive rows; thread t walks columns t, t+NT, ... of all
 * R rows (a warp reads 512 contiguous bytes of each row per step, x_j is unpacked once for R rows).
 * The R accumulate steps of one column are branch-free (qwa_fma) so that they interleave; a step that
 * qwa_fma declines (zero / subnormal / Inf / NaN operand, product above the anchor) is redone out of
 * line.  End of row: shuffle tree per warp, then thread r folds the NT/32 warp windows of row r in order. */
/* one column step of R rows: branch-free accumulate, then the declined steps out of line */
template <int R, int NT>
__device__ __forceinline__ void gv_row_step(qwacc (&acc)[R], uint32_t (&bad)[R], const q128 (&av)[R], const q128 &xv, uint32_t *col0)
{
  const qop X = qop_load_n(xv);
  bool rare = false, rr[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rr[r] = qwa_fma(acc[r], qop_load_n(av[r]), X, col0 + r * QWA_COL_WORDS * NT, NT);
    rare |= rr[r];
  }
  if (rare) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (rr[r]) qwa_fma_rare(acc[r], av[r], xv, bad[r]);
  }
}

template <int R, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_gemv_row_wide(GemvArgs g)
{
  constexpr int NW = NT / 32;
  static_assert(R * NW * 8 <= R * QWA_COL_WORDS * NT, "the reduction records reuse the scratch columns");
  __shared__ __align__(16) uint32_t scr[R * QWA_COL_WORDS * NT];
  uint32_t *sh = scr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * R;
  qwacc acc[R];
  uint32_t bad[R];
  const q128 *ap[R];                     /* running pointers: element (row r, column tid + k * NT) */
#pragma unroll
  for (int r = 0; r < R; ++r) {
    acc[r] = qwa_zero();
    bad[r] = 0;
    const int64_t gr = (row0 + r < g.m) ? row0 + r : g.m - 1;   /* clamp: duplicate work, never stored */
    ap[r] = g.A + gr * g.lda + tid;
    qwa_col_init(scr + r * QWA_COL_WORDS * NT + tid, NT);
  }
  const q128 *xp = g.x + (int64_t)tid * g.incx;
  const int64_t xstep = (int64_t)NT * g.incx;
  /* software pipeline over two register sets (no copies): the loads of step k + 1 are issued before
   * the arithmetic of step k */
  int left = (g.n > tid) ? (int)((g.n - tid + NT - 1) / NT) : 0;   /* column steps of this thread */
  q128 a0[R], a1[R], x0, x1;
  if (left > 0) {
    x0 = ldg128(xp);
#pragma unroll
    for (int r = 0; r < R; ++r) a0[r] = ldg128(ap[r]);
  }
  while (left > 0) {
    if (left > 1) {
      xp += xstep;
      x1 = ldg128(xp);
#pragma unroll
      for (int r = 0; r < R; ++r) { ap[r] += NT; a1[r] = ldg128(ap[r]); }
    }
    gv_row_step<R, NT>(acc, bad, a0, x0, scr + tid);
    if (--left == 0) break;
    if (left > 1) {
      xp += xstep;
      x0 = ldg128(xp);
#pragma unroll
      for (int r = 0; r < R; ++r) { ap[r] += NT; a0[r] = ldg128(ap[r]); }
    }
    gv_row_step<R, NT>(acc, bad, a1, x1, scr + tid);
    --left;
  }
  __syncthreads();                       /* the scratch columns become the reduction records */
#pragma unroll
  for (int r = 0; r < R; ++r) {
    qwide v = qwa_fold(acc[r]);
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) v = gv_merge(v, gv_shfl_down(v, off));
    const uint32_t b = __reduce_or_sync(0xffffffffu, bad[r]);
    if (lane == 0) gv_store(sh + (r * NW + warp) * 8, v, b);
  }
  __syncthreads();
  if (tid < R && row0 + tid < g.m) {
    uint32_t b = 0;
    qwide v = gv_load(sh + (tid * NW) * 8, b);
#pragma unroll 1
    for (int w = 1; w < NW; ++w) v = gv_merge(v, gv_load(sh + (tid * NW + w) * 8, b));
    q128 *yp = g.y + (row0 + tid) * g.incy;
    *yp = gemv_epilogue(g.alpha, qw_finish(v, b), g.beta, *yp);
  }
}

/* Col-major (and row-major transposed): thread per row, consecutive threads read consecutive quads
 * of a column; the column range is split over gridDim.y CTAs so that the grid fills the chip, each
 * writing its window to part[split][row]; k_gemv_col_fin folds the splits in order and applies the
 * epilogue. */
template <int ROWS, int TW>
__global__ void __launch_bounds__(ROWS)
k_gemv_col_wide(GemvArgs g, int64_t jchunk, uint32_t *part)
{
  constexpr int U = 4;                  /* columns in flight per thread, one scratch column each */
  static_assert(TW % U == 0, "tile width");
  __shared__ q128 sx[TW];
  __shared__ uint32_t scr[U * QWA_COL_WORDS * ROWS];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * ROWS + tid;
  const bool live = i < g.m;
  const int64_t jb = (int64_t)blockIdx.y * jchunk;
  const int64_t je = (jb + jchunk < g.n) ? jb + jchunk : g.n;
  qwacc acc = qwa_zero();
  uint32_t bad = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) qwa_col_init(scr + u * QWA_COL_WORDS * ROWS + tid, ROWS);
  for (int64_t j0 = jb; j0 < je; j0 += TW) {
    for (int idx = tid; idx < TW; idx += ROWS) {
      const int64_t j = j0 + idx;
      sx[idx] = (j < je) ? g.x[j * g.incx] : q_zero(0);
    }
    __syncthreads();
    const int lim = (int)((je - j0) < TW ? (je - j0) : TW);
    const q128 *col = g.A + j0 * g.lda + (live ? i : 0);
    int j = 0;
    for (; j + U <= lim; j += U) {
      q128 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = ldg128(col + (int64_t)(j + u) * g.lda);
      bool rare = false, rr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rr[u] = qwa_fma(acc, qop_load_n(a[u]), qop_load_n(sx[j + u]), scr + u * QWA_COL_WORDS * ROWS + tid, ROWS);
        rare |= rr[u];
      }
      if (rare) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (rr[u]) qwa_fma_rare(acc, a[u], sx[j + u], bad);
      }
    }
    for (; j < lim; ++j) {
      const q128 a = ldg128(col + (int64_t)j * g.lda);
      if (qwa_fma(acc, qop_load_n(a), qop_load_n(sx[j]), scr + tid, ROWS)) qwa_fma_rare(acc, a, sx[j], bad);
    }
    __syncthreads();
  }
  if (live) gv_store(part + ((int64_t)blockIdx.y * g.m + i) * 8, qwa_fold(acc), bad);
}

__global__ void k_gemv_col_fin(GemvArgs g, int splits, const uint32_t *part)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.m) return;
  uint32_t b = 0;
  qwide v = gv_load(part + i * 8, b);
#pragma unroll 1
  for (int s = 1; s < splits; ++s) v = gv_merge(v, gv_load(part + ((int64_t)s * g.m + i) * 8, b));
  q128 *yp = g.y + i * g.incy;
  *yp = gemv_epilogue(g.alpha, qw_finish(v, b), g.beta, *yp);
}

static constexpr int GV_COL_ROWS = 128;
static int gemv_col_splits(int64_t m, int64_t n)
{
  /* enough CTAs for ~4 per SM, at least 64 columns per split */
  const int64_t rb = (m + GV_COL_ROWS - 1) / GV_COL_ROWS;
  int64_t s = (148 * 4 + rb - 1) / rb;
  const int64_t smax = (n + 63) / 64;
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return (int)s;
}

/* 16-byte elements of device scratch the fast col-major path needs (32-byte window per split x row) */
int64_t gemv_work_elems(int64_t m, int64_t n, int col_major, int mode)
{
  if (mode == 0 || !col_major || fast_variant() == 0 || m <= 0 || n <= 0) return 0;
  return 2 * (int64_t)gemv_col_splits(m, n) * m;
}

cudaError_t launch_gemv(const GemvArgs &a, int mode, cudaStream_t st)
{
  if (a.m == 0 || a.n == 0) return cudaSuccess; /* level2.hpp:21,59: y untouched */
  if (mode != 0 && fast_variant() != 0) {
    if (!a.col_major) {
      /* R rows per CTA: 4 when that still gives >= 2 CTAs per SM, else fewer rows for more CTAs */
      /* R rows per CTA share the unpacked x_j; 2 x 128 measured best at m = 32768 (96 registers, 5 CTAs per SM);
       * fewer rows per CTA when m alone cannot fill the chip */
#define GV_LAUNCH(R_, NT_, MB_) k_gemv_row_wide<R_, NT_, MB_><<<(unsigned)((a.m + R_ - 1) / R_), NT_, 0, st>>>(a)
      if (a.m >= 2 * 148 * 5) GV_LAUNCH(2, 128, 5);
      else if (a.m >= 148 * 2) GV_LAUNCH(1, 128, 8);
      else GV_LAUNCH(1, 256, 4);
#undef GV_LAUNCH
      count_launch();
    } else {
      const int splits = gemv_col_splits(a.m, a.n);
      if (a.work == nullptr || a.work_elems < 2 * (int64_t)splits * a.m) return cudaErrorInvalidValue;
      const int64_t jchunk = (((a.n + splits - 1) / splits) + 31) / 32 * 32;
      const int gy = (int)((a.n + jchunk - 1) / jchunk);
      dim3 grid((unsigned)((a.m + GV_COL_ROWS - 1) / GV_COL_ROWS), (unsigned)gy);
      k_gemv_col_wide<GV_COL_ROWS, 32><<<grid, GV_COL_ROWS, 0, st>>>(a, jchunk, reinterpret_cast<uint32_t *>(a.work));
      k_gemv_col_fin<<<(unsigned)((a.m + 127) / 128), 128, 0, st>>>(a, gy, reinterpret_cast<const uint32_t *>(a.work));
      count_launch(2);
    }
    return cudaGetLastError();
  }
  if (!a.col_major) {
    if (a.incx == 1) {
      constexpr int ROWS = 64, TW = 32;
      k_gemv_row<2, ROWS, TW><<<(unsigned)((a.m + ROWS - 1) / ROWS), ROWS * 2, 0, st>>>(a);
    } else {
      constexpr int ROWS = 64, TW = 32;
      k_gemv_row<1, ROWS, TW><<<(unsigned)((a.m + ROWS - 1) / ROWS), ROWS, 0, st>>>(a);
    }
  } else {
    constexpr int ROWS = 128, TW = 32;
    k_gemv_col<ROWS, TW><<<(unsigned)((a.m + ROWS - 1) / ROWS), ROWS, 0, st>>>(a);
  }
  count_launch();
  return cudaGetLastError();
}

} // namespace qb

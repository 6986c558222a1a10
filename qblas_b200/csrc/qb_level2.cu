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
 */
#include "qb_internal.h"
#include "q128_chain.cuh"

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

cudaError_t launch_gemv(const GemvArgs &a, int mode, cudaStream_t st)
{
  (void)mode; /* the reference order is already fully parallel over rows; fast mode shares it for now */
  if (a.m == 0 || a.n == 0) return cudaSuccess; /* level2.hpp:21,59: y untouched */
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

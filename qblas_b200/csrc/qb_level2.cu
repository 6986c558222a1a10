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
#include <algorithm>
#include <cstdlib>
#include "qb_internal.h"
#include "q128_chain.cuh"
#include "qwide.cuh"
#include "qslice.cuh"
#include "qb_tc.cuh"

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

template <int R, int NT>
__device__ __forceinline__ void gv_row_wide_body(const GemvArgs &g, const int64_t row0, uint32_t *scr)
{
  constexpr int NW = NT / 32;
  static_assert(R * NW * 8 <= R * QWA_COL_WORDS * NT, "the reduction records reuse the scratch columns");
  uint32_t *sh = scr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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

template <int R, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_gemv_row_wide(GemvArgs g)
{
  __shared__ __align__(16) uint32_t scr[R * QWA_COL_WORDS * NT];
  gv_row_wide_body<R, NT>(g, (int64_t)blockIdx.x * R, scr);
}

/* the rows the sliced kernel below declined (rowflag != 0), one row per CTA pass */
template <int NT>
__global__ void __launch_bounds__(NT, 8)
k_gemv_row_wide_sel(GemvArgs g, const uint8_t *rowflag)
{
  __shared__ __align__(16) uint32_t scr[QWA_COL_WORDS * NT];
  for (int64_t row = blockIdx.x; row < g.m; row += gridDim.x) {
    if (rowflag[row] == 0) continue;                /* uniform over the CTA */
    __syncthreads();                                /* the scratch of the previous row is free */
    gv_row_wide_body<1, NT>(g, row, scr);
  }
}

/* ------------------------------------------------------------------ fast mode on the FP64 pipe (qslice.cuh), both layouts
 * k_gv_xscan: EX = largest exponent field of x, and whether x holds an Inf / NaN / nonzero subnormal (then every row goes to
 * the window kernel).  k_gv_xtab: per x_j a 200-byte table {6 zeros, +X_0..5, 6 zeros, -X_0..5, e(x_j)} — what qs_step reads at the
 * dynamic slice offset, ready to be copied.  k_gemv_f64: ONE THREAD PER ROW (no cross-thread merge, no per-thread staging of x):
 * a CTA owns a block of rows and a range of columns (gridDim.y splits); tiles of A arrive by TMA into a two-stage ring (col-major:
 * box {128 rows, 8 columns}, consecutive threads read consecutive quads; row-major: boxes {128 B, rows} with the 128-byte swizzle,
 * so that 32 threads reading the same column of 32 consecutive rows hit 32 different bank groups), the x tables of the tile's
 * columns by one bulk copy on the same mbarrier; elements past m or n read as zero (TMA fill, zero tables), so there is no edge
 * code.  Each thread runs qs_step on its row (about 72 instructions per element), keeps its 256-bit window in shared memory, and
 * ends with a record {window as qwide, anchor, Dmax, flags}.  k_gemv_f64_fin folds the records of a row in order, applies
 * qs_accept and either stores y_i = fma(alpha, S_i, mul(beta, y_i)) or flags the row for the window kernel
 * (k_gemv_row_wide_sel / k_gemv_col_wide). */
__global__ void k_gv_xscan(GemvArgs g, int32_t *hdr)
{
  int32_t emax = 0;
  uint32_t special = 0;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < g.n; j += (int64_t)gridDim.x * blockDim.x) {
    const qop o = qop_load(g.x[j * g.incx]);
    if (o.e == 0x7fff || (o.e == 0 && (o.m0 | o.m1 | o.m2 | o.m3) != 0u)) special = 1u;
    else emax = max(emax, o.e);
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  special = __reduce_or_sync(0xffffffffu, special);
  if ((threadIdx.x & 31) == 0) {
    if (emax) atomicMax(hdr, emax);
    if (special) atomicOr(reinterpret_cast<unsigned *>(hdr) + 1, 1u);
  }
}

constexpr int GVT_ROWS = 128;            /* rows per CTA = threads per CTA */
constexpr int GVT_COLS = 8;              /* columns per tile: 8 x 16 B = the 128-byte swizzle span */
constexpr int GVT_XDBL = QS_XCOL + 1;    /* doubles per column of the x table */
constexpr int GVT_GRP = 4;               /* elements stepped between two looks at the rare flags */
constexpr int GVT_TILE_BYTES = GVT_ROWS * GVT_COLS * 16;
constexpr int GVT_NB_ROW = 2;           /* column boxes per tile of the row-major kernel (see k_gemv_f64) */
constexpr int gvt_smem(int nb) { return 2 * GVT_TILE_BYTES + 2 * ((GVT_COLS * nb * GVT_XDBL * 8 + 127) / 128 * 128) + 4 * GVT_ROWS * 8 + 64; }

__global__ void k_gv_xtab(GemvArgs g, const int32_t *hdr, double *tab, int64_t npad)
{
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  qs_xrec r;
  if (j < g.n) qs_xrec_make(g.x[j * g.incx], hdr[0], r);
  else { for (int l = 0; l < QS_NS; ++l) r.X[l] = 0.0; r.ex = QS_EXNONE; }
  double *dst = tab + j * GVT_XDBL;
#pragma unroll
  for (int l = 0; l < QS_NS; ++l) { dst[l] = 0.0; dst[QS_NS + l] = r.X[l]; dst[2 * QS_NS + l] = 0.0; dst[3 * QS_NS + l] = -r.X[l]; }
  dst[QS_XCOL] = __hiloint2double(0, r.ex);
}

template <int NT>
__device__ __noinline__ void gv_flush(double c0, double c1, double c2, double c3, double c4, double c5, uint64_t *w)
{
  qs_flush(c0, c1, c2, c3, c4, c5, w, NT);
}

/* the element the hot form left out (it met only zeros there): qs_rare moves the anchor or flags the row, then the element is
 * stepped.  Through memory (t = the six columns, st = {anchor, flags}) so that the call does not pin the hot loop's registers. */
template <int NT>
__device__ __noinline__ void gv_rare_mem(double *t, int32_t *st, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint64_t *w, const double *col)
{
  qs_cols C; C.c0 = t[0]; C.c1 = t[1]; C.c2 = t[2]; C.c3 = t[3]; C.c4 = t[4]; C.c5 = t[5];
  qs_row S; S.anc = st[0]; S.dmax = 0;
  uint32_t flags = (uint32_t)st[1];
  const uint32_t e = (w3 >> 16) & 0x7fffu;
  /* an earlier element of the same group may have moved the anchor above this one */
  const uint32_t sh = ((uint32_t)(e - 1u) >= (uint32_t)S.anc) ? qs_rare(C, S, flags, e, w0, w1, w2, w3, w, NT) : min((uint32_t)S.anc - e, QS_SHMAX);
  if (sh < QS_SHMAX) qs_step(C, w0, w1, w2, w3, sh, col, 1);
  t[0] = C.c0; t[1] = C.c1; t[2] = C.c2; t[3] = C.c3; t[4] = C.c4; t[5] = C.c5;
  st[0] = S.anc; st[1] = (int32_t)flags;
}

/* record of one (row, column split): the window as a qwide + flags (gv_store), then anchor and Dmax */
constexpr int GVT_REC = 12;

/* NB = column boxes of 8 columns per tile: the CTA's 128 threads are RB = 128 / NB rows x NB boxes (a warp = 32 consecutive rows of
 * one box), a tile is RB rows x 8 NB columns.  Row-major uses NB = 2 or 4: a row then contributes 256 / 512 contiguous bytes per
 * tile instead of 128, which is what the DRAM pages want (measured: NB = 1 leaves row-major 10 % behind col-major, whose tile is a
 * run of 2 KB per column); every (row, box) is a partial of its own.  Col-major: NB = 1. */
template <bool COL, int NB>
__global__ void __launch_bounds__(GVT_ROWS, 5)
k_gemv_f64(const __grid_constant__ CUtensorMap tmA, GemvArgs g, const int32_t *hdr, const double *tab, int64_t jsplit, uint32_t *part)
{
  constexpr int RB = GVT_ROWS / NB;                  /* rows per CTA */
  constexpr int TC = GVT_COLS * NB;                  /* columns per tile */
  constexpr int XT_BYTES = TC * GVT_XDBL * 8;
  constexpr int XT_STRIDE = (XT_BYTES + 127) / 128 * 128;
  static_assert(!COL || NB == 1, "col-major tiles are one box");
  /* the kernel has no static shared memory, so the dynamic array starts the CTA's window and the 1024-byte alignment the swizzled
   * tiles need is the declared one (checked below); used directly so that every access stays an LDS / STS */
  extern __shared__ __align__(1024) unsigned char gvt_sm[];
  uint64_t *win = reinterpret_cast<uint64_t *>(gvt_sm + 2 * GVT_TILE_BYTES + 2 * XT_STRIDE);
  uint64_t *full = win + 4 * GVT_ROWS;
  const int tid = threadIdx.x;
  const int box = tid / RB, rr = tid % RB;
  const int64_t row0 = (int64_t)blockIdx.x * RB, row = row0 + rr;
  const int64_t j0 = (int64_t)blockIdx.y * jsplit, j1 = min(g.n, j0 + jsplit);
  const int ntiles = hdr[1] != 0 ? 0 : (int)((j1 - j0 + TC - 1) / TC);   /* Inf / NaN / subnormal in x: nothing here, every row flagged */
  const int32_t EX = hdr[0];
  auto issue = [&](int t, int s) {
    const int64_t j = j0 + (int64_t)t * TC;
    tc::mbar_expect_tx(&full[s], GVT_TILE_BYTES + XT_BYTES);
    if (COL) tc::tma_load_2d(gvt_sm + s * GVT_TILE_BYTES, &tmA, &full[s], (int)(row0 * 2), (int)j);
    else {
#pragma unroll
      for (int bb = 0; bb < NB; ++bb)
        tc::tma_load_2d(gvt_sm + s * GVT_TILE_BYTES + bb * (RB * 128), &tmA, &full[s], (int)((j + bb * GVT_COLS) * 4), (int)row0);
    }
    tc::bulk_load_1d(gvt_sm + 2 * GVT_TILE_BYTES + s * XT_STRIDE, tab + j * GVT_XDBL, XT_BYTES, &full[s]);
  };
  if (tid == 0) {
    if ((tc::smem_u32(gvt_sm) & 1023u) != 0u) __trap();
    tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
    tc::fence_barrier_init();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) win[k * GVT_ROWS + tid] = 0ull;
  __syncthreads();
  if (tid == 0) {
    if (ntiles > 0) issue(0, 0);
    if (ntiles > 1) issue(1, 1);
  }
  qs_cols C = qs_cols_zero();
  int32_t anc = QS_ANCMIN, dmax = QS_EXNONE;
  uint32_t flags = hdr[1] != 0 ? QS_FALLBACK : 0u;
  uint64_t *wn = win + tid;
  const uint32_t aoff = COL ? (uint32_t)tid * 16u : (uint32_t)(box * (RB * 128) + rr * 128);
  for (int t = 0; t < ntiles; ++t) {
    const int s = t & 1;
    tc::mbar_wait(&full[s], (uint32_t)(t >> 1) & 1u);
    const unsigned char *at = gvt_sm + s * GVT_TILE_BYTES + aoff;
    const double *xs = reinterpret_cast<const double *>(gvt_sm + 2 * GVT_TILE_BYTES + s * XT_STRIDE) + box * (GVT_COLS * GVT_XDBL);
    /* groups of GVT_GRP elements: the hot steps are branch-free so that their decode / convert / multiply chains interleave in one
     * instruction stream; an element the hot form cannot take (zero / subnormal / Inf / NaN, or larger than everything before it
     * in this row) meets only the six zeros there and is redone after its group, in order */
#pragma unroll
    for (int c0 = 0; c0 < GVT_COLS; c0 += GVT_GRP) {
      uint4 v[GVT_GRP];
      bool rare[GVT_GRP], any = false;
#pragma unroll
      for (int u = 0; u < GVT_GRP; ++u) {
        const int c = c0 + u;
        v[u] = *reinterpret_cast<const uint4 *>(COL ? at + c * (GVT_ROWS * 16) : at + ((c ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int u = 0; u < GVT_GRP; ++u) {
        const double *col = xs + (c0 + u) * GVT_XDBL;
        const uint32_t e = (v[u].w >> 16) & 0x7fffu;
        rare[u] = (uint32_t)(e - 1u) >= (uint32_t)anc;
        any |= rare[u];
        /* nsh = 21 - min(anc - e, QS_SHMAX) in one add-max; an element above the anchor (the difference wraps) and a zero
         * (anc >= QS_ANCMIN) both get QS_SHMAX and meet only zeros */
        const int32_t nsh = 21 - (int32_t)min((uint32_t)anc - e, QS_SHMAX);
        qs_step_n(C, v[u].x, v[u].y, v[u].z, v[u].w, nsh, col, 1);
        dmax = max(dmax, (int32_t)e + __double2loint(col[QS_XCOL]));
      }
      if (any) {
#pragma unroll
        for (int u = 0; u < GVT_GRP; ++u) {
          if (!rare[u]) continue;
          double tt[6] = {C.c0, C.c1, C.c2, C.c3, C.c4, C.c5};
          int32_t st[2] = {anc, (int32_t)flags};
          gv_rare_mem<GVT_ROWS>(tt, st, v[u].x, v[u].y, v[u].z, v[u].w, wn, xs + (c0 + u) * GVT_XDBL);
          C.c0 = tt[0]; C.c1 = tt[1]; C.c2 = tt[2]; C.c3 = tt[3]; C.c4 = tt[4]; C.c5 = tt[5];
          anc = st[0]; flags = (uint32_t)st[1];
        }
      }
    }
    if ((t & 7) == 7) {                   /* 64 elements: the columns are still exact */
      gv_flush<GVT_ROWS>(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, wn);
      C = qs_cols_zero();
    }
    __syncthreads();                      /* every thread is done with stage s (an mbarrier handshake that lets the warps drift was measured slower) */
    if (tid == 0 && t + 2 < ntiles) issue(t + 2, s);
  }
  gv_flush<GVT_ROWS>(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, wn);
  if (row < g.m) {
    const qwide v = qs_to_qwide(wn, GVT_ROWS, anc, EX);
    uint32_t *dst = part + (((int64_t)blockIdx.y * NB + box) * g.m + row) * GVT_REC;
    gv_store(dst, v, flags);
    dst[8] = (uint32_t)anc; dst[9] = (uint32_t)dmax;
  }
}

/* fold the column splits of a row in order; store the row or flag it */
__global__ void k_gemv_f64_fin(GemvArgs g, const int32_t *hdr, int splits, const uint32_t *part, uint8_t *rowflag)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.m) return;
  uint32_t b = 0;
  const uint32_t *src = part + i * GVT_REC;
  qwide v = gv_load(src, b);
  int32_t am = (int32_t)src[8], dm = (int32_t)src[9];
#pragma unroll 1
  for (int s = 1; s < splits; ++s) {
    src = part + ((int64_t)s * g.m + i) * GVT_REC;
    v = gv_merge(v, gv_load(src, b));
    am = max(am, (int32_t)src[8]); dm = max(dm, (int32_t)src[9]);
  }
  const bool ok = qs_accept(am, hdr[0], dm, b);
  rowflag[i] = ok ? 0 : 1;
  if (ok) {
    q128 *yp = g.y + i * g.incy;
    *yp = gemv_epilogue(g.alpha, qw_finish(v, 0u), g.beta, *yp);
  }
}

/* Col-major (and row-major transposed): thread per row, consecutive threads read consecutive quads
 * of a column; the column range is split over gridDim.y CTAs so that the grid fills the chip, each
 * writing its window to part[split][row]; k_gemv_col_fin folds the splits in order and applies the
 * epilogue. */
template <int ROWS, int TW>
__global__ void __launch_bounds__(ROWS)
k_gemv_col_wide(GemvArgs g, int64_t jchunk, uint32_t *part, const uint8_t *only)
{
  constexpr int U = 4;                  /* columns in flight per thread, one scratch column each */
  static_assert(TW % U == 0, "tile width");
  __shared__ q128 sx[TW];
  __shared__ uint32_t scr[U * QWA_COL_WORDS * ROWS];
  const int tid = threadIdx.x;
  const int64_t i = (int64_t)blockIdx.x * ROWS + tid;
  const bool live = i < g.m;
  if (only != nullptr && !__syncthreads_or(live && only[i] != 0)) return;   /* none of this CTA's rows was declined by the sliced kernel */
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

__global__ void k_gemv_col_fin(GemvArgs g, int splits, const uint32_t *part, const uint8_t *only)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.m || (only != nullptr && only[i] == 0)) return;
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

/* the sliced FP64 kernel pays four small launches around the product: worth it from about a million elements, and the acceptance
 * bound of qslice.cuh is stated for n >= 128 */
static bool gemv_sliced(int64_t m, int64_t n)
{
  if (fast_variant() == 3) return n >= 128 && m >= 1;            /* test hook: the sliced kernel at every size its bound covers */
  return fast_variant() == 2 && n >= 512 && m >= 148 * 4 && m * n >= (1 << 20);
}
static int64_t gemv_sliced_npad(int64_t n) { return (n + 4 * GVT_COLS - 1) / (4 * GVT_COLS) * (4 * GVT_COLS) + 4 * GVT_COLS; }   /* whole tiles of any box count, read past the last one */
/* column splits of the sliced kernel: a function of n ALONE (1024 columns per split, 512 for narrower matrices), so that the bits of
 * a row do not depend on how many rows the call has — a slab of a host matrix or the row block of one GPU gives what the whole call
 * gives — and a block of a few thousand rows still fills the chip (32768 columns: 32 splits x 32 row blocks of 4096 rows) */
static int64_t gemv_sliced_jsplit(int64_t n, int col_major)
{
  int64_t j = (n >= 16384 ? 1024 : 512) * (col_major ? 1 : GVT_NB_ROW);   /* row-major: NB partials per split, the same number of records */
  while ((n + j - 1) / j > 16384) j *= 2;                        /* gridDim.y */
  return j;
}
static int gemv_sliced_splits(int64_t n, int col_major) { return (int)((n + gemv_sliced_jsplit(n, col_major) - 1) / gemv_sliced_jsplit(n, col_major)); }
static int gemv_sliced_records(int64_t n, int col_major) { return gemv_sliced_splits(n, col_major) * (col_major ? 1 : GVT_NB_ROW); }
struct GvSlicedLayout { int64_t tab, flags, part, colwork, total; };   /* offsets in 16-byte elements */
static GvSlicedLayout gemv_sliced_layout(int64_t m, int64_t n, int col_major)
{
  GvSlicedLayout L;
  L.tab = 4;
  L.flags = L.tab + (GVT_XDBL * gemv_sliced_npad(n) + 1) / 2;
  L.part = L.flags + (m + 15) / 16;
  L.colwork = L.part + (int64_t)gemv_sliced_records(n, col_major) * m * GVT_REC / 4;
  L.total = L.colwork + (col_major ? 2 * (int64_t)gemv_col_splits(m, n) * m : 0);
  return L;
}

/* 16-byte elements of device scratch: fast col-major = a 32-byte window per split x row; the sliced kernel adds a 64-byte header,
 * the 200-byte table of every (padded) element of x, one flag byte per row and a 48-byte record per (split, row) */
int64_t gemv_work_elems(int64_t m, int64_t n, int col_major, int mode, int64_t m_plan)
{
  if (mode == 0 || fast_variant() == 0 || m <= 0 || n <= 0) return 0;
  const int64_t mp = m_plan > 0 ? m_plan : m;
  if (gemv_sliced(mp, n)) return gemv_sliced_layout(m, n, col_major).total;
  return col_major ? 2 * (int64_t)gemv_col_splits(m, n) * m : 0;
}

const uint8_t *gemv_sliced_rowflags(const q128 *work, int64_t m, int64_t n, int col_major, int64_t m_plan)
{
  const int64_t mp = m_plan > 0 ? m_plan : m;
  return gemv_sliced(mp, n) ? reinterpret_cast<const uint8_t *>(work + gemv_sliced_layout(m, n, col_major).flags) : nullptr;
}

bool make_quad_map(CUtensorMap *tm, const void *base, int64_t inner, int64_t outer, int64_t stride_bytes, int box_inner, int box_outer, bool swizzle128);

static cudaError_t launch_gemv_col_window(const GemvArgs &a, q128 *work, const uint8_t *only, cudaStream_t st)
{
  const int splits = gemv_col_splits(a.m, a.n);
  const int64_t jchunk = (((a.n + splits - 1) / splits) + 31) / 32 * 32;
  const int gy = (int)((a.n + jchunk - 1) / jchunk);
  dim3 grid((unsigned)((a.m + GV_COL_ROWS - 1) / GV_COL_ROWS), (unsigned)gy);
  k_gemv_col_wide<GV_COL_ROWS, 32><<<grid, GV_COL_ROWS, 0, st>>>(a, jchunk, reinterpret_cast<uint32_t *>(work), only);
  k_gemv_col_fin<<<(unsigned)((a.m + 127) / 128), 128, 0, st>>>(a, gy, reinterpret_cast<const uint32_t *>(work), only);
  count_launch(2);
  return cudaGetLastError();
}

static cudaError_t launch_gemv_sliced(const GemvArgs &a, cudaStream_t st)
{
  const GvSlicedLayout L = gemv_sliced_layout(a.m, a.n, a.col_major);
  if (a.work == nullptr || a.work_elems < L.total) return cudaErrorInvalidValue;
  CUtensorMap tm;
  const bool ok = (reinterpret_cast<uintptr_t>(a.A) & 15u) == 0 &&
                  (a.col_major ? make_quad_map(&tm, a.A, a.m, a.n, a.lda * 16, GVT_ROWS, GVT_COLS, false)
                               : make_quad_map(&tm, a.A, a.n, a.m, a.lda * 16, GVT_COLS, GVT_ROWS / GVT_NB_ROW, true));
  if (!ok) {   /* no tensor map for this matrix (alignment, extents): the window kernel takes the whole call */
    cudaMemsetAsync(a.work + L.flags, 1, (size_t)a.m, st);   /* (qb_gemv_last_declined then reports every row) */
    if (a.col_major) return launch_gemv_col_window(a, a.work + L.colwork, nullptr, st);
    k_gemv_row_wide<2, 128, 5><<<(unsigned)((a.m + 1) / 2), 128, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
  }
  const int64_t npad = gemv_sliced_npad(a.n);
  int32_t *hdr = reinterpret_cast<int32_t *>(a.work);
  double *tab = reinterpret_cast<double *>(a.work + L.tab);
  uint8_t *rowflag = reinterpret_cast<uint8_t *>(a.work + L.flags);
  uint32_t *part = reinterpret_cast<uint32_t *>(a.work + L.part);
  cudaError_t e = cudaMemsetAsync(hdr, 0, 64, st);
  if (e != cudaSuccess) return e;
  k_gv_xscan<<<(unsigned)std::min<int64_t>((a.n + 255) / 256, 148 * 8), 256, 0, st>>>(a, hdr);
  k_gv_xtab<<<(unsigned)((npad + 255) / 256), 256, 0, st>>>(a, hdr, tab, npad);
  const int64_t jsplit = gemv_sliced_jsplit(a.n, a.col_major);
  const int gy = gemv_sliced_splits(a.n, a.col_major);
  if (a.col_major) {
    dim3 grid((unsigned)((a.m + GVT_ROWS - 1) / GVT_ROWS), (unsigned)gy);
    k_gemv_f64<true, 1><<<grid, GVT_ROWS, gvt_smem(1), st>>>(tm, a, hdr, tab, jsplit, part);
  } else {
    constexpr int RB = GVT_ROWS / GVT_NB_ROW;
    dim3 grid((unsigned)((a.m + RB - 1) / RB), (unsigned)gy);
    if (gvt_smem(GVT_NB_ROW) > 48 * 1024) {
      static bool attr_set[QB_MAX_DEVICES] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < QB_MAX_DEVICES && !attr_set[dev]) {
        e = cudaFuncSetAttribute(k_gemv_f64<false, GVT_NB_ROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, gvt_smem(GVT_NB_ROW));
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
      }
    }
    k_gemv_f64<false, GVT_NB_ROW><<<grid, GVT_ROWS, gvt_smem(GVT_NB_ROW), st>>>(tm, a, hdr, tab, jsplit, part);
  }
  k_gemv_f64_fin<<<(unsigned)((a.m + 127) / 128), 128, 0, st>>>(a, hdr, gemv_sliced_records(a.n, a.col_major), part, rowflag);
  count_launch(4);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  /* the rows the sliced kernel declined: the window kernel of the same layout, other rows untouched */
  if (a.col_major) return launch_gemv_col_window(a, a.work + L.colwork, rowflag, st);
  k_gemv_row_wide_sel<128><<<(unsigned)std::min<int64_t>(a.m, 148 * 8), 128, 0, st>>>(a, rowflag);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_gemv(const GemvArgs &a, int mode, cudaStream_t st)
{
  if (a.m == 0 || a.n == 0) return cudaSuccess; /* level2.hpp:21,59: y untouched */
  if (mode != 0 && fast_variant() != 0) {
    if (gemv_sliced(a.m_plan > 0 ? a.m_plan : a.m, a.n)) return launch_gemv_sliced(a, st);
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
      if (a.work == nullptr || a.work_elems < 2 * (int64_t)gemv_col_splits(a.m, a.n) * a.m) return cudaErrorInvalidValue;
      return launch_gemv_col_window(a, a.work, nullptr, st);
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

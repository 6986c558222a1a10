/*
 * qb_level1.cu — binary128 dot / nrm2 (grid reduction) and axpy.
 *
 * Replaces QuadBLAS::dot (/root/reference/include/quadblas/algorithms/level1.hpp:80-137),
 * dot_parallel (:38-77), dot_kernel_vectorized (:14-35), QuadBLAS::axpy (:190-223) and the
 * sqrt in quadblas_qnrm2 / Vector::norm (interface/c_interface.hpp:34-44, cpp_classes.hpp:78-81).
 *
 * REFERENCE ORDER (mode 0): the reference's result depends on T = quadblas_get_num_threads()
 * (threading/openmp_utils.hpp:10-17): unit stride and n >= 500 -> T contiguous chunks of n/T
 * (last takes the remainder), each reduced by the two-lane kernel (even chain, odd chain,
 * lane0+lane1, odd tail), partials folded with add from +0 in tid order; n < 500 -> one two-lane
 * kernel; strided -> same chunking with ONE ascending chain per chunk (:104-120) or a single
 * chain (:128-134).  One GPU thread runs one chain, so the bits are the reference's for any T.
 *
 * FAST (mode 1): a deterministic two-level tree on the unrounded 192-bit window accumulator of
 * qwide.cuh.  Level 1: thread t of a fixed G x B grid accumulates elements t, t+GB, t+2GB, ...
 * (fully coalesced 128-bit loads, ~100 integer instructions per element instead of ~170 for a
 * rounded FMA); the B windows of a CTA are merged by a fixed shuffle/shared-memory tree.  Level 2:
 * one CTA merges the G block windows with the same fixed tree and rounds ONCE.  For fixed
 * (n, G, B) the tree is fixed, hence run-to-run and GPU-to-GPU reproducible.  The previous
 * generation of this path (rounded FMA chains, k_dot_fast_l1) is kept as `fast variant 0` for the
 * bench's side-by-side (qb_set_fast_variant).
 */
#include <cstdlib>
#include "qb_internal.h"
#include "q128_chain.cuh"
#include "qwide.cuh"
#include "qslice.cuh"
#include "qb_tc.cuh"

namespace qb {

__device__ __forceinline__ q128 ldg128_l1(const q128 *p)
{
  uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
  q128 r;
  r.lo = ((uint64_t)v.y << 32) | v.x;
  r.hi = ((uint64_t)v.w << 32) | v.z;
  return r;
}
__device__ __noinline__ q128 l1_add(q128 a, q128 b) { return q_add(a, b); }

/* ------------------------------------------------------------------ reference order */
/* one thread = one chain.  lanes = 2: thread (t, lane) walks x[s + 2p + lane]; lanes = 1: thread t
 * walks the whole chunk.  Output: part[t*lanes + lane]. */
__global__ void k_dot_ref_chains(DotArgs g, int64_t chunk, int nchunks, int lanes)
{
  const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= (int64_t)nchunks * lanes) return;
  const int64_t t = gt / lanes;
  const int lane = (int)(gt % lanes);
  const int64_t s = t * chunk;
  const int64_t e = (t == nchunks - 1) ? g.n : s + chunk;
  const int64_t len = e > s ? e - s : 0;
  qacc acc = qacc_zero();
  if (lanes == 2) {
    const int64_t h = len / 2;
    const q128 *xp = g.x + (s + lane) * g.incx, *yp = g.y + (s + lane) * g.incy;
    for (int64_t p = 0; p < h; ++p)
      qacc_fma(acc, qop_load(ldg128_l1(xp + 2 * p * g.incx)), qop_load(ldg128_l1(yp + 2 * p * g.incy)));
  } else {
    const q128 *xp = g.x + s * g.incx, *yp = g.y + s * g.incy;
    for (int64_t p = 0; p < len; ++p)
      qacc_fma(acc, qop_load(ldg128_l1(xp + p * g.incx)), qop_load(ldg128_l1(yp + p * g.incy)));
  }
  g.work[gt] = qacc_pack(acc);
}

/* per chunk: lane0 + lane1, then the odd tail (level1.hpp:27-32).  In place: work[t] <- chunk t. */
__global__ void k_dot_ref_lanes(DotArgs g, int64_t chunk, int nchunks)
{
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks) return;
  const int64_t s = t * chunk;
  const int64_t e = (t == nchunks - 1) ? g.n : s + chunk;
  const int64_t len = e > s ? e - s : 0;
  q128 r = q_zero(0);
  if (len > 0) {
    r = l1_add(g.work[2 * t], g.work[2 * t + 1]);
    if (len & 1) r = q_fma_slow_packed(g.x[(e - 1) * g.incx], g.y[(e - 1) * g.incy], r);
  }
  g.work[2 * (int64_t)nchunks + t] = r;
}

/* result = fold_t add(result, part[t]) from +0 in index order (level1.hpp:67-73); optional sqrt.
 * A single thread: the order is the contract. `direct` = no fold (n < 500 path, level1.hpp:40-43). */
__global__ void k_fold(const q128 *part, int64_t count, int direct, int do_sqrt, q128 *result)
{
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  q128 r = q_zero(0);
  if (direct) r = part[0];
  else for (int64_t t = 0; t < count; ++t) r = l1_add(r, part[t]);
  if (do_sqrt) r = q_sqrt(r);
  *result = r;
}

/* ------------------------------------------------------------------ fast mode */
template <int B>
__device__ __forceinline__ q128 block_tree(q128 v, q128 *sh)
{
  sh[threadIdx.x] = v;
  __syncthreads();
#pragma unroll 1
  for (int s = B / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = l1_add(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  return sh[0];
}

template <int B, int U>
__global__ void __launch_bounds__(B)
k_dot_fast_l1(DotArgs g)
{
  __shared__ q128 sh[B];
  const int64_t nthreads = (int64_t)gridDim.x * B;
  const int64_t t = (int64_t)blockIdx.x * B + threadIdx.x;
  qacc acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = qacc_zero();
  int64_t i = t;
  /* U independent chains per thread: U loads in flight, U-way ILP in the integer pipes */
  for (; i + (U - 1) * nthreads < g.n; i += U * nthreads) {
    q128 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      xv[u] = ldg128_l1(g.x + (i + u * nthreads) * g.incx);
      yv[u] = ldg128_l1(g.y + (i + u * nthreads) * g.incy);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) qacc_fma(acc[u], qop_load(xv[u]), qop_load(yv[u]));
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (i + u * nthreads < g.n)
      qacc_fma(acc[u], qop_load(ldg128_l1(g.x + (i + u * nthreads) * g.incx)), qop_load(ldg128_l1(g.y + (i + u * nthreads) * g.incy)));
  q128 v = qacc_pack(acc[0]);
#pragma unroll
  for (int u = 1; u < U; ++u) v = l1_add(v, qacc_pack(acc[u]));
  v = block_tree<B>(v, sh);
  if (threadIdx.x == 0) g.work[blockIdx.x] = v;
}

template <int B>
__global__ void __launch_bounds__(B)
k_dot_fast_l2(const q128 *part, int count, int do_sqrt, q128 *result)
{
  __shared__ q128 sh[B];
  q128 v = q_zero(0);
  /* fixed assignment: thread t folds part[t], part[t+B], ... in order */
  for (int i = threadIdx.x; i < count; i += B) v = l1_add(v, part[i]);
  v = block_tree<B>(v, sh);
  if (threadIdx.x == 0) *result = do_sqrt ? q_sqrt(v) : v;
}


/* ---- fast mode, unrounded window accumulator (qwide.cuh) ---- */
/* partial record: 8 words {w0..w5, E, bad} = 32 B */
__device__ __forceinline__ void qw_store(uint32_t *dst, const qwide &s, uint32_t bad)
{
  reinterpret_cast<uint4 *>(dst)[0] = make_uint4(s.w0, s.w1, s.w2, s.w3);
  reinterpret_cast<uint4 *>(dst)[1] = make_uint4(s.w4, s.w5, (uint32_t)s.E, bad);
}
__device__ __forceinline__ qwide qw_load(const uint32_t *src, uint32_t &bad)
{
  const uint4 a = reinterpret_cast<const uint4 *>(src)[0], b = reinterpret_cast<const uint4 *>(src)[1];
  qwide s;
  s.w0 = a.x; s.w1 = a.y; s.w2 = a.z; s.w3 = a.w; s.w4 = b.x; s.w5 = b.y; s.E = (int32_t)b.z;
  bad |= b.w;
  return s;
}

/* the same through L2 (records written by other CTAs of the running grid) */
__device__ __forceinline__ qwide qw_load_cg(const uint32_t *src, uint32_t &bad)
{
  const uint4 a = __ldcg(reinterpret_cast<const uint4 *>(src)), b = __ldcg(reinterpret_cast<const uint4 *>(src) + 1);
  qwide s;
  s.w0 = a.x; s.w1 = a.y; s.w2 = a.z; s.w3 = a.w; s.w4 = b.x; s.w5 = b.y; s.E = (int32_t)b.z;
  bad |= b.w;
  return s;
}

/* Sum of the windows of the B threads of a CTA (result valid in thread 0), by qw_align7 / qw7_add of qwide.cuh: the largest anchor of the
 * CTA (warp max + one exchange through shared memory), every window shifted ONCE to it, then 224-bit integer sums — five shuffle
 * rounds inside the warps, thread 0 adds the warp results.  Integer addition is associative, so the result is the same bits for any
 * order.  (The first version was a tree of pairwise qw_merge: 8 dependent merges per tree, 18 us of fixed cost per kernel.) */
template <int B>
__device__ __forceinline__ int32_t qw_block_emax(int32_t e, uint32_t *sh /* B/32 words */)
{
  e = __reduce_max_sync(0xffffffffu, e);
  __syncthreads();                                   /* sh may still be in use by the phase before */
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = (uint32_t)e;
  __syncthreads();
  int32_t E = (int32_t)sh[0];
#pragma unroll
  for (int w = 1; w < B / 32; ++w) E = max(E, (int32_t)sh[w]);
  __syncthreads();
  return E;
}
template <int B>
__device__ __forceinline__ qwide qw_block_add7(uint32_t (&a)[7], int32_t E, uint32_t &bad, uint32_t *sh /* 8 * B/32 words */)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    uint32_t b[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) b[j] = __shfl_down_sync(0xffffffffu, a[j], off);
    qw7_add(a, b);
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    reinterpret_cast<uint4 *>(sh + 8 * warp)[0] = make_uint4(a[0], a[1], a[2], a[3]);
    reinterpret_cast<uint4 *>(sh + 8 * warp)[1] = make_uint4(a[4], a[5], a[6], bad);
  }
  __syncthreads();
  qwide v = qw_zero();
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int w = 1; w < B / 32; ++w) {
      const uint4 p = reinterpret_cast<const uint4 *>(sh + 8 * w)[0], q = reinterpret_cast<const uint4 *>(sh + 8 * w)[1];
      const uint32_t b[7] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z};
      qw7_add(a, b);
      bad |= q.w;
    }
    v = qw_from7(a, E);
  }
  return v;
}
template <int B>
__device__ __forceinline__ qwide qw_block_tree(qwide v, uint32_t &bad, uint32_t *sh /* 8 * B/32 words */)
{
  const int32_t E = qw_block_emax<B>(v.E, sh);
  uint32_t a[7];
  qw_align7(v, E, a);
  return qw_block_add7<B>(a, E, bad, sh);
}
/* the same over `count` records {w0..w5, E, bad} written by other CTAs (`stride` words apart): thread t takes the records t, t + B, ... */
template <int B>
__device__ __forceinline__ qwide qw_block_fold_records(const uint32_t *rec, int stride, int count, uint32_t &bad, uint32_t *sh)
{
  int32_t e = QW_EMPTY;
  for (int i = threadIdx.x; i < count; i += B) e = max(e, (int32_t)__ldcg(rec + (int64_t)stride * i + 6));
  const int32_t E = qw_block_emax<B>(e, sh);
  uint32_t a[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
  for (int i = threadIdx.x; i < count; i += B) {
    uint32_t b[7];
    qw_align7(qw_load_cg(rec + (int64_t)stride * i, bad), E, b);
    qw7_add(a, b);
  }
  return qw_block_add7<B>(a, E, bad, sh);
}

/* U element pairs (2U 128-bit loads) in flight per thread; the U accumulate steps are branch-free
 * (qwa_fma, one scratch column each) and go into ONE accumulator — the window update is a plain sum,
 * so the order in which fast and declined steps are applied does not matter. */
template <int B, int U, bool SAME, int MINB>
__global__ void __launch_bounds__(B, MINB)
k_dot_wide_l1(DotArgs g)
{
  static_assert(8 * (B / 32) <= U * QWA_COL_WORDS * B, "the reduction records reuse the scratch columns");
  __shared__ __align__(16) uint32_t scr[U * QWA_COL_WORDS * B];
  uint32_t *sh = scr;
  if (g.only_if != nullptr && *g.only_if == 0u) return;   /* queued behind k_sumsq_f64: runs only when that kernel declined */
  const int64_t nthreads = (int64_t)gridDim.x * B;
  const int64_t t = (int64_t)blockIdx.x * B + threadIdx.x;
  qwacc acc = qwa_zero();
  uint32_t bad = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) qwa_col_init(scr + u * QWA_COL_WORDS * B + threadIdx.x, B);
  /* running pointers (element t, t + nthreads, ...): no 64-bit index multiplies in the loop */
  const q128 *xp = g.x + t * g.incx, *yp = g.y + t * g.incy;
  const int64_t xs = nthreads * g.incx, ys = nthreads * g.incy;
  int64_t left = (g.n > t) ? (g.n - t + nthreads - 1) / nthreads : 0;   /* elements of this thread */
  for (; left >= U; left -= U) {
    q128 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      xv[u] = ldg128_l1(xp); xp += xs;
      if (!SAME) { yv[u] = ldg128_l1(yp); yp += ys; }
    }
    bool rare = false, rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const qop a = qop_load_n(xv[u]);
      rr[u] = qwa_fma(acc, a, SAME ? a : qop_load_n(yv[u]), scr + u * QWA_COL_WORDS * B + threadIdx.x, B);
      rare |= rr[u];
    }
    if (rare) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rr[u]) qwa_fma_rare(acc, xv[u], SAME ? xv[u] : yv[u], bad);
    }
  }
  for (; left > 0; --left) {
    const q128 xv = ldg128_l1(xp);
    const q128 yv = SAME ? xv : ldg128_l1(yp);
    xp += xs; yp += ys;
    if (qwa_fma(acc, qop_load_n(xv), qop_load_n(yv), scr + threadIdx.x, B)) qwa_fma_rare(acc, xv, yv, bad);
  }
  __syncthreads();                       /* the scratch columns become the reduction records */
  qwide v = qw_block_tree<B>(qwa_fold(acc), bad, sh);
  /* second level in the same launch: the CTA that takes the last ticket merges all the records, in index order (thread t takes
   * records t, t + B, ...; then the fixed block tree), so the result does not depend on which CTA that is.  Saves the second
   * launch, which is a fifth of the time of a 10^7-element dot. */
  __shared__ int is_last;
  uint32_t *rec = reinterpret_cast<uint32_t *>(g.work);
  if (threadIdx.x == 0) {
    qw_store(rec + 8 * (int64_t)blockIdx.x, v, bad);
    __threadfence();
    is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  v = qw_zero();
  bad = 0;
  v = qw_block_fold_records<B>(rec, 8, (int)gridDim.x, bad, sh);
  if (threadIdx.x == 0) {
    const q128 r = qw_finish(v, bad);
    *g.result = g.do_sqrt ? q_sqrt(r) : r;
    *g.ticket = 0u;                      /* ready for the next call (stream order) */
  }
}

/* k_dot_wide_l1 for two contiguous vectors, fed by the copy engine: the kernel above waits on its own loads (long-scoreboard stalls
 * 3.9 per issue, issue slots 49 % busy at 24 warps per SM, profiles/r2_dot_wide_ncu_full.txt); here tiles of B x U pairs arrive by
 * cp.async.bulk into a two-stage ring and a tile is exactly one interleaved group of U branch-free accumulate steps per thread. */
template <int B, int U, bool SAME, int MINB>
__global__ void __launch_bounds__(B, MINB)
k_dot_wide_tma(DotArgs g)
{
  constexpr int TE = B * U;
  extern __shared__ __align__(128) unsigned char dw_sm[];
  uint4 *tx = reinterpret_cast<uint4 *>(dw_sm);                      /* [2][TE] */
  uint4 *ty = SAME ? tx : tx + 2 * TE;                               /* [2][TE] (x * x: the same tiles) */
  uint32_t *scr = reinterpret_cast<uint32_t *>(tx + (SAME ? 2 : 4) * TE);   /* [U * QWA_COL_WORDS * B] */
  uint64_t *full = reinterpret_cast<uint64_t *>(scr + U * QWA_COL_WORDS * B);
  uint32_t *sh = scr;
  __shared__ int is_last;
  if (g.only_if != nullptr && *g.only_if == 0u) return;
  const int tid = threadIdx.x;
  qwacc acc = qwa_zero();
  uint32_t bad = 0;
#pragma unroll
  for (int u = 0; u < U; ++u) qwa_col_init(scr + u * QWA_COL_WORDS * B + tid, B);
  const int64_t nfull = g.n / TE;
  const int64_t mine = nfull > blockIdx.x ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t i, int s) {
    const int64_t off = (blockIdx.x + i * gridDim.x) * TE;
    tc::mbar_expect_tx(&full[s], (SAME ? 1 : 2) * TE * 16);
    tc::bulk_load_1d(tx + s * TE, g.x + off, TE * 16, &full[s]);
    if (!SAME) tc::bulk_load_1d(ty + s * TE, g.y + off, TE * 16, &full[s]);
  };
  if (tid == 0) {
    tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    if (mine > 0) issue(0, 0);
    if (mine > 1) issue(1, 1);
  }
  auto ld = [](const uint4 &v) { q128 r; r.lo = ((uint64_t)v.y << 32) | v.x; r.hi = ((uint64_t)v.w << 32) | v.z; return r; };
  if (mine > 0) {
    /* an empty accumulator declines every product (its anchor is below everything) and the out-of-line step is slow, U of them in a
     * row per thread (measured: 18.4 -> 16.4 us per call at n = 10^6).  The anchor is therefore set from the first tile before the
     * loop, where the declined steps would have put it: QW_SLACK bits above the largest of this thread's first U products of
     * normal operands. */
    tc::mbar_wait(&full[0], 0u);
    int32_t e0 = QW_EMPTY;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t ea = (tx[u * B + tid].w >> 16) & 0x7fffu, eb = SAME ? ea : ((ty[u * B + tid].w >> 16) & 0x7fffu);
      if (ea - 1u < 0x7ffeu && eb - 1u < 0x7ffeu) e0 = max(e0, (int32_t)(ea + eb));
    }
    if (e0 != QW_EMPTY) acc.E = e0 + QW_SLACK;
  }
  for (int64_t i = 0; i < mine; ++i) {
    const int s = (int)(i & 1);
    tc::mbar_wait(&full[s], (uint32_t)(i >> 1) & 1u);
    q128 xv[U], yv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { xv[u] = ld(tx[s * TE + u * B + tid]); yv[u] = SAME ? xv[u] : ld(ty[s * TE + u * B + tid]); }
    bool rare = false, rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const qop a = qop_load_n(xv[u]);
      rr[u] = qwa_fma(acc, a, SAME ? a : qop_load_n(yv[u]), scr + u * QWA_COL_WORDS * B + tid, B);
      rare |= rr[u];
    }
    if (rare) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (rr[u]) qwa_fma_rare(acc, xv[u], yv[u], bad);
    }
    __syncthreads();
    if (tid == 0 && i + 2 < mine) issue(i + 2, s);
  }
  if (blockIdx.x == (unsigned)(nfull % gridDim.x)) {   /* the partial tile */
    for (int64_t j = nfull * TE + tid; j < g.n; j += B) {
      const q128 xv = ldg128_l1(g.x + j), yv = SAME ? xv : ldg128_l1(g.y + j);
      if (qwa_fma(acc, qop_load_n(xv), qop_load_n(yv), scr + tid, B)) qwa_fma_rare(acc, xv, yv, bad);
    }
  }
  __syncthreads();                       /* the scratch columns become the reduction records */
  qwide v = qw_block_tree<B>(qwa_fold(acc), bad, sh);
  uint32_t *rec = reinterpret_cast<uint32_t *>(g.work);
  if (tid == 0) {
    qw_store(rec + 8 * (int64_t)blockIdx.x, v, bad);
    __threadfence();
    is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  v = qw_zero();
  bad = 0;
  v = qw_block_fold_records<B>(rec, 8, (int)gridDim.x, bad, sh);
  if (tid == 0) {
    const q128 r = qw_finish(v, bad);
    *g.result = g.do_sqrt ? q_sqrt(r) : r;
    *g.ticket = 0u;
  }
}

/* qnrm2 on the FP64 pipe (qslice.cuh, qs_square_step): every thread strides over the vector with U loads in flight, keeps six exact
 * column sums and a 256-bit window (shared memory), and ends as a qwide that takes the same block tree / last-ticket fold as the
 * window kernel.  All terms are positive, so the truncation (< 2^-125 of the sum per element) needs no acceptance test; an Inf, NaN
 * or nonzero subnormal element sets *g.only_if instead of a result, and the window kernel queued behind this one then runs. */
template <int B>
__device__ __noinline__ void sq_flush(double c0, double c1, double c2, double c3, double c4, double c5, uint64_t *w)
{
  qs_flush(c0, c1, c2, c3, c4, c5, w, B);
}
/* the element the hot form left out: its accumulators go to the window first (t: in = the six columns, out = the eight fresh
 * accumulators), qs_rare moves the anchor or flags the call, then the element is stepped.  Through memory so that the call does
 * not pin the hot loop's registers. */
template <int B>
__device__ __noinline__ void sq_rare_mem(double *t, int32_t *st, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint64_t *w)
{
  qs_cols C; C.c0 = t[0]; C.c1 = t[1]; C.c2 = t[2]; C.c3 = t[3]; C.c4 = t[4]; C.c5 = t[5];
  qs_row S; S.anc = st[0]; S.dmax = 0;
  uint32_t flags = (uint32_t)st[1];
  const uint32_t e = (w3 >> 16) & 0x7fffu;
  const uint32_t sh = ((uint32_t)(e - 1u) >= (uint32_t)S.anc) ? qs_rare(C, S, flags, e, w0, w1, w2, w3, w, B, 2u) : min((uint32_t)S.anc - e, QS_SHMAX);
  qs_flush(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, w, B);
  qs_sq_cols Q = qs_sq_zero();
  qs_square_step(Q, w0, w1, w2, w3, sh);
  t[0] = Q.d0; t[1] = Q.o1; t[2] = Q.o2; t[3] = Q.d2; t[4] = Q.o3; t[5] = Q.o4; t[6] = Q.d4; t[7] = Q.o5;
  st[0] = S.anc; st[1] = (int32_t)flags;
}

template <int B, int U, int MINB>
__global__ void __launch_bounds__(B, MINB)
k_sumsq_f64(DotArgs g)
{
  __shared__ uint64_t win[4 * B];
  __shared__ __align__(16) uint32_t sh[8 * (B / 32)];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * B;
  const int64_t t = (int64_t)blockIdx.x * B + tid;
  qs_sq_cols Q = qs_sq_zero();
  int32_t anc = QS_ANCMIN;
  uint32_t flags = 0;
  uint64_t *wn = win + tid;
#pragma unroll
  for (int k = 0; k < 4; ++k) wn[k * B] = 0ull;
  const q128 *xp = g.x + t * g.incx;
  const int64_t xs = nthreads * g.incx;
  int64_t left = (g.n > t) ? (g.n - t + nthreads - 1) / nthreads : 0;
  int tile = 0;
  auto one = [&](const q128 &xv) {
    const uint32_t w0 = (uint32_t)xv.lo, w1 = (uint32_t)(xv.lo >> 32), w2 = (uint32_t)xv.hi, w3 = (uint32_t)(xv.hi >> 32);
    const uint32_t e = (w3 >> 16) & 0x7fffu;
    qs_square_step(Q, w0, w1, w2, w3, min((uint32_t)anc - e, QS_SHMAX));   /* a zero or an element above the anchor adds nothing here */
    if ((uint32_t)(e - 1u) >= (uint32_t)anc) {
      const qs_cols C = qs_sq_columns(Q);
      double tt[8] = {C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, 0.0, 0.0};
      int32_t st[2] = {anc, (int32_t)flags};
      sq_rare_mem<B>(tt, st, w0, w1, w2, w3, wn);
      Q.d0 = tt[0]; Q.o1 = tt[1]; Q.o2 = tt[2]; Q.d2 = tt[3]; Q.o3 = tt[4]; Q.o4 = tt[5]; Q.d4 = tt[6]; Q.o5 = tt[7];
      anc = st[0]; flags = (uint32_t)st[1];
    }
  };
  auto flush = [&]() {
    const qs_cols C = qs_sq_columns(Q);
    sq_flush<B>(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, wn);
    Q = qs_sq_zero();
  };
  /* two register sets of U elements: the loads of the next batch are issued before the arithmetic of this one */
  q128 xa[U], xb[U];
  if (left >= U) {
#pragma unroll
    for (int u = 0; u < U; ++u) { xa[u] = ldg128_l1(xp); xp += xs; }
  }
  while (left >= U) {
    if (left >= 2 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) { xb[u] = ldg128_l1(xp); xp += xs; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(xa[u]);
    left -= U;
    if (left < U) break;
    if (left >= 2 * U) {
#pragma unroll
      for (int u = 0; u < U; ++u) { xa[u] = ldg128_l1(xp); xp += xs; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) one(xb[u]);
    left -= U;
    tile += 2 * U;
    if (tile >= QS_TILE - 2 * U) { tile = 0; flush(); }
  }
  flush();                                /* (the ragged tail below adds fewer than U elements) */
  for (; left > 0; --left) { const q128 xv = ldg128_l1(xp); xp += xs; one(xv); }
  flush();
  qwide v = qw_block_tree<B>(qs_to_qwide(wn, B, anc, anc), flags, sh);
  uint32_t *rec = reinterpret_cast<uint32_t *>(g.work);
  if (tid == 0) {
    qw_store(rec + 8 * (int64_t)blockIdx.x, v, flags);
    __threadfence();
    is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  v = qw_zero();
  flags = 0;
  v = qw_block_fold_records<B>(rec, 8, (int)gridDim.x, flags, sh);
  if (tid == 0) {
    if (flags & QS_FALLBACK) *g.only_if = 1u;
    else {
      *g.only_if = 0u;
      const q128 r = qw_finish(v, 0u);
      *g.result = g.do_sqrt ? q_sqrt(r) : r;
    }
    *g.ticket = 0u;
  }
}

/* the same for a contiguous vector, fed by the copy engine: tiles of B x SQ_PER elements (16 KB) arrive by cp.async.bulk into a
 * two-stage ring — 32 KB in flight per CTA and six CTAs per SM, without a register spent on it (k_sumsq_f64 above waits on its own
 * loads: long-scoreboard stalls 3.9 per issue at 24 warps per SM).  CTA b takes the tiles b, b + gridDim.x, ...; the last, partial
 * tile of the vector is read with plain loads by the CTA it falls to. */
template <int B, int MINB, int SQ_PER>
__global__ void __launch_bounds__(B, MINB)
k_sumsq_tma(DotArgs g)
{
  constexpr int TE = B * SQ_PER;                       /* elements per tile */
  constexpr int FLUSH_TILES = 56 / SQ_PER;             /* 56 elements: the accumulators are still exact */
  __shared__ __align__(128) uint4 tile[2][TE];
  __shared__ uint64_t win[4 * B];
  __shared__ __align__(16) uint32_t sh[8 * (B / 32)];
  __shared__ uint64_t full[2];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  qs_sq_cols Q = qs_sq_zero();
  int32_t anc = QS_ANCMIN;
  uint32_t flags = 0;
  uint64_t *wn = win + tid;
#pragma unroll
  for (int k = 0; k < 4; ++k) wn[k * B] = 0ull;
  const int64_t nfull = g.n / TE;                      /* whole tiles of the vector */
  const int64_t mine = nfull > blockIdx.x ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t i, int s) {
    tc::mbar_expect_tx(&full[s], TE * 16);
    tc::bulk_load_1d(&tile[s][0], g.x + (blockIdx.x + i * gridDim.x) * TE, TE * 16, &full[s]);
  };
  if (tid == 0) {
    tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    if (mine > 0) issue(0, 0);
    if (mine > 1) issue(1, 1);
  }
  auto one = [&](const uint4 &v) {
    const uint32_t e = (v.w >> 16) & 0x7fffu;
    qs_square_step(Q, v.x, v.y, v.z, v.w, min((uint32_t)anc - e, QS_SHMAX));   /* a zero or an element above the anchor adds nothing here */
    if ((uint32_t)(e - 1u) >= (uint32_t)anc) {
      const qs_cols C = qs_sq_columns(Q);
      double tt[8] = {C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, 0.0, 0.0};
      int32_t st[2] = {anc, (int32_t)flags};
      sq_rare_mem<B>(tt, st, v.x, v.y, v.z, v.w, wn);
      Q.d0 = tt[0]; Q.o1 = tt[1]; Q.o2 = tt[2]; Q.d2 = tt[3]; Q.o3 = tt[4]; Q.o4 = tt[5]; Q.d4 = tt[6]; Q.o5 = tt[7];
      anc = st[0]; flags = (uint32_t)st[1];
    }
  };
  auto flush = [&]() {
    const qs_cols C = qs_sq_columns(Q);
    sq_flush<B>(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, wn);
    Q = qs_sq_zero();
  };
  for (int64_t i = 0; i < mine; ++i) {
    const int s = (int)(i & 1);
    tc::mbar_wait(&full[s], (uint32_t)(i >> 1) & 1u);
    const uint4 *tp = &tile[s][tid];
#pragma unroll
    for (int k = 0; k < SQ_PER; ++k) one(tp[k * B]);
    if ((i % FLUSH_TILES) == FLUSH_TILES - 1) flush();
    __syncthreads();                                   /* every thread is done with stage s */
    if (tid == 0 && i + 2 < mine) issue(i + 2, s);
  }
  flush();
  if (blockIdx.x == (unsigned)(nfull % gridDim.x)) {   /* the partial tile */
    for (int64_t j = nfull * TE + tid; j < g.n; j += B) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(g.x + j));
      one(v);
    }
    flush();
  }
  qwide v = qw_block_tree<B>(qs_to_qwide(wn, B, anc, anc), flags, sh);
  uint32_t *rec = reinterpret_cast<uint32_t *>(g.work);
  if (tid == 0) {
    qw_store(rec + 8 * (int64_t)blockIdx.x, v, flags);
    __threadfence();
    is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  v = qw_zero();
  flags = 0;
  v = qw_block_fold_records<B>(rec, 8, (int)gridDim.x, flags, sh);
  if (tid == 0) {
    if (flags & QS_FALLBACK) *g.only_if = 1u;
    else {
      *g.only_if = 0u;
      const q128 r = qw_finish(v, 0u);
      *g.result = g.do_sqrt ? q_sqrt(r) : r;
    }
    *g.ticket = 0u;
  }
}

/* qdot of two contiguous vectors on the FP64 pipe: both factors are sliced on the fly (qs_dot_step), each thread keeps an anchor
 * per vector, tiles of B x DQ_PER elements of x and of y arrive by cp.async.bulk into a two-stage ring.  The truncation is
 * relative to (largest |x| so far) x (largest |y| so far) of the thread, so the call is accepted when some product comes within
 * 2^-12 of the largest anchor pair (qs_accept's test on max e(x_j) + e(y_j)); otherwise — or with an Inf / NaN / subnormal — the
 * kernel sets *g.only_if and the window kernel queued behind it computes the result. */
constexpr int DQ_PER = 4;
template <int B>
__device__ __noinline__ void dq_rare_mem(double *t, int32_t *st, const uint4 *xy, uint64_t *w)
{
  qs_cols C; C.c0 = t[0]; C.c1 = t[1]; C.c2 = t[2]; C.c3 = t[3]; C.c4 = t[4]; C.c5 = t[5];
  qs_row SX, SY; SX.anc = st[0]; SY.anc = st[1]; SX.dmax = SY.dmax = 0;
  uint32_t flags = (uint32_t)st[2];
  const uint4 x = xy[0], y = xy[1];
  const uint32_t ex = (x.w >> 16) & 0x7fffu, ey = (y.w >> 16) & 0x7fffu;
  uint32_t shx = min((uint32_t)SX.anc - ex, QS_SHMAX), shy = min((uint32_t)SY.anc - ey, QS_SHMAX);
  if ((uint32_t)(ex - 1u) >= (uint32_t)SX.anc) shx = qs_rare(C, SX, flags, ex, x.x, x.y, x.z, x.w, w, B);
  if ((uint32_t)(ey - 1u) >= (uint32_t)SY.anc) shy = qs_rare(C, SY, flags, ey, y.x, y.y, y.z, y.w, w, B);
  if (shx < QS_SHMAX && shy < QS_SHMAX) qs_dot_step(C, x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, shx, shy);
  t[0] = C.c0; t[1] = C.c1; t[2] = C.c2; t[3] = C.c3; t[4] = C.c4; t[5] = C.c5;
  st[0] = SX.anc; st[1] = SY.anc; st[2] = (int32_t)flags;
}

template <int B>
__device__ __forceinline__ int block_max(int v, int *shm /* B / 32 ints */)
{
  v = __reduce_max_sync(0xffffffffu, v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = shm[0];
#pragma unroll
  for (int w = 1; w < B / 32; ++w) r = max(r, shm[w]);
  return r;
}

template <int B, int MINB>
__global__ void __launch_bounds__(B, MINB)
k_dot_f64_tma(DotArgs g)
{
  constexpr int TE = B * DQ_PER;
  __shared__ __align__(128) uint4 tx[2][TE];
  __shared__ __align__(128) uint4 ty[2][TE];
  __shared__ uint64_t win[4 * B];
  __shared__ __align__(16) uint32_t sh[8 * (B / 32)];
  __shared__ int shm[B / 32];
  __shared__ uint64_t full[2];
  __shared__ int is_last;
  const int tid = threadIdx.x;
  qs_cols C = qs_cols_zero();
  int32_t ancx = QS_ANCMIN, ancy = QS_ANCMIN, dmax = QS_EXNONE;
  uint32_t flags = 0;
  uint64_t *wn = win + tid;
#pragma unroll
  for (int k = 0; k < 4; ++k) wn[k * B] = 0ull;
  const int64_t nfull = g.n / TE;
  const int64_t mine = nfull > blockIdx.x ? (nfull - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t i, int s) {
    const int64_t off = (blockIdx.x + i * gridDim.x) * TE;
    tc::mbar_expect_tx(&full[s], 2 * TE * 16);
    tc::bulk_load_1d(&tx[s][0], g.x + off, TE * 16, &full[s]);
    tc::bulk_load_1d(&ty[s][0], g.y + off, TE * 16, &full[s]);
  };
  if (tid == 0) {
    tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    if (mine > 0) issue(0, 0);
    if (mine > 1) issue(1, 1);
  }
  auto one = [&](const uint4 &x, const uint4 &y) {
    const uint32_t ex = (x.w >> 16) & 0x7fffu, ey = (y.w >> 16) & 0x7fffu;
    /* a zero factor or one above its anchor gets QS_SHMAX: the hot step adds nothing then */
    qs_dot_step(C, x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w, min((uint32_t)ancx - ex, QS_SHMAX), min((uint32_t)ancy - ey, QS_SHMAX));
    if (((uint32_t)(ex - 1u) >= (uint32_t)ancx) | ((uint32_t)(ey - 1u) >= (uint32_t)ancy)) {
      double tt[6] = {C.c0, C.c1, C.c2, C.c3, C.c4, C.c5};
      int32_t st[3] = {ancx, ancy, (int32_t)flags};
      uint4 xy[2] = {x, y};
      dq_rare_mem<B>(tt, st, xy, wn);
      C.c0 = tt[0]; C.c1 = tt[1]; C.c2 = tt[2]; C.c3 = tt[3]; C.c4 = tt[4]; C.c5 = tt[5];
      ancx = st[0]; ancy = st[1]; flags = (uint32_t)st[2];
      if (ex != 0u && ey != 0u) dmax = max(dmax, (int32_t)(ex + ey));   /* a zero factor is no product */
    } else {
      dmax = max(dmax, (int32_t)(ex + ey));
    }
  };
  auto flush = [&]() { sq_flush<B>(C.c0, C.c1, C.c2, C.c3, C.c4, C.c5, wn); C = qs_cols_zero(); };
  for (int64_t i = 0; i < mine; ++i) {
    const int s = (int)(i & 1);
    tc::mbar_wait(&full[s], (uint32_t)(i >> 1) & 1u);
#pragma unroll
    for (int k = 0; k < DQ_PER; ++k) one(tx[s][k * B + tid], ty[s][k * B + tid]);
    if ((i & 15) == 15) flush();                      /* 64 pairs: the columns are still exact */
    __syncthreads();
    if (tid == 0 && i + 2 < mine) issue(i + 2, s);
  }
  flush();
  if (blockIdx.x == (unsigned)(nfull % gridDim.x)) {   /* the partial tile */
    for (int64_t j = nfull * TE + tid; j < g.n; j += B)
      one(__ldg(reinterpret_cast<const uint4 *>(g.x + j)), __ldg(reinterpret_cast<const uint4 *>(g.y + j)));
    flush();
  }
  int asum = block_max<B>(ancx + ancy, shm);
  int dm = block_max<B>(dmax, shm);
  qwide v = qw_block_tree<B>(qs_to_qwide(wn, B, ancx, ancy), flags, sh);
  uint32_t *rec = reinterpret_cast<uint32_t *>(g.work);
  if (tid == 0) {
    uint32_t *dst = rec + 12 * (int64_t)blockIdx.x;
    qw_store(dst, v, flags);
    dst[8] = (uint32_t)asum; dst[9] = (uint32_t)dm;
    __threadfence();
    is_last = atomicAdd(g.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  v = qw_zero();
  flags = 0;
  asum = 0; dm = QS_EXNONE;
  for (int i = tid; i < (int)gridDim.x; i += B) {
    const uint32_t *src = rec + 12 * (int64_t)i;
    asum = max(asum, (int)__ldcg(src + 8)); dm = max(dm, (int)__ldcg(src + 9));
  }
  asum = block_max<B>(asum, shm);
  dm = block_max<B>(dm, shm);
  __syncthreads();
  v = qw_block_fold_records<B>(rec, 12, (int)gridDim.x, flags, sh);
  if (tid == 0) {
    if ((flags & QS_FALLBACK) || dm < asum - QS_ACCEPT) *g.only_if = 1u;
    else {
      *g.only_if = 0u;
      const q128 r = qw_finish(v, 0u);
      *g.result = g.do_sqrt ? q_sqrt(r) : r;
    }
    *g.ticket = 0u;
  }
}

static constexpr int FAST_B = 256;       /* rounded-chain variant and the second-level CTA */
static constexpr int FAST_GRID = 148 * 4;
static constexpr int WIDE_B = 128;       /* window variant: 4 scratch columns per thread = 24 KB per CTA */
static constexpr int WIDE_GRID = 148 * 6;
static constexpr int SUMSQ_B = 128, SUMSQ_CTAS = 6, SUMSQ_GRID = 148 * SUMSQ_CTAS;

int64_t dot_work_elems(int64_t n, int T, int mode)
{
  (void)n;
  if (mode != 0) return 3 * SUMSQ_GRID;   /* 32-byte window records (48 bytes with the anchors of the sliced dot) */
  return 3 * (int64_t)(T < 1 ? 1 : T) + 4;
}

cudaError_t launch_dot(const DotArgs &a, int mode, cudaStream_t st)
{
  DotArgs g = a;
  if (g.n == 0) { /* level1.hpp:83-84: +0 (sqrt(+0) = +0) */
    cudaError_t e = cudaMemsetAsync(g.result, 0, 16, st);
    return e;
  }
  if (mode != 0) {
    if (fast_variant() == 0) { /* rounded-FMA chains (previous generation, kept for comparison) */
      int grid = FAST_GRID;
      const int64_t need = (g.n + FAST_B - 1) / FAST_B;
      if (need < grid) grid = (int)need;
      k_dot_fast_l1<FAST_B, 4><<<grid, FAST_B, 0, st>>>(g);
      k_dot_fast_l2<FAST_B><<<1, FAST_B, 0, st>>>(g.work, grid, g.do_sqrt, g.result);
    } else {
      const bool same = (g.x == g.y && g.incx == g.incy);
      /* 128 threads x 4 elements in flight, 6 CTAs per SM resident (80 registers, 24 KB of scratch columns):
       * the grid is exactly one wave, every thread strides over the whole vector; the CTA that finishes last folds the
       * per-CTA records (one launch) */
      if (g.ticket == nullptr) return cudaErrorInvalidValue;
      /* from 2^24 elements (measured: at 10^7 the window kernel's single launch still wins, 79 vs 89 us); variant 3 = test hook, any size */
      if (same && g.only_if != nullptr && ((fast_variant() == 2 && g.n >= (1 << 24)) || fast_variant() == 3)) {
        /* sum of squares on the FP64 pipe; the window kernel is queued behind it and runs only if an Inf / NaN / subnormal made
         * the sliced kernel decline */
        /* tiles of 1024 elements at 6 CTAs per SM: measured against 512-element tiles at 8 / 9 / 10 CTAs (0.361 / 0.363 / 0.386 ms at n = 10^8)
         * and 256-element tiles at 12 (0.439 ms) — more, smaller tiles only add barrier rounds; 0.341 ms as built */
        if (g.incx == 1 && (reinterpret_cast<uintptr_t>(g.x) & 15u) == 0) k_sumsq_tma<SUMSQ_B, SUMSQ_CTAS, 8><<<SUMSQ_GRID, SUMSQ_B, 0, st>>>(g);
        else k_sumsq_f64<SUMSQ_B, 4, SUMSQ_CTAS><<<SUMSQ_GRID, SUMSQ_B, 0, st>>>(g);
        k_dot_wide_l1<WIDE_B, 4, true, WIDE_GRID / 148><<<WIDE_GRID, WIDE_B, 0, st>>>(g);
        count_launch(2);
        return cudaGetLastError();
      }
      /* two different contiguous vectors with both factors sliced on the fly: measured SLOWER than the window kernel (n = 10^8:
       * 0.675 vs 0.607 ms — two slicings per product cost more than the integer multiplier saves, and 32 bytes per product leave
       * the window kernel at 0.82 of HBM anyway), so it is not part of the default; fast variant 3 keeps it reachable for the tests */
      if (!same && g.only_if != nullptr && g.incx == 1 && g.incy == 1 && ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.y)) & 15u) == 0 &&
          fast_variant() == 3) {
        k_dot_f64_tma<SUMSQ_B, SUMSQ_CTAS><<<SUMSQ_GRID, SUMSQ_B, 0, st>>>(g);
        k_dot_wide_l1<WIDE_B, 4, false, WIDE_GRID / 148><<<WIDE_GRID, WIDE_B, 0, st>>>(g);
        count_launch(2);
        return cudaGetLastError();
      }
      g.only_if = nullptr;
      /* contiguous vectors, large enough for tiles: the copy-engine-fed form of the window kernel (U = 4 pairs per thread and tile,
       * 3 CTAs per SM: 56 KB of tiles and scratch columns each; measured against U = 2 x 6 CTAs and U = 3 x 4 CTAs) */
      if (g.incx == 1 && g.incy == 1 && ((reinterpret_cast<uintptr_t>(g.x) | reinterpret_cast<uintptr_t>(g.y)) & 15u) == 0 && g.n >= (1 << 19)) {
        constexpr int U = 4, CT = 3;
        auto go = [&](auto kern, int smem) -> cudaError_t {
          cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          if (ae != cudaSuccess) return ae;
          kern<<<148 * CT, WIDE_B, smem, st>>>(g);
          count_launch(1);
          return cudaGetLastError();
        };
        if (same) return go(k_dot_wide_tma<WIDE_B, U, true, CT>, 2 * WIDE_B * U * 16 + U * QWA_COL_WORDS * WIDE_B * 4 + 64);
        return go(k_dot_wide_tma<WIDE_B, U, false, CT>, 4 * WIDE_B * U * 16 + U * QWA_COL_WORDS * WIDE_B * 4 + 64);
      }
      int grid = WIDE_GRID;
      const int64_t need = (g.n + WIDE_B - 1) / WIDE_B;
      if (need < grid) grid = (int)need;
      if (same) k_dot_wide_l1<WIDE_B, 4, true, WIDE_GRID / 148><<<grid, WIDE_B, 0, st>>>(g);
      else k_dot_wide_l1<WIDE_B, 4, false, WIDE_GRID / 148><<<grid, WIDE_B, 0, st>>>(g);
      count_launch(1);
      return cudaGetLastError();
    }
    count_launch(2);
    return cudaGetLastError();
  }
  const bool unit = (g.incx == 1 && g.incy == 1);
  const int T = g.T < 1 ? 1 : g.T;
  int nchunks;
  int64_t chunk;
  if (g.n < 500) { nchunks = 1; chunk = g.n; }         /* PARALLEL_THRESHOLD, core/constants.hpp:15 */
  else { nchunks = T; chunk = g.n / T; }
  const int lanes = unit ? 2 : 1;
  const int64_t nth = (int64_t)nchunks * lanes;
  const int B = 64;
  k_dot_ref_chains<<<(unsigned)((nth + B - 1) / B), B, 0, st>>>(g, chunk, nchunks, lanes);
  const q128 *parts = g.work;
  if (unit) {
    k_dot_ref_lanes<<<(unsigned)((nchunks + B - 1) / B), B, 0, st>>>(g, chunk, nchunks);
    parts = g.work + 2 * (int64_t)nchunks;
    count_launch();
  }
  k_fold<<<1, 32, 0, st>>>(parts, nchunks, g.n < 500 ? 1 : 0, g.do_sqrt, g.result);
  count_launch(2);
  return cudaGetLastError();
}

/* partials of `nchunks` reference-order chunks of a local shard; work must hold 3*nchunks quads */
cudaError_t launch_dot_partials(const DotArgs &a, int64_t chunk, int nchunks, q128 *partials, cudaStream_t st)
{
  DotArgs g = a;
  if (nchunks <= 0) return cudaSuccess;
  const bool unit = (g.incx == 1 && g.incy == 1);
  const int lanes = unit ? 2 : 1;
  const int64_t nth = (int64_t)nchunks * lanes;
  const int B = 64;
  k_dot_ref_chains<<<(unsigned)((nth + B - 1) / B), B, 0, st>>>(g, chunk, nchunks, lanes);
  count_launch();
  const q128 *src = g.work;
  if (unit) {
    k_dot_ref_lanes<<<(unsigned)((nchunks + B - 1) / B), B, 0, st>>>(g, chunk, nchunks);
    count_launch();
    src = g.work + 2 * (int64_t)nchunks;
  }
  cudaError_t e = cudaMemcpyAsync(partials, src, (size_t)nchunks * 16, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_fold(int64_t count, const q128 *partials, int do_sqrt, q128 *result, cudaStream_t st)
{
  k_fold<<<1, 32, 0, st>>>(partials, count, 0, do_sqrt, result);
  count_launch();
  return cudaGetLastError();
}

/* ------------------------------------------------------------------ axpy */
/* y_i = fma(alpha, x_i, y_i) (level1.hpp:140-223); order-free, so one kernel serves every mode */
/* the loads of the next element are issued before the FMA of this one (the FMA is ~280 dependent instructions: without the
 * prefetch every warp waits out a full memory round trip per element) */
__global__ void __launch_bounds__(256, 4) k_axpy(int64_t n, q128 alpha, const q128 *x, int64_t incx, q128 *y, int64_t incy)
{
  const qop al = qop_load(alpha);
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  q128 xv = ldg128_l1(x + i * incx), yv = y[i * incy];
  for (; i < n; i += step) {
    const int64_t in = i + step;
    q128 xn = xv, yn = yv;
    if (in < n) { xn = ldg128_l1(x + in * incx); yn = y[in * incy]; }
    qacc acc = qacc_from(yv);
    qacc_fma(acc, al, qop_load(xv));
    y[i * incy] = qacc_pack(acc);
    xv = xn; yv = yn;
  }
}

cudaError_t launch_axpy(int64_t n, q128 alpha, const q128 *x, int64_t incx, q128 *y, int64_t incy, cudaStream_t st)
{
  if (n <= 0) return cudaSuccess;
  const int B = 256;
  int64_t grid = (n + B - 1) / B;
  if (grid > 148 * 8) grid = 148 * 8;
  k_axpy<<<(unsigned)grid, B, 0, st>>>(n, alpha, x, incx, y, incy);
  count_launch();
  return cudaGetLastError();
}

} // namespace qb

/*
 * qb_level3.cu — binary128 GEMM on the SM integer pipes.
 *
 * Replaces QuadBLAS::gemm and everything under it
 * (/root/reference/include/quadblas/algorithms/level3.hpp:215-336 driver, :128-185 macro kernel,
 * :20-126 micro kernels, :187-213 gemm_simple; blocking from detail/blocking.hpp:21-66).
 *
 * Reference order (what the bits depend on): per C element, k is cut into
 * panels of kc (126); inside a panel s = fma chain from +0 in ascending l; panels are folded with
 * C = fma(alpha, s, mul(q == 0 ? beta : 1, C)).  mc/nc/MR/NR and the thread count do not matter,
 * so the whole M x N plane is parallel and only the k order is kept.
 *
 * Kernel shape: a CTA owns a BM x BN tile of C; A/B k-slices are staged in shared memory; each
 * thread owns TM x TN accumulators kept UNPACKED in registers and steps them with the chain FMA
 * of q128_chain.cuh.  One qFMA costs ~140-170 integer instructions, so operand traffic
 * (1 LDS.128 per several hundred instructions) is irrelevant: the kernel is bound by the dispatch
 * clocks of that instruction stream, not by HBM, L2 or shared memory.
 * Two kernels, same bits: k_gemm (first version: raw quads staged, qacc_fma with an inlined slow-path
 * call per step) and k_gemm_nb (the default: operands staged decoded, branch-free qacc_fma_nb, declined
 * steps redone out of line); qb_set_ref_gemm_kernel / QBLAS_GEMM_KERNEL choose.
 */
#include "qb_internal.h"
#include "q128_chain.cuh"

namespace qb {

/* C = fma(alpha, s, bq * C) with the reference's epilogue (level3.hpp:102-109 / :206-210). */
__device__ __noinline__ q128 gemm_fold(q128 alpha, q128 s, q128 bq, q128 c, int apply_b)
{
  q128 t = apply_b ? q_mul(bq, c) : c; /* mul(1, c) == c exactly; skipping it changes no bit */
  return q_fma(alpha, s, t);
}

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
k_gemm(GemmArgs g)
{
  constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static_assert(NT % 32 == 0, "scratch columns need a multiple of 32 threads");
  __shared__ q128 sA[BK][BM];
  __shared__ q128 sB[BK][BN];
  __shared__ unsigned char zA[BK][BM], zB[BK][BN]; /* trailing-zero counts of the mantissas (jam test) */
  __shared__ uint32_t scr[12 * NT];                /* per-thread scratch columns, [word][thread] */

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
  const q128 fill = q_one(); /* out-of-range rows/cols compute on 1.0 so they stay on the fast path */
  for (int w = tid; w < 12 * NT; w += NT) scr[w] = 0u; /* words 6..11 of every column stay zero */
  qscratch sc;
  sc.col = scr + tid;
  sc.stride = NT;

  qacc acc[TM][TN];
  q128 cv[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      acc[a][b] = qacc_zero();
      const int64_t i = i0 + ty + a * TY, j = j0 + tx + b * TX;
      cv[a][b] = (i < g.m && j < g.n) ? g.C[i * g.sci + j * g.scj] : q_zero(0);
    }

  int64_t pc = 0;     /* position inside the current k-panel */
  int first = 1;      /* first panel uses beta, later ones 1 (level3.hpp:299) */
  const bool a_l_contig = (g.sal == 1);
  const bool b_j_contig = (g.sbj == 1);

  for (int64_t l0 = 0; l0 < g.k; l0 += BK) {
    /* ---- stage A[i0:i0+BM, l0:l0+BK] and B[l0:l0+BK, j0:j0+BN] ---- */
    for (int idx = tid; idx < BM * BK; idx += NT) {
      int i, l;
      if (a_l_contig) { l = idx % BK; i = idx / BK; } else { i = idx % BM; l = idx / BM; }
      const int64_t gi = i0 + i, gl = l0 + l;
      const q128 v = (gi < g.m && gl < g.k) ? g.A[gi * g.sai + gl * g.sal] : fill;
      sA[l][i] = v;
      zA[l][i] = (unsigned char)qop_tz(qop_load(v));
    }
    for (int idx = tid; idx < BN * BK; idx += NT) {
      int j, l;
      if (b_j_contig) { j = idx % BN; l = idx / BN; } else { l = idx % BK; j = idx / BK; }
      const int64_t gj = j0 + j, gl = l0 + l;
      const q128 v = (gj < g.n && gl < g.k) ? g.B[gl * g.sbl + gj * g.sbj] : fill;
      sB[l][j] = v;
      zB[l][j] = (unsigned char)qop_tz(qop_load(v));
    }
    __syncthreads();

    const int lim = (int)((g.k - l0) < BK ? (g.k - l0) : BK);
    for (int l = 0; l < lim; ++l) {
      qop a[TM], b[TN];
      uint32_t za[TM], zb[TN];
#pragma unroll
      for (int x = 0; x < TM; ++x) { a[x] = qop_load(sA[l][ty + x * TY]); za[x] = zA[l][ty + x * TY]; }
#pragma unroll
      for (int x = 0; x < TN; ++x) { b[x] = qop_load(sB[l][tx + x * TX]); zb[x] = zB[l][tx + x * TX]; }
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TN; ++y) qacc_fma_sc(acc[x][y], a[x], b[y], sc, za[x] + zb[y]);
      if (++pc == g.kc) {
#pragma unroll
        for (int x = 0; x < TM; ++x)
#pragma unroll
          for (int y = 0; y < TN; ++y) {
            cv[x][y] = gemm_fold(g.alpha, qacc_pack(acc[x][y]), g.beta, cv[x][y], first);
            acc[x][y] = qacc_zero();
          }
        pc = 0;
        first = 0;
      }
    }
    __syncthreads();
  }
  if (pc > 0) {
#pragma unroll
    for (int x = 0; x < TM; ++x)
#pragma unroll
      for (int y = 0; y < TN; ++y) cv[x][y] = gemm_fold(g.alpha, qacc_pack(acc[x][y]), g.beta, cv[x][y], first);
  }
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int64_t i = i0 + ty + a * TY, j = j0 + tx + b * TX;
      if (i < g.m && j < g.n) g.C[i * g.sci + j * g.scj] = cv[a][b];
    }
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * k_gemm_nb: the same tiling around the branch-free step qacc_fma_nb (q128_chain.cuh).
 *
 * The tiles are staged DECODED: four mantissa words with the implicit bit and one meta word per element (exponent, sign, class,
 * trailing zeros: qstage), so the 32 threads that consume an element do not each repeat its decoding.  The TM x TN steps of one k
 * index are independent and free of branches; the steps that declined (a rare cancellation of more than 13 bits, a product far above
 * the accumulator, an Inf / NaN / subnormal) are redone out of line by the generic q_fma from the same staged operands — same bits.
 * C is not held in registers during the k loop: it is read when a panel is folded (level3.hpp:102-109) and written back. */
__device__ __noinline__ void gemm_redo_step(uint32_t *w)   /* w[0..3] = a, w[4..7] = b, w[8..11] = s  ->  w[8..11] = fma(a, b, s) */
{
  q128 a, b, c;
  a.lo = ((uint64_t)w[1] << 32) | w[0]; a.hi = ((uint64_t)w[3] << 32) | w[2];
  b.lo = ((uint64_t)w[5] << 32) | w[4]; b.hi = ((uint64_t)w[7] << 32) | w[6];
  c.lo = ((uint64_t)w[9] << 32) | w[8]; c.hi = ((uint64_t)w[11] << 32) | w[10];
  const q128 r = q_fma(a, b, c);
  w[8] = (uint32_t)r.lo; w[9] = (uint32_t)(r.lo >> 32); w[10] = (uint32_t)r.hi; w[11] = (uint32_t)(r.hi >> 32);
}

__device__ __forceinline__ void gemm_redo(qacc2 &S, const uint4 &a, uint32_t ma, const uint4 &b, uint32_t mb)
{
  uint32_t w[12];
  qstaged A, B;
  A.m0 = a.x; A.m1 = a.y; A.m2 = a.z; A.m3 = a.w; A.meta = ma;
  B.m0 = b.x; B.m1 = b.y; B.m2 = b.z; B.m3 = b.w; B.meta = mb;
  const q128 pa = qstaged_pack(A), pb = qstaged_pack(B), pc = qacc2_pack(S);
  w[0] = (uint32_t)pa.lo; w[1] = (uint32_t)(pa.lo >> 32); w[2] = (uint32_t)pa.hi; w[3] = (uint32_t)(pa.hi >> 32);
  w[4] = (uint32_t)pb.lo; w[5] = (uint32_t)(pb.lo >> 32); w[6] = (uint32_t)pb.hi; w[7] = (uint32_t)(pb.hi >> 32);
  w[8] = (uint32_t)pc.lo; w[9] = (uint32_t)(pc.lo >> 32); w[10] = (uint32_t)pc.hi; w[11] = (uint32_t)(pc.hi >> 32);
  gemm_redo_step(w);
  q128 r;
  r.lo = ((uint64_t)w[9] << 32) | w[8]; r.hi = ((uint64_t)w[11] << 32) | w[10];
  S = qacc2_from(r);
}

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), 2)
k_gemm_nb(GemmArgs g)
{
  constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static_assert(NT % 32 == 0, "scratch columns need a multiple of 32 threads");
  __shared__ uint4 sA[BK][BM];
  __shared__ uint4 sB[BK][BN];
  __shared__ uint32_t eA[BK][BM], eB[BK][BN];
  __shared__ uint32_t scr[12 * NT];                /* per-thread scratch columns, [word][thread] */

  const int tid = threadIdx.x;
  const int tx = tid % TX, ty = tid / TX;
  const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
  const q128 fill = q_one(); /* out-of-range rows/cols compute on 1.0 */
  for (int w = tid; w < 12 * NT; w += NT) scr[w] = 0u; /* words 6..11 of every column stay zero */
  qnbctx sc;
  sc.col = scr + tid;
  sc.stride = NT;
  sc.zero = (uint32_t)((uint64_t)g.k >> 63);   /* 0 (k > 0), but not a constant the compiler can fold */

  qacc2 acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = qacc2_zero();

  auto fold_all = [&](int first) {
#pragma unroll
    for (int x = 0; x < TM; ++x)
#pragma unroll
      for (int y = 0; y < TN; ++y) {
        const int64_t i = i0 + ty + x * TY, j = j0 + tx + y * TX;
        if (i < g.m && j < g.n) {
          q128 *cp = g.C + i * g.sci + j * g.scj;
          *cp = gemm_fold(g.alpha, qacc2_pack(acc[x][y]), g.beta, *cp, first);
        }
        acc[x][y] = qacc2_zero();
      }
  };

  int64_t pc = 0;     /* position inside the current k-panel */
  int first = 1;      /* first panel uses beta, later ones 1 (level3.hpp:299) */
  const bool a_l_contig = (g.sal == 1);
  const bool b_j_contig = (g.sbj == 1);

  for (int64_t l0 = 0; l0 < g.k; l0 += BK) {
    for (int idx = tid; idx < BM * BK; idx += NT) {
      int i, l;
      if (a_l_contig) { l = idx % BK; i = idx / BK; } else { i = idx % BM; l = idx / BM; }
      const int64_t gi = i0 + i, gl = l0 + l;
      const qstaged v = qstage((gi < g.m && gl < g.k) ? g.A[gi * g.sai + gl * g.sal] : fill);
      sA[l][i] = make_uint4(v.m0, v.m1, v.m2, v.m3);
      eA[l][i] = v.meta;
    }
    for (int idx = tid; idx < BN * BK; idx += NT) {
      int j, l;
      if (b_j_contig) { j = idx % BN; l = idx / BN; } else { l = idx % BK; j = idx / BK; }
      const int64_t gj = j0 + j, gl = l0 + l;
      const qstaged v = qstage((gj < g.n && gl < g.k) ? g.B[gl * g.sbl + gj * g.sbj] : fill);
      sB[l][j] = make_uint4(v.m0, v.m1, v.m2, v.m3);
      eB[l][j] = v.meta;
    }
    __syncthreads();

    const int lim = (int)((g.k - l0) < BK ? (g.k - l0) : BK);
    for (int l = 0; l < lim; ++l) {
      uint4 a[TM], b[TN];
      uint32_t ma[TM], mb[TN];
#pragma unroll
      for (int x = 0; x < TM; ++x) { a[x] = sA[l][ty + x * TY]; ma[x] = eA[l][ty + x * TY]; }
#pragma unroll
      for (int x = 0; x < TN; ++x) { b[x] = sB[l][tx + x * TX]; mb[x] = eB[l][tx + x * TX]; }
      bool declined[TM][TN], any = false;
#pragma unroll
      for (int x = 0; x < TM; ++x)
#pragma unroll
        for (int y = 0; y < TN; ++y) {
          declined[x][y] = qacc_fma_nb(acc[x][y], a[x].x, a[x].y, a[x].z, a[x].w, b[y].x, b[y].y, b[y].z, b[y].w, ma[x] + mb[y], sc);
          any = any || declined[x][y];
        }
      if (any) {
#pragma unroll
        for (int x = 0; x < TM; ++x)
#pragma unroll
          for (int y = 0; y < TN; ++y)
            if (declined[x][y]) gemm_redo(acc[x][y], a[x], ma[x], b[y], mb[y]);
      }
      if (++pc == g.kc) {
        fold_all(first);
        pc = 0;
        first = 0;
      }
    }
    __syncthreads();
  }
  if (pc > 0) fold_all(first);
}

cudaError_t launch_gemm(const GemmArgs &a, int mode, cudaStream_t st)
{
  if (a.m == 0 || a.n == 0 || a.k == 0) return cudaSuccess; /* level3.hpp:221: C untouched */
  GemmArgs g = a;
  if (mode != 0 || g.kc <= 0) g.kc = (mode != 0) ? g.k : 126;
  constexpr int BM = 32, BN = 32, BK = 16, TM = 2, TN = 2;
  dim3 grid((unsigned)((g.n + BN - 1) / BN), (unsigned)((g.m + BM - 1) / BM));
  if (ref_gemm_kernel() != 0) k_gemm_nb<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g);
  else k_gemm<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g);   /* qb_set_ref_gemm_kernel(0) / QBLAS_GEMM_KERNEL=0: the first version, kept for the side-by-side */
  count_launch();
  return cudaGetLastError();
}

} // namespace qb

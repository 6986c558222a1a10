/*
 * qb_elem.cu — elementwise scalar-op kernels (parity probes for the arithmetic core against
 * Sleef_{fma,mul,add,sqrt}q1_u05 and the casts) and the register-resident qFMA microbenchmark
 * that gives the empirical integer-pipe roofline for qgemm.
 */
#include "qb_internal.h"
#include "q128_chain.cuh"
#include "qwide.cuh"

namespace qb {

__global__ void k_elementwise(int op, int64_t n, const q128 *a, const q128 *b, const q128 *c, q128 *out)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    q128 r;
    switch (op) {
    case 0: r = q_fma(a[i], b[i], c[i]); break;
    case 1: r = q_fma_fast(a[i], b[i], c[i]); break;
    case 2: r = q_mul(a[i], b[i]); break;
    case 3: r = q_add(a[i], b[i]); break;
    case 4: r = q_sqrt(a[i]); break;
    default: r = q_from_double_bits(q_to_double_bits(a[i])); break;
    }
    out[i] = r;
  }
}

cudaError_t launch_elementwise(int op, int64_t n, const q128 *a, const q128 *b, const q128 *c, q128 *out, cudaStream_t st)
{
  if (n <= 0) return cudaSuccess;
  const int B = 128;
  int64_t grid = (n + B - 1) / B;
  if (grid > 148 * 16) grid = 148 * 16;
  k_elementwise<<<(unsigned)grid, B, 0, st>>>(op, n, a, b, c, out);
  count_launch();
  return cudaGetLastError();
}

/* ILP independent accumulators per thread, operands generated in registers (xorshift mantissas,
 * exponents near the bias, random signs) so that no global memory traffic is involved.  Every step
 * is a full correctly rounded FMA through the same qacc_fma / qacc_fma_sc the kernels use.
 * SC = scratch-column form (the one k_gemm runs). */
template <int ILP, bool SC, int NT>
__global__ void __launch_bounds__(NT) k_fma_microbench(int iters, q128 *sink)
{
  __shared__ uint32_t scr[SC ? 12 * NT : 1];
  qscratch sc;
  sc.col = scr + threadIdx.x;
  sc.stride = NT;
  if (SC) {
    for (int w = threadIdx.x; w < 12 * NT; w += NT) scr[w] = 0u;
    __syncthreads();
  }
  uint32_t s = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  auto next = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  qop a[ILP], b;
  uint32_t tza[ILP];
  qacc acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) {
    a[u].m0 = next() | 1u; a[u].m1 = next(); a[u].m2 = next(); a[u].m3 = (next() & 0xffffu) | 0x10000u;
    a[u].e = 16383 - (int)(next() & 3); a[u].s = next() & 1;
    tza[u] = 0;
    acc[u] = qacc_zero();
  }
  b.m0 = next() | 1u; b.m1 = next(); b.m2 = next(); b.m3 = (next() & 0xffffu) | 0x10000u; b.e = 16383; b.s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (SC) qacc_fma_sc(acc[u], a[u], b, sc, tza[u]);  /* odd mantissas: tz(a) + tz(b) = 0 */
      else qacc_fma(acc[u], a[u], b);
    }
    /* perturb the shared operand so that nothing is loop invariant (bit 0 stays set) */
    b.m0 += 0x9e3779b8u; b.m1 ^= b.m0; b.s ^= (b.m0 >> 7) & 1u;
  }
  q128 r = qacc_pack(acc[0]);
#pragma unroll
  for (int u = 1; u < ILP; ++u) { q128 t = qacc_pack(acc[u]); r.lo ^= t.lo; r.hi ^= t.hi; }
  if (r.lo == 0x1234567 && r.hi == 0x7654321) sink[0] = r; /* keep the result alive */
}

/* Same harness for the fast-mode window accumulate (qwide.cuh: qwa_fma on a scratch column, exactly what
 * the qdot / qnrm2 / qgemv kernels run per element): the register-resident ceiling of those kernels.
 * Every word of both operands changes every step, so that no partial product is loop invariant. */
template <int ILP, int NT>
__global__ void __launch_bounds__(NT) k_wide_microbench(int iters, q128 *sink)
{
  __shared__ uint32_t scr[ILP * QWA_COL_WORDS * NT];
  uint32_t s = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  auto next = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  qop a[ILP], b;
  qwacc acc[ILP];
  uint32_t bad = 0;
#pragma unroll
  for (int u = 0; u < ILP; ++u) {
    a[u].m0 = next() | 1u; a[u].m1 = next(); a[u].m2 = next(); a[u].m3 = (next() & 0xffffu) | 0x10000u;
    a[u].e = 16383 - (int)(next() & 3); a[u].s = next() & 1;
    acc[u] = qwa_zero();
    qwa_col_init(scr + u * QWA_COL_WORDS * NT + threadIdx.x, NT);
  }
  b.m0 = next() | 1u; b.m1 = next(); b.m2 = next(); b.m3 = (next() & 0xffffu) | 0x10000u; b.e = 16383; b.s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u)
      if (qwa_fma(acc[u], a[u], b, scr + u * QWA_COL_WORDS * NT + threadIdx.x, NT)) qwa_fma_rare(acc[u], qop_pack(a[u]), qop_pack(b), bad);
    b.m0 += 0x9e3779b8u; b.m1 ^= b.m0; b.m2 += b.m1 | 1u; b.m3 = ((b.m3 + (b.m2 >> 20)) & 0xffffu) | 0x10000u; b.s ^= (b.m0 >> 7) & 1u;
    b.e = 16383 - (int)((b.m0 >> 9) & 7u);   /* alignment shifts 0..10 bits, as U(-1,1) data give */
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      a[u].m0 ^= b.m1; a[u].m1 += b.m2; a[u].m2 ^= b.m0; a[u].m3 = ((a[u].m3 + (b.m1 >> 24)) & 0xffffu) | 0x10000u;
      a[u].s ^= (b.m1 >> 3) & 1u;
    }
  }
  q128 r = qw_finish(qwa_fold(acc[0]), bad);
#pragma unroll
  for (int u = 1; u < ILP; ++u) { q128 t = qw_finish(qwa_fold(acc[u]), bad); r.lo ^= t.lo; r.hi ^= t.hi; }
  if (r.lo == 0x1234567 && r.hi == 0x7654321) sink[0] = r;
}

/* variant = 100 * SC + ILP (ILP in {1,2,4}); 1000 + ILP = window accumulate; threads must be 128 or 256 */
cudaError_t launch_fma_microbench(int variant, int blocks, int threads, int iters, q128 *sink, int64_t *n_fma, cudaStream_t st)
{
  if (variant >= 1000) {
    const int il = variant - 1000;
    if ((il != 1 && il != 2 && il != 4) || (threads != 128 && threads != 256)) return cudaErrorInvalidValue;
    if (threads == 128) { if (il == 1) k_wide_microbench<1, 128><<<blocks, 128, 0, st>>>(iters, sink); else if (il == 2) k_wide_microbench<2, 128><<<blocks, 128, 0, st>>>(iters, sink); else k_wide_microbench<4, 128><<<blocks, 128, 0, st>>>(iters, sink); }
    else { if (il == 1) k_wide_microbench<1, 256><<<blocks, 256, 0, st>>>(iters, sink); else if (il == 2) k_wide_microbench<2, 256><<<blocks, 256, 0, st>>>(iters, sink); else k_wide_microbench<4, 256><<<blocks, 256, 0, st>>>(iters, sink); }
    if (n_fma) *n_fma = (int64_t)blocks * threads * iters * il;
    count_launch();
    return cudaGetLastError();
  }
  const int ilp = variant % 100, scv = variant / 100;
  if ((ilp != 1 && ilp != 2 && ilp != 4) || (threads != 128 && threads != 256)) return cudaErrorInvalidValue;
#define QB_MB(I, S, T) k_fma_microbench<I, S, T><<<blocks, T, 0, st>>>(iters, sink)
  if (threads == 128) {
    if (!scv) { if (ilp == 1) QB_MB(1, false, 128); else if (ilp == 2) QB_MB(2, false, 128); else QB_MB(4, false, 128); }
    else      { if (ilp == 1) QB_MB(1, true, 128);  else if (ilp == 2) QB_MB(2, true, 128);  else QB_MB(4, true, 128); }
  } else {
    if (!scv) { if (ilp == 1) QB_MB(1, false, 256); else if (ilp == 2) QB_MB(2, false, 256); else QB_MB(4, false, 256); }
    else      { if (ilp == 1) QB_MB(1, true, 256);  else if (ilp == 2) QB_MB(2, true, 256);  else QB_MB(4, true, 256); }
  }
#undef QB_MB
  if (n_fma) *n_fma = (int64_t)blocks * threads * iters * ilp;
  count_launch();
  return cudaGetLastError();
}

} // namespace qb

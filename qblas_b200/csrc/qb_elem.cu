/*
 * qb_elem.cu — elementwise scalar-op kernels (parity probes for the arithmetic core against
 * Sleef_{fma,mul,add,sqrt}q1_u05 and the casts) and the register-resident qFMA microbenchmark
 * that gives the empirical integer-pipe roofline for qgemm.
 */
#include "qb_internal.h"
#include "q128_chain.cuh"

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
 * exponents near the bias, random signs) so that no memory traffic is involved.  Every step is a
 * full correctly rounded FMA through the same qacc_fma the kernels use. */
template <int ILP>
__global__ void k_fma_microbench(int iters, q128 *sink)
{
  uint32_t s = 0x9e3779b9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
  auto next = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  qop a[ILP], b;
  qacc acc[ILP];
#pragma unroll
  for (int u = 0; u < ILP; ++u) {
    a[u].m0 = next(); a[u].m1 = next(); a[u].m2 = next(); a[u].m3 = (next() & 0xffffu) | 0x10000u;
    a[u].e = 16383 - (int)(next() & 3); a[u].s = next() & 1;
    acc[u] = qacc_zero();
  }
  b.m0 = next(); b.m1 = next(); b.m2 = next(); b.m3 = (next() & 0xffffu) | 0x10000u; b.e = 16383; b.s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < ILP; ++u) qacc_fma(acc[u], a[u], b);
    /* perturb the shared operand so that nothing is loop invariant */
    b.m0 += 0x9e3779b9u; b.m1 ^= b.m0; b.s ^= (b.m0 >> 7) & 1u;
  }
  q128 r = qacc_pack(acc[0]);
#pragma unroll
  for (int u = 1; u < ILP; ++u) { q128 t = qacc_pack(acc[u]); r.lo ^= t.lo; r.hi ^= t.hi; }
  if (r.lo == 0x1234567 && r.hi == 0x7654321) sink[0] = r; /* keep the result alive */
}

cudaError_t launch_fma_microbench(int variant, int blocks, int threads, int iters, q128 *sink, int64_t *n_fma, cudaStream_t st)
{
  int ilp = 1;
  switch (variant) {
  case 1: k_fma_microbench<1><<<blocks, threads, 0, st>>>(iters, sink); ilp = 1; break;
  case 2: k_fma_microbench<2><<<blocks, threads, 0, st>>>(iters, sink); ilp = 2; break;
  default: k_fma_microbench<4><<<blocks, threads, 0, st>>>(iters, sink); ilp = 4; break;
  }
  if (n_fma) *n_fma = (int64_t)blocks * threads * iters * ilp;
  count_launch();
  return cudaGetLastError();
}

} // namespace qb

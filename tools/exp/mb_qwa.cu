// Standalone microbenchmark of the fast-mode window accumulate (qwide.cuh: qwa_fma) with knobs that
// remove one stage at a time, to find which SM resource bounds it.  Development tool, not shipped.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I qblas_b200/csrc -I include tools/exp/mb_qwa.cu -o tools/exp/_mb_qwa
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "q128.cuh"
#include "q128_chain.cuh"
#include "qwide.cuh"
using namespace qb;

// MODE bit0: no shared-memory round trip; bit1: no multiply; bit2: no accumulate
template <int MODE>
__device__ __forceinline__ bool step(qwacc &S, const qop &A, const qop &B, uint32_t *col, uint32_t stride)
{
  if (MODE == 0) return qwa_fma(S, A, B, col, stride);
  const bool normal = ((uint32_t)A.e - 1u < 0x7ffeu) && ((uint32_t)B.e - 1u < 0x7ffeu);
  const int32_t d = normal ? S.E - A.e - B.e : -1;
  const bool rare = d < 0;
  uint32_t p2, p3, p4, p5, p6, p7;
  if (MODE & 2) { p2 = A.m0 ^ B.m1; p3 = A.m1 ^ B.m2; p4 = A.m2 + B.m0; p5 = A.m3 ^ B.m3; p6 = A.m0 + B.m2; p7 = A.m1 ^ B.m0; }
  else mul4x4_top6(A.m0, A.m1, A.m2, A.m3, B.m0, B.m1, B.m2, B.m3, p2, p3, p4, p5, p6, p7);
  uint32_t du = (uint32_t)d;
  du = du > 223u ? 223u : du;
  const uint32_t wq = du >> 5;
  const uint32_t r = (du & 31u) + 1u;
  const uint32_t M = 0x80000000u >> (du & 31u);
  const uint32_t mask = rare ? 0u : (0u - (A.s ^ B.s));
  uint32_t t1, t2, t3, t4, t5, t6;
  if (MODE & 1) { t1 = p2 ^ mask; t2 = p3 ^ mask; t3 = p4 ^ mask; t4 = p5 ^ mask; t5 = p6 ^ mask; t6 = (p7 + wq) ^ mask; }
  else {
    col[0] = p2; col[stride] = p3; col[2 * stride] = p4; col[3 * stride] = p5; col[4 * stride] = p6; col[5 * stride] = p7;
    const uint32_t *q = col + wq * stride;
    t1 = q[0] ^ mask; t2 = q[stride] ^ mask; t3 = q[2 * stride] ^ mask; t4 = q[3 * stride] ^ mask; t5 = q[4 * stride] ^ mask; t6 = q[5 * stride] ^ mask;
  }
  if (MODE & 4) {
    S.e01 ^= ((uint64_t)t2 << 32) | t1; S.e23 += ((uint64_t)t4 << 32) | t3; S.e45 ^= ((uint64_t)t6 << 32) | t5; S.v0 += M + r;
    return rare;
  }
  const uint32_t t1s = __funnelshift_rc(t1, 0u, r);
  asm("{\n\t"
      ".reg .u32 cy, a0, a1, a2, a3, a4, a5, b1, b2, b3, b4;\n\t"
      "mov.b64 {a0, a1}, %0;\n\t" "mov.b64 {a2, a3}, %1;\n\t" "mov.b64 {a4, a5}, %2;\n\t"
      "mov.b64 {b1, b2}, %4;\n\t" "mov.b64 {b3, b4}, %5;\n\t"
      "add.cc.u32      cy, %13, %13;\n\t"
      "madc.lo.cc.u32  a0, %8, %14, a0;\n\t"   "madc.hi.cc.u32 a1, %8, %14, a1;\n\t"
      "madc.lo.cc.u32  a2, %10, %14, a2;\n\t"  "madc.hi.cc.u32 a3, %10, %14, a3;\n\t"
      "madc.lo.cc.u32  a4, %12, %14, a4;\n\t"  "madc.hi.u32    a5, %12, %14, a5;\n\t"
      "add.cc.u32      %3, %3, %7;\n\t"
      "madc.lo.cc.u32  b1, %9, %14, b1;\n\t"   "madc.hi.cc.u32 b2, %9, %14, b2;\n\t"
      "madc.lo.cc.u32  b3, %11, %14, b3;\n\t"  "madc.hi.cc.u32 b4, %11, %14, b4;\n\t"
      "madc.lo.u32     %6, %13, %14, %6;\n\t"
      "mov.b64 %0, {a0, a1};\n\t" "mov.b64 %1, {a2, a3};\n\t" "mov.b64 %2, {a4, a5};\n\t"
      "mov.b64 %4, {b1, b2};\n\t" "mov.b64 %5, {b3, b4};\n\t"
      "}"
      : "+l"(S.e01), "+l"(S.e23), "+l"(S.e45), "+r"(S.v0), "+l"(S.v12), "+l"(S.v34), "+r"(S.v5)
      : "r"(t1s), "r"(t2), "r"(t3), "r"(t4), "r"(t5), "r"(t6), "r"(mask), "r"(M));
  return rare;
}

template <int ILP, int NT, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB) k_mb(int iters, q128 *sink)
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
    acc[u].E = 2 * 16383 + 20;
    qwa_col_init(scr + u * QWA_COL_WORDS * NT + threadIdx.x, NT);
  }
  b.m0 = next() | 1u; b.m1 = next(); b.m2 = next(); b.m3 = (next() & 0xffffu) | 0x10000u; b.e = 16383; b.s = 0;
  for (int it = 0; it < iters; ++it) {
    bool rare = false;
#pragma unroll
    for (int u = 0; u < ILP; ++u) rare |= step<MODE>(acc[u], a[u], b, scr + u * QWA_COL_WORDS * NT + threadIdx.x, NT);
    if (rare) bad++;
    /* every operand word changes every step (nothing hoistable): xorshift on b, a rotated by b */
    b.m0 += 0x9e3779b8u; b.m1 ^= b.m0; b.m2 += b.m1 | 1u; b.m3 = ((b.m3 + (b.m2 >> 20)) & 0xffffu) | 0x10000u; b.s ^= (b.m0 >> 7) & 1u;
    b.e = 16383 - (int)((b.m0 >> 9) & 7u);
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      a[u].m0 ^= b.m1; a[u].m1 += b.m2; a[u].m2 ^= b.m0; a[u].m3 = ((a[u].m3 + (b.m1 >> 24)) & 0xffffu) | 0x10000u; a[u].s ^= (b.m1 >> 3) & 1u;
    }
  }
  qwide f = qwa_fold(acc[0]);
  q128 r = qw_finish(f, bad);
#pragma unroll
  for (int u = 1; u < ILP; ++u) { q128 t = qw_finish(qwa_fold(acc[u]), bad); r.lo ^= t.lo; r.hi ^= t.hi; }
  if (r.lo == 0x1234567 && r.hi == 0x7654321) sink[0] = r;
}

template <int ILP, int NT, int MINB, int MODE>
static void run(const char *name, q128 *sink)
{
  const int iters = 4000, blocks = 148 * MINB;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_mb<ILP, NT, MINB, MODE><<<blocks, NT>>>(64, sink);
  cudaEventRecord(e0);
  k_mb<ILP, NT, MINB, MODE><<<blocks, NT>>>(iters, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * NT * iters * ILP;
  cudaError_t err = cudaGetLastError();
  printf("%-34s ILP=%d NT=%d CTAs/SM=%d warps/SM=%2d : %7.1f Gacc/s  (%s)\n", name, ILP, NT, MINB, MINB * NT / 32, n / ms / 1e6, cudaGetErrorString(err));
}

int main()
{
  q128 *sink; cudaMalloc(&sink, 64);
  run<2, 128, 5, 0>("full", sink);
  run<4, 128, 3, 0>("full", sink);
  run<1, 128, 8, 0>("full", sink);
  run<1, 256, 4, 0>("full", sink);
  run<1, 128, 4, 0>("full", sink);
  run<1, 128, 2, 0>("full", sink);
  run<1, 128, 8, 1>("no smem", sink);
  run<1, 128, 8, 2>("no mul", sink);
  run<1, 128, 8, 4>("no accumulate", sink);
  run<1, 128, 8, 3>("no smem, no mul", sink);
  run<1, 128, 8, 5>("no smem, no accumulate", sink);
  run<1, 128, 8, 6>("no mul, no accumulate", sink);
  run<1, 128, 8, 7>("only operand update/align control", sink);
  return 0;
}

// Is the FP64 pipe of sm_100a (B200) a useful multiplier for the fast-mode accumulate?  Measures DFMA / DADD issue cost per SM
// sub-partition, alone and mixed with ALU work, and a register-resident prototype of the sliced accumulate (22-bit slices of the
// binary128 significand as doubles, 21 DFMAs per element into 6 exact column sums).  Development tool.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb_f64 tools/exp/mb_f64.cu && /tmp/mb_f64
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int KIND>
__global__ void __launch_bounds__(256, 4) k(int iters, double *out, double seed)
{
  double a[8], b[8]; uint32_t u[8], v[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = a[i] * 0.5 + 1.0; u[i] = (uint32_t)(threadIdx.x * 77 + i); v[i] = u[i] ^ 0x9e3779b9u; }
  const double m = seed * 1e-3 + 1.0;
  const uint32_t mu = (uint32_t)seed | 1u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (KIND == 0) {
#define X(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 1) {
#define X(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 2) {   // DFMA + LOP3
#define X(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(b[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(mu));
        REP8(X)
#undef X
      } else if (KIND == 3) {   // DFMA + 2 LOP3
#define X(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(b[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(mu)); \
             asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(u[i]), "r"(mu));
        REP8(X)
#undef X
      } else if (KIND == 4) {   // DFMA + IMAD.WIDE
#define X(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(b[i])); \
             asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(mu), "r"(v[i]));
        REP8(X)
#undef X
      } else if (KIND == 5) {   // LOP3 + SHF (ALU only, for reference)
#define X(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(mu)); \
             asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(u[i]), "r"(mu));
        REP8(X)
#undef X
      }
    }
  }
  double r = 0; uint32_t q = 0;
  for (int i = 0; i < 8; ++i) { r += a[i] + b[i]; q ^= u[i] ^ v[i]; }
  if (r == 0.123 || q == 0x12345678u) out[0] = r;
}

template <int KIND>
static void run(const char *name, int ninstr, double *out)
{
  const int iters = 2000, blocks = 148 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<blocks, 256>>>(10, out, 1.25);
  cudaEventRecord(e0);
  k<KIND><<<blocks, 256>>>(iters, out, 1.25);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double wi = 8.0 * iters * 4 * ninstr;       // warp instructions per sub-partition (8 warps each)
  printf("%-36s %3d instr/rep : %6.2f clk per warp-instr per SMSP (at 1.92 GHz), %.3f ms  [%s]\n", name, ninstr, ms * 1e-3 * 1.92e9 / wi, ms,
         cudaGetErrorString(cudaGetLastError()));
}

// ---------------------------------------------------------------------------------------------------------------------------------
// prototype of the sliced accumulate: R rows per thread share the x slices of a column step (staged in a shared-memory column
// with 6 zero entries in front, read at a dynamic slice offset q), each row element is shifted by r < 22 bits, cut into 6 slices
// of 22 bits at fixed positions, converted with the 2^52 trick and multiplied into 6 column sums (21 DFMAs).
constexpr int SL = 22, NT = 128;

struct Cols { double c0, c1, c2, c3, c4, c5; };

__device__ __forceinline__ void setlo(double &P, uint32_t lo)
{
  asm("{.reg .b32 l, h; mov.b64 {l, h}, %0; mov.b64 %0, {%1, h};}" : "+d"(P) : "r"(lo));
}

struct Pairs { double p0, p1, p2, p3, p4, p5, pm; };   /* hi halves hold 0x43300000 for the whole loop: only the low halves are written */

__device__ __forceinline__ void step(Cols &C, Pairs &P, int32_t &dmax, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int32_t anc, int32_t ex, const double *col)
{
  const uint32_t e = (w3 >> 16) & 0x7fffu;
  uint32_t sh = (uint32_t)(anc - (int32_t)e);
  sh = sh > 153u ? 153u : sh;
  const uint32_t q = (sh * 2979u) >> 16;
  const uint32_t s = 21u - (sh - q * 22u);
  const uint32_t m3 = (w3 & 0xffffu) | 0x10000u;
  const uint32_t v0 = w0 << s, v1 = __funnelshift_l(w0, w1, s), v2 = __funnelshift_l(w1, w2, s), v3 = __funnelshift_l(w2, m3, s), v4 = m3 >> (32u - s);
  const uint32_t MK = (1u << SL) - 1u;
  const uint32_t sx = (uint32_t)((int32_t)w3 >> 31);
  setlo(P.p0, (__funnelshift_r(v3, v4, 16) ^ sx) & MK); setlo(P.p1, (__funnelshift_r(v2, v3, 26) ^ sx) & MK); setlo(P.p2, ((v2 >> 4) ^ sx) & MK);
  setlo(P.p3, (__funnelshift_r(v1, v2, 14) ^ sx) & MK); setlo(P.p4, (__funnelshift_r(v0, v1, 24) ^ sx) & MK); setlo(P.p5, ((v0 >> 2) ^ sx) & MK);
  setlo(P.pm, sx & MK);
  const double d0 = P.p0 - P.pm, d1 = P.p1 - P.pm, d2 = P.p2 - P.pm, d3 = P.p3 - P.pm, d4 = P.p4 - P.pm, d5 = P.p5 - P.pm;
  const double *xq = col + (6 - (int)q) * NT;
  const double x0 = xq[0], x1 = xq[NT], x2 = xq[2 * NT], x3 = xq[3 * NT], x4 = xq[4 * NT], x5 = xq[5 * NT];
  C.c0 = fma(d0, x0, C.c0);
  C.c1 = fma(d0, x1, fma(d1, x0, C.c1));
  C.c2 = fma(d0, x2, fma(d1, x1, fma(d2, x0, C.c2)));
  C.c3 = fma(d0, x3, fma(d1, x2, fma(d2, x1, fma(d3, x0, C.c3))));
  C.c4 = fma(d0, x4, fma(d1, x3, fma(d2, x2, fma(d3, x1, fma(d4, x0, C.c4)))));
  C.c5 = fma(d0, x5, fma(d1, x4, fma(d2, x3, fma(d3, x2, fma(d4, x1, fma(d5, x0, C.c5))))));
  dmax = max(dmax, (int32_t)e + ex);
}

template <int R>
__global__ void __launch_bounds__(NT, 4) k_proto(int iters, double *out, uint32_t seed)
{
  __shared__ double colsm[12 * NT];
  const int tid = threadIdx.x;
  for (int kx = 0; kx < 6; ++kx) colsm[kx * NT + tid] = 0.0;
  Cols C[R];
  Pairs P; P.p0 = P.p1 = P.p2 = P.p3 = P.p4 = P.p5 = P.pm = __hiloint2double(0x43300000, 0);
  int32_t dmax[R];
  uint32_t w[R][4];
  for (int r = 0; r < R; ++r) {
    C[r] = Cols{0, 0, 0, 0, 0, 0}; dmax[r] = -100000;
    for (int j = 0; j < 4; ++j) w[r][j] = seed * (tid * 4 + r * 17 + j + 1);
  }
  double xs[6];
  for (int j = 0; j < 6; ++j) xs[j] = (double)((seed * (tid + j + 3)) & 0x3fffff);
  const int32_t anc = 0x3fff;
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < 6; ++j) { colsm[(6 + j) * NT + tid] = xs[j]; xs[j] = (double)(((uint32_t)xs[j] * 2654435761u + it) & 0x3fffff); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      // next pseudo element: all words change; exponent within 0..40 below the anchor
      w[r][0] = w[r][0] * 1664525u + 1013904223u; w[r][1] ^= w[r][0] >> 3; w[r][2] += w[r][1];
      const uint32_t ee = (uint32_t)anc - ((w[r][0] >> 9) % 41u);
      w[r][3] = (w[r][2] & 0x8000ffffu) | (ee << 16);
      step(C[r], P, dmax[r], w[r][0], w[r][1], w[r][2], w[r][3], anc, 0x3fff, colsm + tid);
    }
    if ((it & 63) == 63) {
      for (int r = 0; r < R; ++r) { C[r].c0 *= 0x1p-40; C[r].c1 *= 0x1p-40; C[r].c2 *= 0x1p-40; C[r].c3 *= 0x1p-40; C[r].c4 *= 0x1p-40; C[r].c5 *= 0x1p-40; }
    }
  }
  double s = 0;
  for (int r = 0; r < R; ++r) s += C[r].c0 + C[r].c1 + C[r].c2 + C[r].c3 + C[r].c4 + C[r].c5 + dmax[r];
  if (s == 0.125) out[0] = s;
}


// variant B: unsigned slices converted with I2F.F64.U32; the sign of the element selects the +x or the -x copy of the slices in shared memory
__device__ __forceinline__ void stepB(Cols &C, int32_t &dmax, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, int32_t anc, int32_t ex, const double *col)
{
  const uint32_t e = (w3 >> 16) & 0x7fffu;
  uint32_t sh = (uint32_t)(anc - (int32_t)e);
  sh = sh > 153u ? 153u : sh;
  const uint32_t q = (sh * 2979u) >> 16;
  const uint32_t s = 21u - (sh - q * 22u);
  const uint32_t m3 = (w3 & 0xffffu) | 0x10000u;
  const uint32_t v0 = w0 << s, v1 = __funnelshift_l(w0, w1, s), v2 = __funnelshift_l(w1, w2, s), v3 = __funnelshift_l(w2, m3, s), v4 = m3 >> (32u - s);
  const uint32_t MK = (1u << SL) - 1u;
  const double d0 = (double)__funnelshift_r(v3, v4, 16), d1 = (double)(__funnelshift_r(v2, v3, 26) & MK), d2 = (double)((v2 >> 4) & MK),
               d3 = (double)(__funnelshift_r(v1, v2, 14) & MK), d4 = (double)(__funnelshift_r(v0, v1, 24) & MK), d5 = (double)((v0 >> 2) & MK);
  const double *xq = col + (6 * NT) - (int)q * NT + (w3 >> 31) * (12 * NT);
  const double x0 = xq[0], x1 = xq[NT], x2 = xq[2 * NT], x3 = xq[3 * NT], x4 = xq[4 * NT], x5 = xq[5 * NT];
  C.c0 = fma(d0, x0, C.c0);
  C.c1 = fma(d0, x1, fma(d1, x0, C.c1));
  C.c2 = fma(d0, x2, fma(d1, x1, fma(d2, x0, C.c2)));
  C.c3 = fma(d0, x3, fma(d1, x2, fma(d2, x1, fma(d3, x0, C.c3))));
  C.c4 = fma(d0, x4, fma(d1, x3, fma(d2, x2, fma(d3, x1, fma(d4, x0, C.c4)))));
  C.c5 = fma(d0, x5, fma(d1, x4, fma(d2, x3, fma(d3, x2, fma(d4, x1, fma(d5, x0, C.c5))))));
  dmax = max(dmax, (int32_t)e + ex);
}

template <int R>
__global__ void __launch_bounds__(NT, 4) k_protoB(int iters, double *out, uint32_t seed)
{
  __shared__ double colsm[24 * NT];
  const int tid = threadIdx.x;
  for (int kx = 0; kx < 6; ++kx) colsm[kx * NT + tid] = colsm[(12 + kx) * NT + tid] = 0.0;
  Cols C[R];
  int32_t dmax[R];
  uint32_t w[R][4];
  for (int r = 0; r < R; ++r) {
    C[r] = Cols{0, 0, 0, 0, 0, 0}; dmax[r] = -100000;
    for (int j = 0; j < 4; ++j) w[r][j] = seed * (tid * 4 + r * 17 + j + 1);
  }
  double xs[6];
  for (int j = 0; j < 6; ++j) xs[j] = (double)((seed * (tid + j + 3)) & 0x3fffff);
  const int32_t anc = 0x3fff;
  for (int it = 0; it < iters; ++it) {
    for (int j = 0; j < 6; ++j) { colsm[(6 + j) * NT + tid] = xs[j]; colsm[(18 + j) * NT + tid] = -xs[j]; xs[j] = (double)(((uint32_t)xs[j] * 2654435761u + it) & 0x3fffff); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      w[r][0] = w[r][0] * 1664525u + 1013904223u; w[r][1] ^= w[r][0] >> 3; w[r][2] += w[r][1];
      const uint32_t ee = (uint32_t)anc - ((w[r][0] >> 9) % 41u);
      w[r][3] = (w[r][2] & 0x8000ffffu) | (ee << 16);
      stepB(C[r], dmax[r], w[r][0], w[r][1], w[r][2], w[r][3], anc, 0x3fff, colsm + tid);
    }
    if ((it & 63) == 63) {
      for (int r = 0; r < R; ++r) { C[r].c0 *= 0x1p-40; C[r].c1 *= 0x1p-40; C[r].c2 *= 0x1p-40; C[r].c3 *= 0x1p-40; C[r].c4 *= 0x1p-40; C[r].c5 *= 0x1p-40; }
    }
  }
  double s = 0;
  for (int r = 0; r < R; ++r) s += C[r].c0 + C[r].c1 + C[r].c2 + C[r].c3 + C[r].c4 + C[r].c5 + dmax[r];
  if (s == 0.125) out[0] = s;
}

// the generator alone (what the prototypes spend outside the accumulate)
template <int R>
__global__ void __launch_bounds__(NT, 4) k_gen(int iters, double *out, uint32_t seed)
{
  const int tid = threadIdx.x;
  uint32_t w[R][4], acc = 0;
  for (int r = 0; r < R; ++r) for (int j = 0; j < 4; ++j) w[r][j] = seed * (tid * 4 + r * 17 + j + 1);
  const int32_t anc = 0x3fff;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      w[r][0] = w[r][0] * 1664525u + 1013904223u; w[r][1] ^= w[r][0] >> 3; w[r][2] += w[r][1];
      const uint32_t ee = (uint32_t)anc - ((w[r][0] >> 9) % 41u);
      w[r][3] = (w[r][2] & 0x8000ffffu) | (ee << 16);
      acc ^= w[r][3] + w[r][1];
    }
  }
  if (acc == 0x1234567u) out[0] = acc;
}

template <int R, int V>
static void run_proto(double *out, int blocks_per_sm)
{
  const int iters = 4000, blocks = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto launch = [&](int n) { if (V == 0) k_proto<R><<<blocks, NT>>>(n, out, 12345u); else if (V == 1) k_protoB<R><<<blocks, NT>>>(n, out, 12345u); else k_gen<R><<<blocks, NT>>>(n, out, 12345u); };
  launch(10);
  cudaEventRecord(e0);
  launch(iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double elems = (double)blocks * NT * iters * R;
  const double per_smsp_warp = (double)blocks_per_sm * iters * R;      // warp-elements per sub-partition (NT/32 = 4 warps per CTA, one per SMSP)
  printf("prototype %c R=%d, %d CTAs/SM: %.1f G elements/s = %.2f TB/s of A at 16 B; %.1f clk per warp-element per SMSP (at 1.92 GHz)  [%s]\n", "ABg"[V], R, blocks_per_sm,
         elems / ms * 1e-6, elems * 16 / ms * 1e-9, ms * 1e-3 * 1.92e9 / per_smsp_warp, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  double *out; cudaMalloc(&out, 64);
  run<0>("DFMA", 8, out);
  run<1>("DADD", 8, out);
  run<2>("DFMA + LOP3", 16, out);
  run<3>("DFMA + LOP3 + SHF", 24, out);
  run<4>("DFMA + IMAD", 16, out);
  run<5>("LOP3 + SHF", 16, out);
  run_proto<2, 0>(out, 4); run_proto<4, 0>(out, 4);
  run_proto<2, 1>(out, 4); run_proto<4, 1>(out, 4); run_proto<2, 1>(out, 6);
  run_proto<2, 2>(out, 4); run_proto<4, 2>(out, 4);
  return 0;
}

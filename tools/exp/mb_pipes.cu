// Calibrates the issue cost (clocks per warp instruction per SM sub-partition) of the integer instructions the
// binary128 kernels are made of, alone and mixed, on sm_100a.  Development tool.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int KIND>
__global__ void __launch_bounds__(256, 4) k(int iters, uint32_t *out, uint32_t seed)
{
  uint32_t a[8], b[8]; uint64_t w[8];
  __shared__ uint32_t sm[8 * 256];
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i + 1); b[i] = a[i] ^ 0x9e3779b9u; w[i] = ((uint64_t)a[i] << 32) | b[i]; sm[i * 256 + threadIdx.x] = a[i]; }
  const uint32_t m = seed | 1u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (KIND == 0) {       // IMAD lo
#define X(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 1) { // IMAD.WIDE
#define X(i) asm volatile("{.reg .u32 l,h; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(w[i]) : "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 14) { // IMAD.WIDE, c = 0 (pure 32x32->64 product)
#define X(i) asm volatile("{.reg .u32 l,h; mov.b64 {l,h}, %0; mul.wide.u32 %0, l, h;}" : "+l"(w[i]));
        REP8(X)
#undef X
      } else if (KIND == 15) { // 64-bit add (IADD3 + IADD3.X)
#define X(i) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) & 7]));
        REP8(X)
#undef X
      } else if (KIND == 16) { // IMAD.WIDE + LOP3 on independent data
#define X(i) asm volatile("{.reg .u32 l,h; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(w[i]) : "r"(m)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 17) { // IMAD.WIDE + IADD3 on independent data
#define X(i) asm volatile("{.reg .u32 l,h; mov.b64 {l,h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(w[i]) : "r"(m)); \
             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 18) { // IMAD + IADD3 on independent data
#define X(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(b[(i+1)&7])); \
             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i+1)&7]));
        REP8(X)
#undef X
      } else if (KIND == 19) { // LOP3 + IADD3 on independent data
#define X(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[(i+1)&7]), "r"(m)); \
             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i+1)&7]));
        REP8(X)
#undef X
      } else if (KIND == 2) { // IMAD.HI
#define X(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 3) { // wide via lo.cc/hi pair with carry chain across two
#define X(i) asm volatile("{.reg .u32 l,h; mov.b64 {l,h}, %0; mad.lo.cc.u32 l, %2, %3, l; madc.hi.cc.u32 h, %2, %3, h; mov.b64 %0, {l,h};\n\t" \
                          " mov.b64 {l,h}, %1; madc.lo.cc.u32 l, %4, %3, l; madc.hi.u32 h, %4, %3, h; mov.b64 %1, {l,h};}" : "+l"(w[i]), "+l"(w[(i + 4) & 7]) : "r"(a[i]), "r"(m), "r"(b[i]) : "memory");
        X(0) X(1) X(2) X(3)
#undef X
      } else if (KIND == 4) { // IADD3
#define X(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        REP8(X)
#undef X
      } else if (KIND == 5) { // LOP3
#define X(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 6) { // SHF
#define X(i) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 7) { // 1 WIDE + 2 LOP3
#define X(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(m)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 8) { // 1 WIDE + 1 IMAD
#define X(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m)); \
             asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(m), "r"(a[i]));
        REP8(X)
#undef X
      } else if (KIND == 9) { // STS + LDS 32-bit
#define X(i) asm volatile("st.shared.u32 [%1], %0;" :: "r"(a[i]), "r"((uint32_t)__cvta_generic_to_shared(&sm[i * 256 + threadIdx.x])) : "memory"); \
             asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b[i]) : "r"((uint32_t)__cvta_generic_to_shared(&sm[((i + 1) & 7) * 256 + threadIdx.x])) : "memory");
        REP8(X)
#undef X
      } else if (KIND == 10) { // 1 IMAD + 1 LOP3
#define X(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(b[i])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(m));
        REP8(X)
#undef X
      } else if (KIND == 11) { // 1 WIDE + 1 LOP3 + 1 IADD + 1 STS
#define X(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(m)); \
             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); \
             asm volatile("st.shared.u32 [%1], %0;" :: "r"(a[i]), "r"((uint32_t)__cvta_generic_to_shared(&sm[i * 256 + threadIdx.x])) : "memory");
        REP8(X)
#undef X
      } else if (KIND == 12) { // IADD3 with carry chain (add.cc / addc.cc)
#define X(i) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(m), "r"(seed));
        REP8(X)
#undef X
      } else if (KIND == 13) { // SEL
#define X(i) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.u32 %0, %0, %1, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(m));
        REP8(X)
#undef X
      }
    }
  }
  uint32_t r = 0;
  for (int i = 0; i < 8; ++i) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ sm[i * 256 + threadIdx.x];
  if (r == 0x12345678u) out[0] = r;
}

template <int KIND>
static void run(const char *name, int ninstr, uint32_t *out)
{
  const int iters = 2000, blocks = 148 * 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<KIND><<<blocks, 256>>>(10, out, 12345u);
  cudaEventRecord(e0);
  k<KIND><<<blocks, 256>>>(iters, out, 12345u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  // warp-instructions per SMSP: 8 warps/SMSP * iters * 4 reps * ninstr
  const double wi = 8.0 * iters * 4 * ninstr;
  const double clks = ms * 1e-3 * 1.92e9;   // assume ~1.92 GHz under this load (reported alongside)
  printf("%-44s %3d instr/rep : %6.2f clk per warp-instr per SMSP (at 1.92 GHz), %.3f ms  [%s]\n", name, ninstr, clks / wi, ms, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  uint32_t *out; cudaMalloc(&out, 64);
  run<0>("IMAD (mad.lo)", 8, out);
  run<1>("IMAD.WIDE (mad.wide)", 8, out);
  run<14>("IMAD.WIDE c=0 (mul.wide)", 8, out);
  run<15>("add.u64 (IADD3 + IADD3.X)", 8, out);
  run<16>("IMAD.WIDE + LOP3 independent", 16, out);
  run<17>("IMAD.WIDE + IADD3 independent", 16, out);
  run<18>("IMAD + IADD3 independent", 16, out);
  run<19>("LOP3 + IADD3 independent", 16, out);
  run<2>("IMAD.HI (mad.hi)", 8, out);
  run<3>("WIDE pairs with carry (lo.cc/hi.cc) x2", 8, out);
  run<4>("IADD3 (add)", 8, out);
  run<5>("LOP3", 8, out);
  run<6>("SHF", 8, out);
  run<12>("add.cc + addc", 16, out);
  run<13>("ISETP + SEL", 16, out);
  run<7>("1 WIDE + 2 LOP3", 24, out);
  run<8>("1 WIDE + 1 IMAD", 16, out);
  run<10>("1 IMAD + 1 LOP3", 16, out);
  run<9>("STS + LDS (32-bit)", 16, out);
  run<11>("1 WIDE + LOP3 + IADD + STS", 32, out);
  return 0;
}

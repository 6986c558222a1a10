/*
 * qb_tc.cuh — thin inline-PTX layer for the Blackwell (sm_100a) tensor path: mbarrier, TMA
 * (cp.async.bulk.tensor), tcgen05 alloc / mma kind::i8 / commit / ld, shared-memory matrix and
 * instruction descriptors.  Used by qb_ozaki.cu (tensor path) and qb_level2.cu (TMA tiles of the sliced qgemv).  No reference counterpart: the reference
 * (/root/reference) has no accelerator code at all.
 */
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* ---------------------------------------------------------------- mbarrier */
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* Wait for the phase with the given parity to complete.  A wait that lasts longer than ~2 s of SM
 * clocks can only be a protocol bug: trap instead of hanging the GPU. */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  const uint32_t a = smem_u32(bar);
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
    if (ok) break;
    if ((it & 0xfffu) == 0xfffu) {
      const long long t = clock64();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000LL) __trap();
    }
  }
}

/* ---------------------------------------------------------------- TMA */
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tm)
{
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}
/* 3-D tiled load global -> shared, completion on an mbarrier (complete_tx::bytes) */
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2)
{
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

/* 2-D tiled load global -> shared, completion on an mbarrier */
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
/* contiguous bulk copy global -> shared (bytes a multiple of 16, both addresses 16-byte aligned), completion on an mbarrier */
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"((uint64_t)src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

/* ---------------------------------------------------------------- tcgen05 */
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) /* whole warp */
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) /* whole warp */
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
/* all previously issued tcgen05.mma of this thread arrive on `bar` when they have completed */
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
/* D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, one CTA, issued by ONE thread */
__device__ __forceinline__ void mma_i8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
/* 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread t <-> TMEM lane base+t) */
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

/* Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with the
 * 128-byte swizzle (what a TMA box {128 B, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes): 8-row
 * groups are 1024 B apart (SBO), LBO is unused for swizzled K-major layouts, descriptor version 1
 * (sm_100), layout type 2 = SWIZZLE_128B.  The tile base must be 1024-byte aligned; stepping
 * along K inside the 128-byte row is done by adding the byte offset >> 4 to the start field. */
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

/* Instruction descriptor, kind::i8: D = S32 (c_format 2, bits [4,6)), A and B signed 8-bit
 * (format 1, bits [7,10) and [10,13)), both K-major (bits 15, 16 = 0), N >> 3 at [17,23),
 * M >> 4 at [24,29), no saturation (bit 3 = 0): the int32 accumulate wraps, which never happens
 * here because the slice planner bounds K * pairs * 2^14 below 2^31. */
__host__ __device__ constexpr uint32_t make_idesc_i8(uint32_t M, uint32_t N)
{
  return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

} // namespace tc
} // namespace qb

// quadblas/memory/allocation.hpp (B200 build) — aligned_alloc<T> / aligned_free
// (/root/reference/include/quadblas/memory/allocation.hpp:18-41; README:133-136).
// Buffers come from the CUDA library as page-locked host memory (32-byte aligned or better), so
// the staging copies of the host-pointer entry points run at full PCIe/C2C rate; nullptr on
// failure, like the reference.  aligned_free accepts anything aligned_alloc returned.
#ifndef QUADBLAS_MEMORY_ALLOCATION_HPP
#define QUADBLAS_MEMORY_ALLOCATION_HPP
#include "../core/constants.hpp"
#include "../core/platform.hpp"
#include <cstddef>
namespace QuadBLAS
{
  template <typename T>
  inline T *aligned_alloc(size_t count)
  {
    return static_cast<T *>(qb_host_alloc(count * sizeof(T)));
  }
  template <typename T>
  inline void aligned_free(T *ptr)
  {
    qb_host_free(const_cast<void *>(static_cast<const void *>(ptr)));
  }
} // namespace QuadBLAS
#endif // QUADBLAS_MEMORY_ALLOCATION_HPP

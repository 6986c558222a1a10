// quadblas/detail/blocking.hpp (B200 build).  The reference derives mc/kc/nc from host cache sizes
// (/root/reference/include/quadblas/detail/blocking.hpp:21-66); only kc reaches the results (the
// k-panel of the reference order).  The struct is kept for source compatibility;
// kc follows the reference's formula and is what qb_set_kc() defaults to (126 on x86-64 Linux).
#ifndef QUADBLAS_DETAIL_BLOCKING_HPP
#define QUADBLAS_DETAIL_BLOCKING_HPP
#include "../core/constants.hpp"
#include <algorithm>
namespace QuadBLAS
{
  struct BlockingParams
  {
    size_t mc, kc, nc;
    BlockingParams(size_t m, size_t n, size_t k)
    {
      const size_t quad = 32;                                  // the reference budgets 32 bytes per element
      const size_t kc_cache = (L1_CACHE_SIZE / quad - 16) / 8; // = 126
      kc = std::max<size_t>(4, std::min({kc_cache, size_t(256), k}));
      mc = std::max<size_t>(4, std::min<size_t>(64, m) / 4 * 4);
      nc = std::max<size_t>(4, std::min<size_t>(4, n));
    }
  };
} // namespace QuadBLAS
#endif // QUADBLAS_DETAIL_BLOCKING_HPP

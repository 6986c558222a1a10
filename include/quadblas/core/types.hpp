// quadblas/core/types.hpp (B200 build) — Layout and the container forward declarations
// (/root/reference/include/quadblas/core/types.hpp:8-19).
#ifndef QUADBLAS_CORE_TYPES_HPP
#define QUADBLAS_CORE_TYPES_HPP
namespace QuadBLAS
{
  enum class Layout { RowMajor, ColMajor };
  template <Layout layout> class Matrix;
  template <Layout layout> class Vector;
  namespace b200
  {
    inline char layout_char(Layout l) { return l == Layout::ColMajor ? 'C' : 'R'; }
  }
} // namespace QuadBLAS
#endif // QUADBLAS_CORE_TYPES_HPP

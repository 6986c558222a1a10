// quadblas/interface/c_interface.hpp (B200 build).  In the reference the C entry points are inline
// definitions in this header (/root/reference/include/quadblas/interface/c_interface.hpp:15-146);
// here they are exported by libqblas_b200.so under the same names and signatures and only declared
// (include/qblas_b200.h section A): quadblas_qdot, quadblas_qnrm2, quadblas_qaxpy, quadblas_qgemv,
// quadblas_qgemm, quadblas_set_num_threads, quadblas_get_num_threads, quadblas_get_version,
// quadblas_is_aligned.
#ifndef QUADBLAS_INTERFACE_C_INTERFACE_HPP
#define QUADBLAS_INTERFACE_C_INTERFACE_HPP
#include "../../qblas_b200.h"
#endif // QUADBLAS_INTERFACE_C_INTERFACE_HPP

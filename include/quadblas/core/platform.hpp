// quadblas/core/platform.hpp (B200 build) — what the reference's platform header provides to user
// code (/root/reference/include/quadblas/core/platform.hpp:5-24): architecture macros, the SLEEF
// quad type header and the SLEEF_QUAD_C fallback.  No OpenMP here: threading is the GPU's.
#ifndef QUADBLAS_CORE_PLATFORM_HPP
#define QUADBLAS_CORE_PLATFORM_HPP

#if defined(__x86_64__) || defined(_M_X64)
#define QUADBLAS_X86_64
#elif defined(__aarch64__) || defined(_M_ARM64)
#define QUADBLAS_AARCH64
#endif
#define QUADBLAS_B200 1 /* this build forwards every routine to libqblas_b200.so (CUDA, sm_100a) */

#if defined(QUADBLAS_USE_SLEEF) || (defined(__has_include) && !defined(QUADBLAS_NO_SLEEF))
#if defined(QUADBLAS_USE_SLEEF)
#include <sleefquad.h>
#elif __has_include(<sleefquad.h>)
#include <sleefquad.h>
#else
#include "../b200/sleefquad_compat.h"
#endif
#else
#include "../b200/sleefquad_compat.h"
#endif

#ifndef SLEEF_QUAD_C
#define SLEEF_QUAD_C(x) Sleef_cast_from_doubleq1(x)
#endif

#include "../../qblas_b200.h" /* the C ABI of the CUDA library */

#endif // QUADBLAS_CORE_PLATFORM_HPP

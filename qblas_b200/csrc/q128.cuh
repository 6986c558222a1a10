/* The integer-limb binary128 core lives with the public headers (the host-side SLEEF compatibility
 * layer include/quadblas/b200/sleefquad_compat.h compiles the same source for callers). */
#pragma once
#include "../../include/quadblas/b200/q128.cuh"

/* placeholder: chain-form primitives come next; for now alias to the generic routine */
#pragma once
#include "q128.cuh"
namespace qb {
QB_HD q128 q_fma_fast(q128 a, q128 b, q128 c) { return q_fma(a, b, c); }
}

"""TEST INFRASTRUCTURE: exact (big-integer) inner products of quads, rounded once to binary128.

The fast-mode tensor path (csrc/qb_ozaki.cu) computes each inner product of
/root/reference/include/quadblas/algorithms/level3.hpp:77-85 exactly and rounds once, so its
checker is exact rational arithmetic, not the reference's rounding order."""
from fractions import Fraction

import numpy as np

from qblas_b200 import quad

BIAS = 16383


def to_int_exp(q):
    """quads (..., 2) uint64 -> (python int mantissas with sign, exponents) lists; value = m * 2^e."""
    flat = q.reshape(-1, 2)
    ms, es = [], []
    for lo, hi in flat:
        hi = int(hi); lo = int(lo)
        ef = (hi >> 48) & 0x7FFF
        m = ((hi & ((1 << 48) - 1)) << 64) | lo
        if ef == 0x7FFF:
            raise ValueError("inf/nan")
        if ef:
            m |= 1 << 112
        else:
            ef = 1
        ms.append(-m if hi >> 63 else m)
        es.append(ef - BIAS - 112)
    return ms, es


def exact_matmul_rounded(A, lda, B, ldb, m, n, k, layout="R"):
    """round_to_binary128(sum_l A(i,l) B(l,j)) for all i, j -> (m*n, 2) uint64, row-major (i, j)."""
    am, ae = to_int_exp(A)
    bm, be = to_int_exp(B)
    a_at = (lambda i, l: i * lda + l) if layout == "R" else (lambda i, l: l * lda + i)
    b_at = (lambda l, j: l * ldb + j) if layout == "R" else (lambda l, j: j * ldb + l)
    emin = min(ae) + min(be)
    out = np.zeros((m * n, 2), dtype=np.uint64)
    for i in range(m):
        arow = [(am[a_at(i, l)], ae[a_at(i, l)]) for l in range(k)]
        for j in range(n):
            acc = 0
            for l in range(k):
                x, ex = arow[l]
                p = b_at(l, j)
                if x and bm[p]:
                    acc += (x * bm[p]) << (ex + be[p] - emin)
            hi, lo = quad.from_fraction(Fraction(acc) * Fraction(2) ** emin) if acc else (0, 0)
            out[i * n + j, 0] = lo
            out[i * n + j, 1] = hi
    return out

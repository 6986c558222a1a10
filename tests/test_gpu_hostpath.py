"""GPU tests of the pipelined all-host qgemm path (csrc/qb_abi.cu qb_gemm: shared operand first, then C cut into four
slabs whose uploads / compute / downloads overlap on three streams).  Each slab is an ordinary qgemm on device pointers,
so the host-buffer result must equal the device-buffer result of the SAME call bit for bit — in reference-order mode
(where every bit is defined by /root/reference/include/quadblas/algorithms/level3.hpp:215-336) and in fast mode."""
import numpy as np
import pytest
import torch

import qgen
from gpu_util import to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


def _run_both(qb, layout, ta, tb, m, n, k, lda, ldb, ldc, A, B, C0, alpha, beta):
    dC = to_dev(C0)
    qb.gemm(layout, m, n, k, alpha, to_dev(A), lda, to_dev(B), ldb, beta, dC, ldc, transa=ta, transb=tb)
    torch.cuda.synchronize()
    hC = C0.copy()
    qb.gemm(layout, m, n, k, alpha, A, lda, B, ldb, beta, hC, ldc, transa=ta, transb=tb)   # numpy arrays: host path
    return to_host(dC), hC


@pytest.mark.parametrize("layout,ta,tb", [("R", "N", "N"), ("C", "N", "N"), ("R", "T", "N"), ("C", "N", "T")])
def test_pipelined_host_path_equals_device_path_reference_order(qb, layout, ta, tb):
    rng = np.random.default_rng(ord(layout) + ord(ta) * 3 + ord(tb) * 5)
    m, n, k = (1536, 1400, 1000) if layout == "R" else (1400, 1536, 1000)   # the split dimension (m for R, n for C) is 1536 -> 4 slabs of 384
    col = layout == "C"
    tA, tB = ta == "T", tb == "T"
    # storage shapes (outer, inner) exactly as qb_gemm derives them
    a_shape = (k, m) if (col != tA) else (m, k)
    b_shape = (n, k) if (col != tB) else (k, n)
    c_shape = (n, m) if col else (m, n)
    lda, ldb, ldc = a_shape[1] + 3, b_shape[1] + 1, c_shape[1] + 2
    A = qgen.matrix(rng, a_shape[0], a_shape[1], "D113", lda); B = qgen.matrix(rng, b_shape[0], b_shape[1], "D113", ldb)
    C0 = qgen.matrix(rng, c_shape[0], c_shape[1], "D113", ldc)
    alpha, beta = quad.random_quads(rng, 2)
    assert A.nbytes + B.nbytes + C0.nbytes >= 64 << 20          # large enough for the pipelined branch
    qb.set_mode(qb.MODE_REFERENCE); qb.set_honor_trans(tA or tB)
    try:
        dev, host = _run_both(qb, layout, ta, tb, m, n, k, lda, ldb, ldc, A, B, C0, alpha, beta)
    finally:
        qb.set_honor_trans(False)
    assert quad.same_bits(dev, host).all(), f"{(~quad.same_bits(dev, host)).sum()} entries differ"
    assert not quad.same_bits(host, C0).all()


def test_pipelined_host_path_fast_mode(qb):
    rng = np.random.default_rng(9)
    m, n, k = 2048, 1024, 1024
    A = qgen.matrix(rng, m, k, "D113", k); B = qgen.matrix(rng, k, n, "D113", n); C0 = qgen.matrix(rng, m, n, "D113", n)
    alpha, beta = quad.random_quads(rng, 2)
    qb.set_mode(qb.MODE_FAST)
    try:
        dev, host = _run_both(qb, "R", "N", "N", m, n, k, k, n, n, A, B, C0, alpha, beta)
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    assert st["pairs"] > 0                                        # the slabs went through the tensor path
    assert quad.same_bits(dev, host).all(), f"{(~quad.same_bits(dev, host)).sum()} entries differ"

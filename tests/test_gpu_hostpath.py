"""GPU tests of the pipelined all-host paths (csrc/qb_abi.cu).  qb_gemm: the shared operand is uploaded first, then the rows
stream in, are multiplied and stream out on three streams — in reference-order mode as slabs that are ordinary qgemms on device
pointers, in fast mode as ONE tensor-path call whose row passes wait for their slabs (passes outer, streamed rows).  qb_gemv: A is
uploaded in row slabs while the earlier slabs are multiplied.  The host-buffer result must equal the device-buffer result of the
SAME call bit for bit — in reference-order mode (where every bit is defined by
/root/reference/include/quadblas/algorithms/level3.hpp:215-336, level2.hpp:15-82) and in fast mode."""
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


@pytest.mark.parametrize("layout,m,n,k", [("R", 2300, 1100, 900), ("C", 1100, 2300, 900), ("R", 4200, 520, 640)])
def test_streamed_host_path_fast_mode_layouts_and_ragged_passes(qb, layout, m, n, k):
    """The fast-mode all-host call is one tensor-path qgemm with streamed rows: ragged slabs / passes, padded leading dimensions, and
    col-major through the exchanged roles (C^T = B^T A^T: the columns of C are the streamed 'rows')."""
    rng = np.random.default_rng(m + n)
    col = layout == "C"
    a_shape = (k, m) if col else (m, k); b_shape = (n, k) if col else (k, n); c_shape = (n, m) if col else (m, n)
    lda, ldb, ldc = a_shape[1] + 3, b_shape[1] + 1, c_shape[1] + 2
    A = qgen.matrix(rng, a_shape[0], a_shape[1], "D113", lda); B = qgen.matrix(rng, b_shape[0], b_shape[1], "D113", ldb)
    C0 = qgen.matrix(rng, c_shape[0], c_shape[1], "D113", ldc)
    alpha, beta = quad.random_quads(rng, 2)
    assert A.nbytes + B.nbytes + C0.nbytes >= 64 << 20
    qb.set_mode(qb.MODE_FAST)
    try:
        dev, host = _run_both(qb, layout, "N", "N", m, n, k, lda, ldb, ldc, A, B, C0, alpha, beta)
        st = qb.oz_last_stats()
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    assert st["pairs"] > 0 and st["row_passes"] >= 3, st
    assert quad.same_bits(dev, host).all(), f"{(~quad.same_bits(dev, host)).sum()} entries differ"
    idx = np.array([[(i * ldc + j) if not col else (j * ldc + i) for j in range(n)] for i in range(m)]).reshape(-1)
    mask = np.ones(C0.shape[0], dtype=bool); mask[idx] = False
    assert (host[mask] == C0[mask]).all()                           # padding between the rows is not touched


@pytest.mark.parametrize("layout,mode", [("R", "ref"), ("C", "ref"), ("R", "fast"), ("C", "fast")])
def test_pipelined_host_gemv_equals_device_path(qb, layout, mode):
    """qb_gemv with a large host A: uploaded in row slabs (contiguous for row-major, 2-D copies of a row range for col-major), each
    slab an ordinary qgemv on its rows: same bits as the device-resident call."""
    rng = np.random.default_rng(ord(layout) + len(mode))
    m, n = 2100, 2060
    lda = (n if layout == "R" else m) + 5
    A = qgen.matrix(rng, m if layout == "R" else n, n if layout == "R" else m, "D113", lda)
    x = quad.random_quads(rng, n * 2); y0 = quad.random_quads(rng, m * 3)
    alpha, beta = quad.random_quads(rng, 2)
    assert A.nbytes >= 64 << 20
    qb.set_mode(qb.MODE_FAST if mode == "fast" else qb.MODE_REFERENCE)
    try:
        dy = to_dev(y0)
        qb.gemv(layout, m, n, alpha, to_dev(A), lda, to_dev(x), 2, beta, dy, 3)
        torch.cuda.synchronize()
        hy = y0.copy()
        qb.gemv(layout, m, n, alpha, A, lda, x, 2, beta, hy, 3)
    finally:
        qb.set_mode(qb.MODE_REFERENCE)
    assert quad.same_bits(to_host(dy), hy).all()
    assert not quad.same_bits(hy, y0).all()


@pytest.mark.parametrize("layout,mode", [("R", "fast"), ("C", "fast"), ("R", "ref")])
@pytest.mark.parametrize("beta_bits", [0, 1 << 63])
def test_pipelined_host_path_beta_zero_sends_classes_of_c(qb, layout, mode, beta_bits):
    """beta = +-0 with contiguous rows of C: the host path classifies C_in (finite sign / Inf-NaN) instead of uploading it and the
    device multiplies beta by a stand-in of the same class — the result must carry the bits of the device-resident call, which
    evaluates beta * C_in itself (level3.hpp:107: 0 * NaN and 0 * Inf are NaN, 0 * negative is -0), and of the same call with the
    class path switched off."""
    rng = np.random.default_rng(31 + (beta_bits >> 63))
    m, n, k = (2048, 1024, 1024) if layout == "R" else (1024, 2048, 1024)
    col = layout == "C"
    a_shape = (k, m) if col else (m, k); b_shape = (n, k) if col else (k, n); c_shape = (n, m) if col else (m, n)
    A = qgen.matrix(rng, a_shape[0], a_shape[1], "D113"); B = qgen.matrix(rng, b_shape[0], b_shape[1], "D113")
    C0 = qgen.matrix(rng, c_shape[0], c_shape[1], "D113")
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); nan = np.array([1, 0x7FFF << 48], dtype=np.uint64)
    C0[5] = inf; C0[6] = inf ^ np.array([0, 1 << 63], dtype=np.uint64); C0[7] = nan; C0[8] = 0; C0[9] = np.array([0, 1 << 63], dtype=np.uint64)
    C0[10] = np.array([3, 0], dtype=np.uint64); C0[c_shape[1] * 700 + 3] = nan; C0[-1] = inf
    # a zero row of A: alpha * S = 0 there, so the sign of 0 * C decides the sign of the result
    if not col:
        A[3 * k: 4 * k] = 0
    alpha = quad.random_quads(rng, 1)[0]
    beta = np.array([0, beta_bits], dtype=np.uint64)
    assert A.nbytes + B.nbytes + C0.nbytes >= 64 << 20
    qb.set_mode(qb.MODE_FAST if mode == "fast" else qb.MODE_REFERENCE)
    try:
        dev, host = _run_both(qb, layout, "N", "N", m, n, k, a_shape[1], b_shape[1], c_shape[1], A, B, C0, alpha, beta)
        qb.set_beta0_classes(0)
        hC = C0.copy()
        qb.gemm(layout, m, n, k, alpha, A, a_shape[1], B, b_shape[1], beta, hC, c_shape[1])
    finally:
        qb.set_beta0_classes(1)
        qb.set_mode(qb.MODE_REFERENCE)
    assert quad.same_bits(dev, host).all(), f"{(~quad.same_bits(dev, host)).sum()} entries differ"
    assert quad.same_bits(hC, host).all()
    assert quad.is_nan(host[[5, 6, 7]]).all() and not quad.is_nan(host[[8, 9, 10]]).any()

"""Pins oracle/qoracle.c (the CPU restatement) bitwise against the reference's own loops compiled
from /root/reference (oracle/_ref/libqref.so) with full-113-bit-mantissa inputs — the only inputs
for which reduction order matters (SURVEY exec-summary)."""
import numpy as np
import pytest

import qgen
from qblas_b200 import quad

GEMM_SHAPES = [(70, 9, 300), (65, 66, 127), (130, 5, 253), (33, 70, 64), (64, 64, 64), (3, 200, 126),
               (47, 31, 23), (1, 1, 1), (2, 2, 2), (16, 16, 16), (66, 4, 1), (5, 65, 2), (4, 4, 125), (4, 4, 252)]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("kind", ["D113", "Dexp"])
def test_gemm_rowmajor(oracle, ref, m, n, k, kind):
    rng = np.random.default_rng(m * 10007 + n * 101 + k)
    lda, ldb, ldc = k + 3, n + 1, n + 2
    A = qgen.matrix(rng, m, k, kind, lda); B = qgen.matrix(rng, k, n, kind, ldb); C0 = qgen.matrix(rng, m, n, kind, ldc)
    alpha = quad.random_quads(rng, 1)[0]; beta = quad.random_quads(rng, 1)[0]
    Cr, Co = C0.copy(), C0.copy()
    ref.gemm("R", m, n, k, alpha, A, lda, B, ldb, beta, Cr, ldc)
    oracle.gemm("R", m, n, k, alpha, A, lda, B, ldb, beta, Co, ldc)
    assert quad.same_bits(Cr, Co).all()


def test_gemm_structured_inputs(oracle, ref):
    """zeros of both signs (triangular rows, zero rows / columns), exact cancellation inside a panel, Inf / NaN / a subnormal, three
    panels of k, beta = random / +0 / -0: the oracle must give the reference's bits here too — these are the inputs on which the
    product's branch-free qgemm kernel takes its special paths (tests/test_gpu_gemm.py runs both kernels on the same case)."""
    rng = np.random.default_rng(2026)
    m, n, k = 70, 45, 300
    A, B, C0 = qgen.structured_gemm_case(rng, m, n, k)
    alpha, beta = quad.random_quads(rng, 2)
    for bt in (beta, quad.from_double(np.array([0.0]))[0], quad.from_double(np.array([-0.0]))[0]):
        Cr, Co = C0.copy(), C0.copy()
        ref.gemm("R", m, n, k, alpha, A, k, B, n, bt, Cr, n)
        oracle.gemm("R", m, n, k, alpha, A, k, B, n, bt, Co, n)
        nan_r, nan_o = quad.is_nan(Cr), quad.is_nan(Co)
        assert (nan_r == nan_o).all() and nan_r.any()
        assert quad.same_bits(Cr[~nan_r], Co[~nan_o]).all()          # a NaN's payload is the arithmetic library's business
        zero = (Cr[:, 0] == 0) & ((Cr[:, 1] & np.uint64(0x7FFFFFFFFFFFFFFF)) == 0)
        assert zero.any() or bt is beta                                # the zero rows give signed zeros when beta is a zero


@pytest.mark.parametrize("m,n,k", [(40, 33, 64), (7, 5, 3), (64, 64, 64)])
def test_gemm_colmajor_simple_path(oracle, ref, m, n, k):
    """ColMajor is only correct in the reference when all dims <= 64 (gemm_simple); SURVEY bug 1."""
    rng = np.random.default_rng(m + n + k)
    lda, ldb, ldc = m + 1, k + 2, m + 3
    A = qgen.matrix(rng, k, m, "D113", lda); B = qgen.matrix(rng, n, k, "D113", ldb); C0 = qgen.matrix(rng, n, m, "D113", ldc)
    alpha = quad.random_quads(rng, 1)[0]; beta = quad.random_quads(rng, 1)[0]
    Cr, Co = C0.copy(), C0.copy()
    ref.gemm("C", m, n, k, alpha, A, lda, B, ldb, beta, Cr, ldc)
    oracle.gemm("C", m, n, k, alpha, A, lda, B, ldb, beta, Co, ldc)
    assert quad.same_bits(Cr, Co).all()


def test_c_qgemm_ignores_trans_and_widens_scalars(oracle, ref):
    rng = np.random.default_rng(5)
    m, n, k = 20, 17, 130
    A = qgen.matrix(rng, m, k); B = qgen.matrix(rng, k, n); C0 = qgen.matrix(rng, m, n)
    Cr, Co = C0.copy(), C0.copy()
    ref.c_qgemm("R", "T", "T", m, n, k, 2.5, A, k, B, n, 1.5, Cr, n)
    oracle.c_qgemm("R", "T", "T", m, n, k, 2.5, A, k, B, n, 1.5, Co, n)
    assert quad.same_bits(Cr, Co).all()


@pytest.mark.parametrize("layout,trans", [("R", "N"), ("R", "T"), ("C", "N"), ("C", "T")])
@pytest.mark.parametrize("m,n", [(77, 131), (3, 3), (1, 7), (600, 5), (20, 501)])
@pytest.mark.parametrize("incx,incy", [(1, 1), (2, 3)])
def test_c_qgemv(oracle, ref, layout, trans, m, n, incx, incy):
    rng = np.random.default_rng(m * 31 + n)
    # storage: rows x cols with ld; for 'R' the stored matrix is m x n, for 'C' it is n x m walk
    rows, cols = (m, n) if layout == "R" else (n, m)
    lda = cols + 2
    A = qgen.matrix(rng, rows, cols, "D113", lda)
    xn, yn = (n, m) if trans == "N" else (m, n)
    x = quad.random_quads(rng, (xn - 1) * incx + 1); y0 = quad.random_quads(rng, (yn - 1) * incy + 1)
    yr, yo = y0.copy(), y0.copy()
    ref.c_qgemv(layout, trans, m, n, 1.25, A, lda, x, incx, -0.75, yr, incy)
    oracle.c_qgemv(layout, trans, m, n, 1.25, A, lda, x, incx, -0.75, yo, incy)
    assert quad.same_bits(yr, yo).all()


@pytest.mark.parametrize("T", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("n,incx,incy", [(100003, 1, 1), (499, 1, 1), (500, 1, 1), (1, 1, 1), (7, 1, 1), (1000, 3, 2), (300, 2, 1), (5, 1, 1)])
def test_dot_depends_on_T(oracle, ref, T, n, incx, incy):
    rng = np.random.default_rng(n + T)
    x = quad.random_quads(rng, (n - 1) * incx + 1); y = quad.random_quads(rng, (n - 1) * incy + 1)
    ref.set_num_threads(T)
    assert quad.same_bits(ref.dot(n, x, incx, y, incy), oracle.dot(n, x, incx, y, incy, T)).all()
    assert quad.same_bits(ref.nrm2(n, x, incx), oracle.nrm2(n, x, incx, T)).all()


def test_axpy(oracle, ref):
    rng = np.random.default_rng(9)
    for n, incx, incy in [(5000, 1, 1), (700, 2, 3), (10, 1, 1)]:
        x = quad.random_quads(rng, (n - 1) * incx + 1); y0 = quad.random_quads(rng, (n - 1) * incy + 1)
        alpha = quad.random_quads(rng, 1)[0]
        yr, yo = y0.copy(), y0.copy()
        ref.axpy(n, alpha, x, incx, yr, incy); oracle.axpy(n, alpha, x, incx, yo, incy)
        assert quad.same_bits(yr, yo).all()

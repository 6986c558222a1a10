"""GPU parity of qgemv / qdot / qnrm2 / qaxpy against the oracle (reference order, bit exact),
through the C ABI.  Shapes follow test_quadblas.cpp:205-431 and SURVEY §8d cfg5."""
import numpy as np
import pytest
import torch

import qgen
from gpu_util import dev_random, to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout,trans", [("R", "N"), ("R", "T"), ("C", "N"), ("C", "T")])
@pytest.mark.parametrize("m,n", [(3, 3), (100, 100), (77, 131), (1, 7), (7, 1), (600, 5), (20, 501), (2000, 333), (129, 64)])
@pytest.mark.parametrize("incx,incy", [(1, 1), (2, 3)])
def test_c_qgemv_bitexact(qb, oracle, layout, trans, m, n, incx, incy):
    rng = np.random.default_rng(m * 31 + n + incx)
    rows, cols = (m, n) if layout == "R" else (n, m)
    lda = cols + (m % 3)
    kind = ["D113", "Dexp", "D53"][(m + n) % 3]
    A = qgen.matrix(rng, rows, cols, kind, lda)
    xn, yn = (n, m) if trans == "N" else (m, n)
    x = quad.random_quads(rng, (xn - 1) * incx + 1, kind); y0 = quad.random_quads(rng, (yn - 1) * incy + 1, kind)
    yg, yo = y0.copy(), y0.copy()
    qb.quadblas_qgemv(layout, trans, m, n, 1.5, A, lda, x, incx, 0.5, yg, incy)
    oracle.c_qgemv(layout, trans, m, n, 1.5, A, lda, x, incx, 0.5, yo, incy)
    assert quad.same_bits(yg, yo).all()


def test_gemv_quad_scalars_device_path_and_specials(qb, oracle):
    rng = np.random.default_rng(3)
    m, n = 300, 257
    A = qgen.matrix(rng, m, n); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    sa, sb, sc = qgen.triples(rng, 50, "specials")
    A[rng.integers(0, m * n, 50)] = sa; x[rng.integers(0, n, 10)] = sb[:10]; y0[rng.integers(0, m, 10)] = sc[:10]
    alpha, beta = quad.random_quads(rng, 2)
    for layout in "RC":
        mm, nn = (m, n) if layout == "R" else (n, m)
        xx = x[:nn] if nn <= n else np.concatenate([x, quad.random_quads(rng, nn - n)])
        yy = y0[:mm] if mm <= m else np.concatenate([y0, quad.random_quads(rng, mm - m)])
        yo = yy.copy(); oracle.gemv(layout, mm, nn, alpha, A, n if layout == "R" else n, xx, 1, beta, yo, 1)
        dy = to_dev(yy)
        qb.gemv(layout, mm, nn, alpha, to_dev(A), n, to_dev(xx), 1, beta, dy, 1)
        assert quad.same_bits(to_host(dy), yo).all(), layout
    for beta0 in (0.0,):  # beta = 0 with NaN/Inf in y propagates (level2.hpp:48)
        yg, yo = y0.copy(), y0.copy()
        qb.gemv("R", m, n, 1.0, A, n, x, 1, beta0, yg, 1); oracle.gemv("R", m, n, 1.0, A, n, x, 1, beta0, yo, 1)
        assert quad.same_bits(yg, yo).all()


def test_gemv_empty_is_noop(qb):
    y0 = quad.random_quads(np.random.default_rng(1), 5)
    yg = y0.copy()
    qb.quadblas_qgemv("R", "N", 0, 5, 1.0, y0, 5, y0, 1, 0.0, yg, 1)
    qb.quadblas_qgemv("C", "N", 5, 0, 1.0, y0, 5, y0, 1, 0.0, yg, 1)
    assert quad.same_bits(yg, y0).all()  # level2.hpp:21,59: y NOT scaled


@pytest.mark.parametrize("T", [1, 2, 3, 8, 64])
@pytest.mark.parametrize("n,incx,incy", [(10, 1, 1), (100003, 1, 1), (499, 1, 1), (500, 1, 1), (501, 1, 1), (1, 1, 1),
                                         (1000, 3, 2), (300, 2, 1), (63, 1, 1), (0, 1, 1)])
def test_dot_nrm2_reference_order(qb, oracle, T, n, incx, incy):
    rng = np.random.default_rng(n + T)
    x = quad.random_quads(rng, max((n - 1) * incx + 1, 1)); y = quad.random_quads(rng, max((n - 1) * incy + 1, 1))
    qb.set_mode(qb.MODE_REFERENCE)
    qb.quadblas_set_num_threads(T)
    try:
        assert quad.same_bits(qb.dot(n, x, incx, y, incy), oracle.dot(n, x, incx, y, incy, T)).all()
        assert quad.same_bits(qb.nrm2(n, x, incx), oracle.nrm2(n, x, incx, T)).all()
        # reference C ABI returns double (c_interface.hpp:30,43)
        assert qb.quadblas_qdot(n, x, incx, y, incy) == oracle.to_double(oracle.dot(n, x, incx, y, incy, T))
        if n:
            assert qb.quadblas_qnrm2(n, x, incx) == oracle.to_double(oracle.nrm2(n, x, incx, T))
    finally:
        qb.quadblas_set_num_threads(0)


def test_dot_fast_mode_bound_determinism_and_cancellation(qb, oracle):
    from fractions import Fraction
    rng = np.random.default_rng(21)
    n = 1_000_003
    x = quad.random_quads(rng, n, "Dexp"); y = quad.random_quads(rng, n, "D113")
    qb.set_mode(qb.MODE_FAST)
    try:
        r1 = qb.dot(n, x, 1, y, 1); r2 = qb.dot(n, x, 1, y, 1)
        assert quad.same_bits(r1, r2).all()  # deterministic tree
        exact = oracle.dot(n, x, 1, y, 1, 1)
        absx = x.copy(); absx[:, 1] &= np.uint64((1 << 63) - 1); absy = y.copy(); absy[:, 1] &= np.uint64((1 << 63) - 1)
        ab = oracle.dot(n, absx, 1, absy, 1, 1)
        f = lambda v: quad.to_fraction(int(v[1]), int(v[0]))
        u = Fraction(1, 2 ** 113); gam = n * u / (1 - n * u)
        assert abs(f(r1) - f(exact)) <= 2 * gam * f(ab)
        xv = np.zeros(10); xv[:3] = [1e20, 1.0, -1e20]
        assert f(qb.dot(10, quad.from_double(xv), 1, quad.from_double(np.ones(10)), 1)) == 1
    finally:
        qb.set_mode(qb.MODE_REFERENCE)


def test_axpy(qb, oracle):
    rng = np.random.default_rng(9)
    for n, incx, incy in [(5000, 1, 1), (700, 2, 3), (10, 1, 1), (1, 1, 1)]:
        x = quad.random_quads(rng, (n - 1) * incx + 1, "Dexp"); y0 = quad.random_quads(rng, (n - 1) * incy + 1, "Dexp")
        alpha = quad.random_quads(rng, 1)[0]
        yg, yo = y0.copy(), y0.copy()
        qb.axpy(n, alpha, x, incx, yg, incy); oracle.axpy(n, alpha, x, incx, yo, incy)
        assert quad.same_bits(yg, yo).all()
    yg = y0.copy(); qb.quadblas_qaxpy(0, 2.0, x, 1, yg, 1); assert quad.same_bits(yg, y0).all()
    yg, yo = y0.copy(), y0.copy()
    qb.quadblas_qaxpy(n, 2.5, x, incx, yg, incy); oracle.axpy(n, 2.5, x, incx, yo, incy)
    assert quad.same_bits(yg, yo).all()


def test_gemv_full_size_property(qb, oracle):
    """BASELINE config 2 scale: 8192 x 8192 row-major on device; 64 sampled rows recomputed on the CPU
    in reference order (bit exact) + linearity-free idempotence: the same call twice gives the same bits."""
    n = 8192
    A = dev_random((n * n,), "D113", 5); x = dev_random((n,), "D113", 6); y = dev_random((n,), "D113", 7)
    y2 = y.clone(); yin = to_host(y)
    qb.gemv("R", n, n, 1.0, A, n, x, 1, 0.0, y, 1)
    qb.gemv("R", n, n, 1.0, A, n, x, 1, 0.0, y2, 1)
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    rows = np.random.default_rng(2).integers(0, n, 64)
    xh = to_host(x); got = to_host(y)
    for r in rows:
        Ar = to_host(A[r * n:(r + 1) * n])
        yo = yin[r:r + 1].copy()
        oracle.gemv("R", 1, n, 1.0, Ar, n, xh, 1, 0.0, yo, 1)
        assert quad.same_bits(got[r], yo[0]).all()

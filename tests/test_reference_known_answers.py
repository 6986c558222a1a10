"""Every known-answer value the reference's own tests hold for the hot path, checked
on the oracle (CPU).  The GPU twins live in tests/test_gpu_known_answers.py."""
from fractions import Fraction

import numpy as np

from qblas_b200 import quad


def qd(vals):
    return quad.from_double(np.asarray(vals, dtype=np.float64))


def as_int(q):
    return quad.to_fraction(int(q[1]), int(q[0]))


def test_dot_1_to_10(oracle):  # test_quadblas.cpp:205-224
    x = qd(np.arange(1, 11)); y = qd(np.ones(10))
    assert as_int(oracle.dot(10, x, 1, y, 1, 8)) == 55


def test_dot_1_to_5_and_norm(oracle):  # tests/debug_test.cpp:50-85, tests/test_sleef_simd.cpp:64-85
    x = qd(np.arange(1, 6)); y = qd(np.ones(5))
    assert as_int(oracle.dot(5, x, 1, y, 1, 8)) == 15
    r = oracle.nrm2(5, x, 1, 8)
    assert (int(r[1]), int(r[0])) == tuple(int(v) for v in oracle.sqrt(qd([55.0]))[0][::-1])


def test_norm_closed_form(oracle):  # test_quadblas.cpp:290-312
    n = 1000
    x = qd(np.arange(1, n + 1))
    r = oracle.nrm2(n, x, 1, 8)
    expect = oracle.sqrt(qd([n * (n + 1) * (2 * n + 1) / 6.0]))[0]
    assert quad.same_bits(r, expect)


def test_gemv_3x3_and_2x2(oracle):  # test_quadblas.cpp:327-350, tests/debug_test.cpp:87-110
    A = qd(np.arange(1, 10)); x = qd(np.ones(3)); y = qd(np.zeros(3))
    oracle.gemv("R", 3, 3, 1.0, A, 3, x, 1, 0.0, y, 1)
    assert [as_int(v) for v in y] == [6, 15, 24]
    A = qd([1, 2, 3, 4]); x = qd([1, 1]); y = qd([0, 0])
    oracle.gemv("R", 2, 2, 1.0, A, 2, x, 1, 0.0, y, 1)
    assert [as_int(v) for v in y] == [3, 7]


def test_gemm_2x2(oracle):  # test_quadblas.cpp:446-464
    A = qd([1, 2, 3, 4]); B = qd([5, 6, 7, 8]); C = qd(np.zeros(4))
    oracle.gemm("R", 2, 2, 2, 1.0, A, 2, B, 2, 0.0, C, 2)
    assert [as_int(v) for v in C] == [19, 22, 43, 50]


def test_identity_50(oracle):  # test_quadblas.cpp:691-712
    rng = np.random.default_rng(42)
    n = 50
    A = quad.random_quads(rng, n * n); I = qd(np.eye(n).ravel()); C = qd(np.zeros(n * n))
    oracle.gemm("R", n, n, n, 1.0, I, n, A, n, 0.0, C, n)
    assert quad.same_bits(C, A).all()


def test_cancellation_1e20(oracle):  # test_quadblas.cpp:715-739, README:141-158, benchmark.cpp:249-270
    n = 10
    xv = np.zeros(n); xv[0], xv[1], xv[2] = 1e20, 1.0, -1e20
    x = qd(xv); y = qd(np.ones(n))
    for T in (1, 2, 8):
        assert as_int(oracle.dot(n, x, 1, y, 1, T)) == 1
    # as a gemv row and a gemm row as well
    yv = qd([0.0]); oracle.gemv("R", 1, n, 1.0, x, n, y, 1, 0.0, yv, 1)
    assert as_int(yv[0]) == 1
    C = qd([0.0]); oracle.gemm("R", 1, 1, n, 1.0, x, n, y, 1, 0.0, C, 1)
    assert as_int(C[0]) == 1


def test_quadvector_lane_values(oracle):  # tests/debug_test.cpp:31-48, tests/test_sleef_simd.cpp:39-42
    a = qd([2, 3]); b = qd([4, 5]); c = qd([6, 8])
    s = oracle.add(a, b)
    assert as_int(oracle.add(s[0:1], s[1:2])[0]) == 14          # (2,3)+(4,5), hsum
    f = oracle.fma(a, b, c)
    assert [as_int(v) for v in f] == [14, 23]                   # (2,3)*(4,5)+(6,8)
    assert as_int(oracle.add(qd([2.0]), qd([3.0]))[0]) == 5     # debug_test.cpp:20-28
    assert as_int(oracle.mul(qd([2.0]), qd([3.0]))[0]) == 6
    assert as_int(oracle.fma(qd([2.0]), qd([3.0]), qd([1.0]))[0]) == 7


def test_one_third_bits():  # SURVEY §8 [probe]
    assert quad.from_fraction(Fraction(1, 3)) == (0x3FFD555555555555, 0x5555555555555555)

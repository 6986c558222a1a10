"""GPU checks of the FAST-mode qdot / qnrm2 / qgemv (window accumulator, csrc/qwide.cuh) through the C ABI.

Fast mode re-associates, so the checker is exact rational arithmetic (not the reference's rounding
order): each inner sum must be the exact value rounded ONCE (bitwise, up to the stated 2^-133 window
truncation), then the reference epilogue y = fma(alpha, S, mul(beta, y)) (level2.hpp:48) through the
oracle's scalar ops.  That is far inside the fast-mode contract |r^ - r| <= gamma_n sum|a||b|, which is
asserted as well."""
from fractions import Fraction

import numpy as np
import pytest
import torch

import qgen
from gpu_util import dev_random, to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


def _frac(q):
    return quad.to_fraction(int(q[1]), int(q[0]))


def _fr_vec(v):
    return [_frac(q) for q in v.reshape(-1, 2)]


def _round(fr):
    hi, lo = quad.from_fraction(fr) if fr != 0 else (0, 0)
    return np.array([lo, hi], dtype=np.uint64)


@pytest.fixture()
def fast(qb):
    qb.set_mode(qb.MODE_FAST)
    qb.set_fast_variant(1)                          # the window accumulator everywhere (tests/test_gpu_gemv_sliced.py covers variant 2)
    yield qb
    qb.set_fast_variant(2)
    qb.set_mode(qb.MODE_REFERENCE)


@pytest.mark.parametrize("n,incx,incy,kind", [(1, 1, 1, "D113"), (31, 1, 1, "D113"), (1000, 1, 1, "Dexp"), (4099, 1, 1, "D113"),
                                              (70001, 1, 1, "D53"), (3000, 3, 2, "Dexp"), (160000, 1, 1, "D113")])
def test_fast_dot_is_exact_sum_rounded_once(fast, n, incx, incy, kind):
    rng = np.random.default_rng(n)
    x = quad.random_quads(rng, (n - 1) * incx + 1, kind); y = quad.random_quads(rng, (n - 1) * incy + 1, "D113")
    fx = _fr_vec(x[::incx][:n]); fy = _fr_vec(y[::incy][:n])
    tot = sum(a * b for a, b in zip(fx, fy)); sab = sum(abs(a * b) for a, b in zip(fx, fy))
    r = fast.dot(n, x, incx, y, incy)
    assert quad.same_bits(r, fast.dot(n, x, incx, y, incy)).all()     # deterministic tree
    err = abs(_frac(r) - tot)
    u = Fraction(1, 2 ** 113)
    assert err <= n * u / (1 - n * u) * sab
    assert err <= abs(tot) * u + n * sab / 2 ** 133
    # nrm2 = sqrt of the once-rounded exact sum of squares (one load per element in the kernel)
    s2 = sum(a * a for a in fx)
    want = _round(s2)
    got = fast.nrm2(n, x, incx)
    # sqrt is correctly rounded by the library core (checked against the oracle elsewhere): compare squares
    assert abs(_frac(got) ** 2 - _frac(want)) <= 4 * u * _frac(want)


def test_fast_dot_specials_and_cancellation(fast):
    rng = np.random.default_rng(2)
    n = 5000
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); ninf = np.array([0, 0xFFFF << 48], dtype=np.uint64)
    nan = np.array([1, 0x7FFF << 48], dtype=np.uint64); one = quad.from_double(np.array([1.0]))[0]
    xs, ys = x.copy(), y.copy(); xs[777] = inf; ys[777] = one
    assert quad.same_bits(fast.dot(n, xs, 1, ys, 1), inf).all()
    xs[4000] = ninf; ys[4000] = one
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    xs, ys = x.copy(), y.copy(); xs[12] = nan
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    xs, ys = x.copy(), y.copy(); xs[12] = inf; ys[12] = 0
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    # zeros, signed zeros and subnormals are exact in the window
    xs, ys = x[:300].copy(), y[:300].copy()
    xs[::7] = 0; ys[3::11] = 0; xs[5] = (np.uint64(99), np.uint64(0x0000_0000_0000_0001))
    ys[5, 1] = (ys[5, 1] & np.uint64(0x8000_FFFF_FFFF_FFFF)) | (np.uint64(0x7E00) << np.uint64(48))
    tot = sum(a * b for a, b in zip(_fr_vec(xs), _fr_vec(ys)))
    assert quad.same_bits(fast.dot(300, xs, 1, ys, 1), _round(tot)).all()
    xv = np.zeros(10); xv[:3] = [1e20, 1.0, -1e20]
    assert _frac(fast.dot(10, quad.from_double(xv), 1, quad.from_double(np.ones(10)), 1)) == 1
    assert _frac(fast.dot(0, x, 1, y, 1)) == 0


@pytest.mark.parametrize("layout,trans", [("R", "N"), ("R", "T"), ("C", "N"), ("C", "T")])
@pytest.mark.parametrize("m,n", [(3, 3), (77, 131), (1, 700), (300, 5), (129, 1025), (700, 64)])
@pytest.mark.parametrize("incx,incy", [(1, 1), (2, 3)])
def test_fast_gemv_exact_row_sums(fast, oracle, layout, trans, m, n, incx, incy):
    rng = np.random.default_rng(m * 31 + n + incx)
    rows, cols = (m, n) if layout == "R" else (n, m)
    lda = cols + (m % 3)
    kind = ["D113", "Dexp", "D53"][(m + n) % 3]
    A = qgen.matrix(rng, rows, cols, kind, lda)
    xn, yn = (n, m) if trans == "N" else (m, n)
    x = quad.random_quads(rng, (xn - 1) * incx + 1, kind); y0 = quad.random_quads(rng, (yn - 1) * incy + 1, kind)
    yg = y0.copy()
    fast.quadblas_qgemv(layout, trans, m, n, 1.5, A, lda, x, incx, 0.5, yg, incy)
    # exact: op(A)(i, j) with the C ABI's relabelling (c_interface.hpp:64-90)
    fx = _fr_vec(x[::incx][:xn])
    FA = {}
    def a_at(i, j):  # element (i, j) of the m x n matrix as laid out by `layout`
        p = i * lda + j if layout == "R" else j * lda + i
        if p not in FA:
            FA[p] = _frac(A[p])
        return FA[p]
    al, be = oracle.from_double(1.5), oracle.from_double(0.5)
    for i in range(yn):
        if trans == "N":
            S = sum(a_at(i, j) * fx[j] for j in range(n))
        else:
            S = sum(a_at(j, i) * fx[j] for j in range(m))
        want = oracle.fma(al, _round(S), oracle.mul(be, y0[i * incy]))
        assert quad.same_bits(yg[i * incy], want).all(), (i, layout, trans)
    # untouched gaps of a strided y
    if incy > 1:
        mask = np.ones(len(y0), dtype=bool); mask[::incy] = False
        assert quad.same_bits(yg[mask], y0[mask]).all()


def test_fast_gemv_specials_device_path(fast, oracle):
    rng = np.random.default_rng(8)
    m, n = 260, 333
    A = qgen.matrix(rng, m, n); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); nan = np.array([1, 0x7FFF << 48], dtype=np.uint64)
    A[5 * n + 7] = inf; A[9 * n + 1] = nan; A[11 * n: 12 * n] = 0; x[100] = 0
    alpha, beta = quad.random_quads(rng, 2)
    for layout in "RC":
        mm, nn = (m, n) if layout == "R" else (n, m)
        xx = x[:nn] if nn <= n else np.concatenate([x, quad.random_quads(rng, nn - n)])
        yy = y0[:mm] if mm <= m else np.concatenate([y0, quad.random_quads(rng, mm - m)])
        dy = to_dev(yy)
        fast.gemv(layout, mm, nn, alpha, to_dev(A), n, to_dev(xx), 1, beta, dy, 1)
        got = to_host(dy)
        fx = _fr_vec(xx)
        for i in range(mm):
            row = [A[i * n + j] if layout == "R" else A[j * n + i] for j in range(nn)]
            cls = [int(q[1]) >> 48 & 0x7FFF for q in row]
            if 0x7FFF in cls:
                r = got[i]
                nonfinite = (int(r[1]) >> 48 & 0x7FFF) == 0x7FFF
                assert nonfinite, (layout, i)
                continue
            S = sum(_frac(q) * f for q, f in zip(row, fx))
            want = oracle.fma(alpha, _round(S), oracle.mul(beta, yy[i]))
            assert quad.same_bits(got[i], want).all(), (layout, i)


def test_fast_gemv_full_size_sampled(fast, oracle):
    """BASELINE config 2 scale (8192^2, both layouts) on device; 24 sampled rows against exact sums."""
    n = 8192
    A = dev_random((n * n,), "D113", 5); x = dev_random((n,), "D113", 6); y = dev_random((n,), "D113", 7)
    yin = to_host(y); xh = to_host(x); fx = _fr_vec(xh)
    one = oracle.from_double(1.0); zero = oracle.from_double(0.0)
    for layout in "RC":
        y1 = y.clone(); y2 = y.clone()
        fast.gemv(layout, n, n, 1.0, A, n, x, 1, 0.0, y1, 1)
        fast.gemv(layout, n, n, 1.0, A, n, x, 1, 0.0, y2, 1)
        torch.cuda.synchronize()
        assert torch.equal(y1, y2)
        got = to_host(y1)
        for r in np.random.default_rng(2).integers(0, n, 24):
            Ar = to_host(A[r * n:(r + 1) * n]) if layout == "R" else to_host(A.view(n, n, 2)[:, r].contiguous())
            S = sum(a * b for a, b in zip(_fr_vec(Ar), fx))
            want = oracle.fma(one, _round(S), oracle.mul(zero, yin[r]))
            assert quad.same_bits(got[r], want).all(), (layout, r)


def test_fast_variant_zero_still_available(qb, oracle):
    """the rounded-FMA-chain generation of the fast dot stays selectable and inside the contract"""
    rng = np.random.default_rng(4)
    n = 20000
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    qb.set_mode(qb.MODE_FAST); qb.set_fast_variant(0)
    try:
        r0 = qb.dot(n, x, 1, y, 1)
        qb.set_fast_variant(1)
        r1 = qb.dot(n, x, 1, y, 1)
    finally:
        qb.set_fast_variant(2); qb.set_mode(qb.MODE_REFERENCE)
    fx, fy = _fr_vec(x), _fr_vec(y)
    tot = sum(a * b for a, b in zip(fx, fy)); sab = sum(abs(a * b) for a, b in zip(fx, fy))
    u = Fraction(1, 2 ** 113)
    assert abs(_frac(r0) - tot) <= n * u * sab and abs(_frac(r1) - tot) <= abs(tot) * u + n * sab / 2 ** 133


# ---------------------------------------------------------------------------------------------------------------------------
# the copy-engine-fed window kernel (k_dot_wide_tma): contiguous vectors from 2^19 elements

@pytest.mark.parametrize("kind,n", [("D113", (1 << 19) + 3), ("Dexp", 700001), ("D53", 1 << 21), ("D113", 2 * 512 * 444 + 511)])
def test_fast_dot_tiles_against_the_long_accumulator(fast, oracle, kind, n):
    """tiles of 512 pairs by cp.async.bulk, the ragged tail by plain loads: the exact sum rounded once (up to the window truncation,
    2^-133 of sum|x||y| per term), deterministic, and x . x (one set of tiles) as well"""
    rng = np.random.default_rng(n)
    x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, "D53" if kind == "D53" else "D113")
    x[::13] = 0; y[5::17] = 0
    none = np.array([[0, 0]], dtype=np.int64)
    got = fast.dot(n, x, 1, y, 1)
    assert quad.same_bits(got, fast.dot(n, x, 1, y, 1)).all()
    exact, ratio, _ = oracle.exact_dot_check("R", n, x, n, y, 1, none, got.reshape(1, 2))
    assert ratio[0] <= 1.0 / n + 2.0 ** -20 + 1e-12
    if kind != "Dexp":
        assert quad.same_bits(got, exact[0]).all()
    got2 = fast.dot(n, x, 1, x, 1)
    exact2, ratio2, _ = oracle.exact_dot_check("R", n, x, n, x, 1, none, got2.reshape(1, 2))
    assert ratio2[0] <= 1.0 / n + 2.0 ** -20 + 1e-12
    if kind != "Dexp":
        assert quad.same_bits(got2, exact2[0]).all()
    # device vectors that start 16 bytes into their allocations
    dx, dy = to_dev(x), to_dev(y)
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    fast.dot(n - 1, dx[1:], 1, dy[1:], 1, out)
    got3 = to_host(out.view(1, 2))
    _, ratio3, _ = oracle.exact_dot_check("R", n - 1, x[1:], n - 1, y[1:], 1, none, got3.reshape(1, 2))
    assert ratio3[0] <= 1.0 / n + 2.0 ** -20 + 1e-12


def test_fast_dot_tiles_specials_and_cancellation(fast):
    rng = np.random.default_rng(3)
    n = (1 << 20) + 77
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); ninf = np.array([0, 0xFFFF << 48], dtype=np.uint64)
    nan = np.array([1, 0x7FFF << 48], dtype=np.uint64); one = quad.from_double(np.array([1.0]))[0]
    xs, ys = x.copy(), y.copy(); xs[777777] = inf; ys[777777] = one
    assert quad.same_bits(fast.dot(n, xs, 1, ys, 1), inf).all()
    xs[400] = ninf; ys[400] = one
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    xs, ys = x.copy(), y.copy(); xs[n - 2] = nan                                   # in the ragged tail
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    xs, ys = x.copy(), y.copy(); xs[12] = inf; ys[12] = 0
    assert quad.is_nan(fast.dot(n, xs, 1, ys, 1).reshape(1, 2)).all()
    assert ((int(fast.nrm2(n, np.concatenate([x[:5], inf[None, :], x[6:]]), 1)[1]) >> 48) & 0x7fff) == 0x7fff
    # subnormals and a product far above everything seen so far (the out-of-line exact path) inside a tile
    xs = np.zeros((n, 2), dtype=np.uint64); ys = quad.from_double(np.ones(n))
    xs[:] = quad.from_double(np.array([2.0 ** -60]))[0]
    xs[1000] = quad.from_double(np.array([1e20]))[0]; xs[900000] = quad.from_double(np.array([-1e20]))[0]
    xs[5] = (np.uint64(99), np.uint64(0))                                           # 99 * 2^-16494
    tot = Fraction(n - 3, 2 ** 60) + Fraction(99, 2 ** 16494)
    assert quad.same_bits(fast.dot(n, xs, 1, ys, 1), _round(tot)).all()
    xv = np.zeros(n); xv[[3, 500000, n - 1]] = [1e20, 1.0, -1e20]
    assert _frac(fast.dot(n, quad.from_double(xv), 1, quad.from_double(np.ones(n)), 1)) == 1

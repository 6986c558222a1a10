"""GPU checks of the fast-mode qgemv on the FP64 pipe (k_gemv_f64, csrc/qslice.cuh; both layouts) through the C ABI.

The checker is the oracle's long accumulator (oracle/qoracle.c: exact sums rounded once, and err / (n u sum|a||x|) per row): every
row must keep the fast-mode contract (ratio <= 1) — accepted rows in fact the tighter bound of the qslice.cuh header, under which
the result is the exact sum rounded once except when that sum lies within 2^-14 ulp-ish of a rounding boundary — and the rows the
kernel declines must carry exactly the bits of the window kernel (fast variant 1), which recomputes them."""
import numpy as np
import pytest
import torch

import qgen
from gpu_util import dev_random, to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast(qb):
    qb.set_mode(qb.MODE_FAST)
    qb.set_fast_variant(2)
    yield qb
    qb.set_fast_variant(2)
    qb.set_mode(qb.MODE_REFERENCE)


def _gemv(qb, m, n, A, lda, x, y0, alpha=1.0, beta=0.0, incx=1, incy=1, variant=2, layout="R"):
    qb.set_fast_variant(variant)
    try:
        dy = to_dev(y0)
        qb.gemv(layout, m, n, alpha, to_dev(A), lda, to_dev(x), incx, beta, dy, incy)
        declined = qb.gemv_last_declined()
        return to_host(dy), declined
    finally:
        qb.set_fast_variant(2)


def _rows_check(oracle, m, n, A, lda, x, got, layout="R"):
    idx = np.stack([np.arange(m), np.zeros(m, dtype=np.int64)], axis=1)
    exact, ratio, klass = oracle.exact_dot_check(layout, n, A, lda, x, 1 if layout == "R" else n, idx, got)
    return exact, ratio, klass


@pytest.mark.parametrize("layout", ["R", "C"])
@pytest.mark.parametrize("kind,m,n,pad", [("D113", 1500, 1100, 3), ("D53", 1501, 1027, 0), ("Dexp", 2000, 700, 5), ("D113", 640, 4096, 0)])
def test_sliced_gemv_every_row_against_the_long_accumulator(fast, oracle, layout, kind, m, n, pad):
    rng = np.random.default_rng(m + n)
    lda = (n if layout == "R" else m) + pad
    A = qgen.matrix(rng, m if layout == "R" else n, n if layout == "R" else m, kind, lda)
    x = quad.random_quads(rng, n, "D53" if kind == "D53" else "D113"); y0 = quad.random_quads(rng, m)
    got, declined = _gemv(fast, m, n, A, lda, x, y0, layout=layout)
    assert declined >= 0, "the call did not take the sliced path"
    exact, ratio, klass = _rows_check(oracle, m, n, A, lda, x, got, layout)
    assert (klass == 0).all()
    assert ratio.max() <= 1.0, ratio.max()                       # the contract, every row
    same = quad.same_bits(got, exact)
    assert same.mean() > 0.995, same.mean()                      # ... and nearly always the exact sum rounded once
    assert ratio.max() <= 2.0 / n + 1e-6                         # one rounding of a sum is at most ~1/n of n u sum|a||x|
    if kind != "Dexp":
        assert declined == 0
    got2, _ = _gemv(fast, m, n, A, lda, x, y0, layout=layout)
    assert quad.same_bits(got, got2).all()                       # deterministic


@pytest.mark.parametrize("layout", ["R", "C"])
def test_sliced_gemv_declined_rows_are_the_window_kernels(fast, oracle, layout):
    """exponents spread over 2^+-100 on both sides: the kernel declines the rows where no product comes within 2^-12 of
    (largest |a_ij|) x (largest |x_j|) — predicted here from the data — and those rows carry the window kernel's bits"""
    rng = np.random.default_rng(3)
    m, n = 1200, 1024
    A = quad.random_quads(rng, m * n, emin=-100, emax=100); x = quad.random_quads(rng, n, emin=-100, emax=100); y0 = quad.random_quads(rng, m)
    lda = n if layout == "R" else m
    got, declined = _gemv(fast, m, n, A, lda, x, y0, layout=layout)
    ref, none = _gemv(fast, m, n, A, lda, x, y0, variant=1, layout=layout)
    assert none == -1
    ea = ((A[:, 1] >> np.uint64(48)).astype(np.int64) & 0x7fff).reshape((m, n) if layout == "R" else (n, m)); ex = (x[:, 1] >> np.uint64(48)).astype(np.int64) & 0x7fff
    ea = ea if layout == "R" else ea.T
    want_declined = (ea + ex[None, :]).max(axis=1) < ea.max(axis=1) + ex.max() - 12
    assert declined == int(want_declined.sum()) and 0 < declined < m, (declined, int(want_declined.sum()))
    assert quad.same_bits(got[want_declined], ref[want_declined]).all()
    exact, ratio, klass = _rows_check(oracle, m, n, A, lda, x, got, layout)
    assert ratio.max() <= 1.0


def test_sliced_gemv_zeros_subnormals_specials(fast, oracle):
    rng = np.random.default_rng(4)
    m, n = 800, 1400
    A = qgen.matrix(rng, m, n, "D113", n); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); nan = np.array([1, 0x7FFF << 48], dtype=np.uint64); sub = np.array([5, 0], dtype=np.uint64)
    A[::5] = 0; x[::11] = 0                       # zeros are skipped
    A[3 * n: 4 * n] = 0                           # an all-zero row: exactly +0
    A[7 * n + 100] = inf; A[9 * n + 1] = nan; A[13 * n + 700] = sub; A[15 * n + 5] = inf; A[15 * n + 6] = inf ^ np.array([0, 1 << 63], dtype=np.uint64)
    got, declined = _gemv(fast, m, n, A, n, x, y0)
    ref, _ = _gemv(fast, m, n, A, n, x, y0, variant=1)
    assert declined == 5                          # the four rows with an Inf / NaN / subnormal and the all-zero row
    for i in (3, 7, 9, 13, 15):
        assert quad.same_bits(got[i], ref[i]).all()
    assert (got[3] == 0).all()
    exact, ratio, klass = _rows_check(oracle, m, n, A, n, x, got)
    fin = klass == 0
    assert ratio[fin].max() <= 1.0 and fin.sum() == m - 3
    nonfin = ~fin
    assert ((got[nonfin, 1] >> np.uint64(48)) & np.uint64(0x7fff) == np.uint64(0x7fff)).all()
    # an Inf in x: every row goes to the window kernel, bit for bit
    x2 = x.copy(); x2[17] = inf
    got2, declined2 = _gemv(fast, m, n, A, n, x2, y0)
    ref2, _ = _gemv(fast, m, n, A, n, x2, y0, variant=1)
    assert declined2 == m and quad.same_bits(got2, ref2).all()
    # all of x zero: the same
    x3 = np.zeros_like(x)
    got3, declined3 = _gemv(fast, m, n, A, n, x3, y0)
    ref3, _ = _gemv(fast, m, n, A, n, x3, y0, variant=1)
    assert quad.same_bits(got3, ref3).all()


def test_sliced_gemv_strides_and_epilogue(fast, oracle):
    """incx / incy / lda and y = fma(alpha, S, mul(beta, y)) (level2.hpp:48): S from the long accumulator through the oracle's scalar ops"""
    rng = np.random.default_rng(5)
    m, n, lda, incx, incy = 1300, 900, 911, 2, 3
    A = qgen.matrix(rng, m, n, "D113", lda); x = quad.random_quads(rng, n * incx); y0 = quad.random_quads(rng, m * incy)
    alpha, beta = quad.random_quads(rng, 2)
    got, declined = _gemv(fast, m, n, A, lda, x, y0, alpha, beta, incx, incy)
    assert declined == 0
    xs = np.ascontiguousarray(x[::incx][:n])
    idx = np.stack([np.arange(m), np.zeros(m, dtype=np.int64)], axis=1)
    exact, _, _ = oracle.exact_dot_check("R", n, A, lda, xs, 1, idx)
    want = np.stack([oracle.fma(alpha, exact[i], oracle.mul(beta, y0[i * incy])) for i in range(m)])
    same = quad.same_bits(got[::incy][:m], want)
    assert same.mean() > 0.995
    mask = np.ones(len(y0), dtype=bool); mask[::incy] = False
    assert quad.same_bits(got[mask], y0[mask]).all()             # the gaps of a strided y are untouched


@pytest.mark.parametrize("layout", ["R", "C"])
def test_sliced_gemv_full_size(fast, oracle, layout):
    """BASELINE config 2 scale (8192^2) on device: sampled rows against the long accumulator, deterministic"""
    n = 8192
    A = dev_random((n * n,), "D113", 5); x = dev_random((n,), "D113", 6); y = dev_random((n,), "D113", 7)
    y1 = y.clone(); y2 = y.clone()
    fast.gemv(layout, n, n, 1.0, A, n, x, 1, 0.0, y1, 1)
    assert fast.gemv_last_declined() == 0
    fast.gemv(layout, n, n, 1.0, A, n, x, 1, 0.0, y2, 1)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    got = to_host(y1); xh = to_host(x)
    rows = np.random.default_rng(2).integers(0, n, 64)
    Ar = np.concatenate([to_host(A[r * n:(r + 1) * n]) if layout == "R" else to_host(A.view(n, n, 2)[:, r].contiguous()) for r in rows])
    idx = np.stack([np.arange(len(rows)), np.zeros(len(rows), dtype=np.int64)], axis=1)
    exact, ratio, klass = oracle.exact_dot_check("R", n, Ar, n, xh, 1, idx, got[rows])
    assert ratio.max() <= 2.0 / n + 1e-6
    assert quad.same_bits(got[rows], exact).mean() > 0.98


# ---------------------------------------------------------------------------------------------------------------------------
# qnrm2 / qdot(x, x): the sum of squares on the FP64 pipe (k_sumsq_f64)

@pytest.fixture()
def fast3(qb):
    """the sliced kernels without their size thresholds"""
    qb.set_mode(qb.MODE_FAST)
    qb.set_fast_variant(3)
    yield qb
    qb.set_fast_variant(2)
    qb.set_mode(qb.MODE_REFERENCE)


@pytest.mark.parametrize("kind,n,incx", [("D113", 300000, 1), ("D53", 1 << 20, 1), ("Dexp", 400001, 1), ("wide", 300000, 2), ("D113", 1000, 1), ("D113", 7, 3)])
def test_sliced_sum_of_squares(fast3, oracle, kind, n, incx):
    fast = fast3
    rng = np.random.default_rng(n)
    x = quad.random_quads(rng, (n - 1) * incx + 1, emin=-150, emax=150) if kind == "wide" else quad.random_quads(rng, (n - 1) * incx + 1, kind)
    xs = np.ascontiguousarray(x[::incx][:n])
    exact, _, _ = oracle.exact_dot_check("R", n, xs, n, xs, 1, np.array([[0, 0]], dtype=np.int64))
    got = fast.dot(n, x, incx, x, incx)
    assert quad.same_bits(got, fast.dot(n, x, incx, x, incx)).all()           # deterministic
    _, ratio, _ = oracle.exact_dot_check("R", n, xs, n, xs, 1, np.array([[0, 0]], dtype=np.int64), got.reshape(1, 2))
    assert ratio[0] <= 2.0 / n + 1e-9                                            # one rounding (1 / n of the contract) plus n 2^-125
    if kind != "wide":
        assert quad.same_bits(got, exact[0]).all()
    nr = fast.nrm2(n, x, incx)
    fast.set_fast_variant(1)
    try:
        nr1 = fast.nrm2(n, x, incx)                                              # the window kernel: sqrt of the same rounded sum
    finally:
        fast.set_fast_variant(3)
    if kind != "wide":
        assert quad.same_bits(nr, nr1).all()


def test_sliced_sum_of_squares_declines_specials(fast3):
    fast = fast3
    rng = np.random.default_rng(11)
    n = 300000
    x = quad.random_quads(rng, n)
    x[::7] = 0
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); sub = np.array([77, 0], dtype=np.uint64)
    for special in (inf, sub):
        y = x.copy(); y[123457] = special
        got = fast.nrm2(n, y, 1)
        fast.set_fast_variant(1)
        try:
            want = fast.nrm2(n, y, 1)
        finally:
            fast.set_fast_variant(3)
        assert quad.same_bits(got, want).all()
    assert ((int(fast.nrm2(n, np.concatenate([x[:5], inf[None, :], x[6:]]), 1)[1]) >> 48) & 0x7fff) == 0x7fff
    z = np.zeros_like(x)
    r = fast.nrm2(n, z, 1)
    assert int(r[0]) == 0 and int(r[1]) == 0


@pytest.mark.parametrize("layout", ["R", "C"])
def test_row_blocks_planned_as_the_whole_call_keep_its_bits(fast, layout):
    """qb_gemv_rows_dev: a block of rows computed with m_total = the whole qgemv (slabs of the pipelined host path, row blocks of the
    multi-GPU qgemv) takes the same kernel and column splits, so every y_i is bit-identical to the unsplit call"""
    rng = np.random.default_rng(21)
    m, n = 4100, 2048
    lda = n if layout == "R" else m
    A = qgen.matrix(rng, m if layout == "R" else n, n if layout == "R" else m, "D113", lda); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    alpha, beta = quad.random_quads(rng, 2)
    dA, dx = to_dev(A), to_dev(x)
    dy = to_dev(y0)
    fast.gemv(layout, m, n, alpha, dA, lda, dx, 1, beta, dy, 1)
    whole = to_host(dy)
    dy2 = to_dev(y0)
    for lo, hi in ((0, 700), (700, 1900), (1900, 4100)):          # the first block alone would not even take the sliced kernel
        blk = dA[lo * lda:] if layout == "R" else dA[lo:]
        fast.gemv(layout, hi - lo, n, alpha, blk, lda, dx, 1, beta, dy2[lo:hi], 1, m_total=m)
    assert quad.same_bits(to_host(dy2), whole).all()


def test_sliced_sum_of_squares_default_threshold(fast, oracle):
    """the default (variant 2) takes the sliced kernel from 2^24 elements: one such call against the long accumulator"""
    n = (1 << 24) + 5
    x = dev_random((n,), "D113", 9)
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    fast.dot(n, x, 1, x, 1, out)
    got = to_host(out.view(1, 2))
    xh = to_host(x)
    _, ratio, _ = oracle.exact_dot_check("R", n, xh, n, xh, 1, np.array([[0, 0]], dtype=np.int64), got.reshape(1, 2))
    assert ratio[0] <= 2.0 / n + 1e-9


@pytest.mark.parametrize("layout,m,n", [("R", 37, 200), ("C", 300, 128), ("R", 129, 1031)])
def test_sliced_gemv_small_shapes(fast3, oracle, layout, m, n):
    """variant 3: the sliced kernel on shapes below its size threshold (ragged tiles, fewer rows than a CTA, a single column split)"""
    rng = np.random.default_rng(m * n)
    lda = (n if layout == "R" else m) + 2
    A = qgen.matrix(rng, m if layout == "R" else n, n if layout == "R" else m, "D113", lda); x = quad.random_quads(rng, n); y0 = quad.random_quads(rng, m)
    dy = to_dev(y0)
    fast3.gemv(layout, m, n, 1.0, to_dev(A), lda, to_dev(x), 1, 0.0, dy, 1)
    assert fast3.gemv_last_declined() == 0
    got = to_host(dy)
    exact, ratio, klass = _rows_check(oracle, m, n, A, lda, x, got, layout)
    assert ratio.max() <= 2.0 / n + 1e-6 and quad.same_bits(got, exact).mean() > 0.99


def test_sliced_gemv_random_shapes(fast3, oracle):
    """a sweep of random shapes, paddings, strides and layouts through the sliced kernel (variant 3: no size threshold): ragged tiles and
    splits, fewer rows than a CTA, sprinkled zeros; every row against the long accumulator"""
    rng = np.random.default_rng(2024)
    for trial in range(28):
        layout = "RC"[trial % 2]
        m = int(rng.integers(1, 700)); n = int(rng.integers(128, 3000))
        pad = int(rng.integers(0, 4)); incx = int(rng.integers(1, 3)); incy = int(rng.integers(1, 4))
        kind = ["D113", "D53", "Dexp"][trial % 3]
        lda = (n if layout == "R" else m) + pad
        A = qgen.matrix(rng, m if layout == "R" else n, n if layout == "R" else m, kind, lda)
        x = quad.random_quads(rng, (n - 1) * incx + 1, "D53" if kind == "D53" else "D113"); y0 = quad.random_quads(rng, (m - 1) * incy + 1)
        if trial % 4 == 0:
            A[rng.integers(0, len(A), len(A) // 7)] = 0; x[rng.integers(0, len(x), len(x) // 9)] = 0
        dy = to_dev(y0)
        fast3.gemv(layout, m, n, 1.0, to_dev(A), lda, to_dev(x), incx, 0.0, dy, incy)
        declined = fast3.gemv_last_declined()
        got = to_host(dy)
        xs = np.ascontiguousarray(x[::incx][:n])
        idx = np.stack([np.arange(m), np.zeros(m, dtype=np.int64)], axis=1)
        exact, ratio, klass = oracle.exact_dot_check(layout, n, A, lda, xs, 1 if layout == "R" else n, idx, np.ascontiguousarray(got[::incy][:m]))
        assert (klass == 0).all() and ratio.max() <= 1.0, (trial, layout, m, n, lda, incx, incy, kind, float(ratio.max()))
        assert quad.same_bits(got[::incy][:m], exact).mean() > 0.98, (trial, layout, m, n)
        assert declined >= 0 and (kind == "Dexp" or trial % 4 == 0 or declined == 0), (trial, declined)
        if incy > 1:
            mask = np.ones(len(y0), dtype=bool); mask[::incy] = False
            assert quad.same_bits(got[mask], y0[mask]).all()


# ---------------------------------------------------------------------------------------------------------------------------
# qdot of two vectors: both factors sliced on the fly (k_dot_f64_tma)

@pytest.mark.parametrize("kind,n", [("D113", 300000), ("D53", 1 << 20), ("Dexp", 400001), ("D113", 2000), ("D113", 5)])
def test_sliced_dot_of_two_vectors(fast3, oracle, kind, n):
    fast = fast3
    rng = np.random.default_rng(n + 1)
    x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, "D53" if kind == "D53" else "D113")
    x[::13] = 0; y[::17] = 0
    got = fast.dot(n, x, 1, y, 1)
    assert quad.same_bits(got, fast.dot(n, x, 1, y, 1)).all()                 # deterministic
    exact, ratio, _ = oracle.exact_dot_check("R", n, x, n, y, 1, np.array([[0, 0]], dtype=np.int64), got.reshape(1, 2))
    assert ratio[0] <= 1.0
    fast.set_fast_variant(1)
    try:
        ref = fast.dot(n, x, 1, y, 1)                                          # the window kernel
    finally:
        fast.set_fast_variant(3)
    _, ratio_ref, _ = oracle.exact_dot_check("R", n, x, n, y, 1, np.array([[0, 0]], dtype=np.int64), ref.reshape(1, 2))
    assert ratio[0] <= max(4.0 * ratio_ref[0], 4.0 / n)                        # as tight as the window accumulator up to its own rounding


def test_sliced_dot_of_two_vectors_declines(fast3):
    """an Inf / subnormal factor, or products that all lie far below (largest |x|) x (largest |y|): the window kernel's bits"""
    fast = fast3
    rng = np.random.default_rng(5)
    n = 300000
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); sub = np.array([77, 0], dtype=np.uint64)
    cases = []
    for special in (inf, sub):
        a = x.copy(); a[1234] = special; cases.append((a, y))
        b = y.copy(); b[4321] = special; cases.append((x, b))
    # the largest x meets a zero y and vice versa, everything else 2^-40 smaller: no product near the anchor pair
    a = x.copy(); b = y.copy()
    a[:, 1] = (a[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (np.uint64(16383 - 40) << np.uint64(48)); b[:, 1] = (b[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (np.uint64(16383 - 40) << np.uint64(48))
    a[7, 1] = np.uint64(16383 + 5) << np.uint64(48); b[7] = 0; b[9, 1] = np.uint64(16383 + 5) << np.uint64(48); a[9] = 0
    cases.append((a, b))
    for u, v in cases:
        got = fast.dot(n, u, 1, v, 1)
        fast.set_fast_variant(1)
        try:
            want = fast.dot(n, u, 1, v, 1)
        finally:
            fast.set_fast_variant(3)
        assert quad.same_bits(got, want).all()

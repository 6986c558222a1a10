"""CPU checks of the fast-mode window accumulator (qblas_b200/csrc/qwide.cuh, host/device dual source,
built with g++ through tests/host/qwide_host.cpp) against exact rational arithmetic: the result must be
the exact sum up to ONE rounding plus the stated window truncation (< 2^-133 of the largest product per
term) — far inside the fast-mode contract gamma_n * sum|x_i||y_i|."""
import ctypes as C
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from qblas_b200 import quad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def qw(tmp_path_factory):
    so = tmp_path_factory.mktemp("qwide") / "libqwide_host.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=gnu++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "host", "qwide_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
    lib.qwide_dot.argtypes = [i64, vp, i64, vp, i64, ci, ci, vp, vp]
    lib.qwide_tree.argtypes = [i64, vp, vp, vp]
    lib.qwide_blocksum.argtypes = [i64, vp, vp, ci, vp]
    return lib


def _dot(lib, x, y, lanes=1, variant=0, incx=1, incy=1, n=None):
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    n = len(x) if n is None else n
    out = np.zeros((1, 2), dtype=np.uint64); bad = np.zeros(1, dtype=np.uint32)
    lib.qwide_dot(n, x.ctypes.data, incx, y.ctypes.data, incy, lanes, variant, out.ctypes.data, bad.ctypes.data)
    return out[0], int(bad[0])


def _frac(q):
    return quad.to_fraction(int(q[1]), int(q[0]))


def _exact(x, y):
    tot = Fraction(0); mx = Fraction(0); sab = Fraction(0)
    for a, b in zip(x, y):
        p = _frac(a) * _frac(b)
        tot += p; sab += abs(p); mx = max(mx, abs(p))
    return tot, mx, sab


def _check(r, x, y):
    tot, mx, sab = _exact(x, y)
    err = abs(_frac(r) - tot)
    n = len(x)
    assert err <= abs(tot) / 2 ** 113 + n * mx / 2 ** 133, (float(err), float(tot), float(mx))
    u = Fraction(1, 2 ** 113)
    assert err <= n * u / (1 - n * u) * sab  # the fast-mode contract


@pytest.mark.parametrize("kind", ["D113", "Dexp", "D53"])
@pytest.mark.parametrize("lanes", [1, 3, 32])
def test_window_dot_vs_exact(qw, kind, lanes):
    rng = np.random.default_rng(len(kind) * 1000 + lanes * 7 + ord(kind[1]))
    n = 700
    x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, "D113")
    for variant in (0, 1, 2):
        r, bad = _dot(qw, x, y, lanes, variant)
        assert bad == 0
        _check(r, x, y)


@pytest.mark.parametrize("variant", [0, 2])
def test_window_rounds_exact_sum_once(qw, variant):
    """without cancellation the result is the correctly rounded exact sum (bitwise)"""
    rng = np.random.default_rng(5)
    n = 300
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    x[:, 1] &= np.uint64((1 << 63) - 1); y[:, 1] &= np.uint64((1 << 63) - 1)  # all positive
    r, _ = _dot(qw, x, y, 4, variant=variant)
    tot, _, _ = _exact(x, y)
    hi, lo = quad.from_fraction(tot)
    assert int(r[1]) == hi and int(r[0]) == lo


@pytest.mark.parametrize("variant", [0, 2])
def test_window_cancellation_and_wide_exponents(qw, variant):
    one = quad.from_double(np.array([1.0]))[0]
    xv = np.zeros(10); xv[:3] = [1e20, 1.0, -1e20]
    x = quad.from_double(xv); y = quad.from_double(np.ones(10))
    for lanes in (1, 2, 3, 7):
        r, _ = _dot(qw, x, y, lanes, variant=variant)
        assert quad.same_bits(r, one).all()   # test_quadblas.cpp:715-739 must be exactly 1
    # exponents spread over +-3000 binades, random signs, increasing and decreasing magnitude
    rng = np.random.default_rng(11)
    n = 400
    base = quad.random_quads(rng, n)
    for order in (1, -1):
        x = base.copy()
        e = (np.linspace(-3000, 3000, n)[::order]).astype(np.int64)
        ef = ((x[:, 1] >> np.uint64(48)) & np.uint64(0x7FFF)).astype(np.int64) + e
        x[:, 1] = (x[:, 1] & ~(np.uint64(0x7FFF) << np.uint64(48))) | (ef.astype(np.uint64) << np.uint64(48))
        y = quad.random_quads(rng, n)
        for lanes in (1, 5):
            r, _ = _dot(qw, x, y, lanes, variant=variant)
            _check(r, x, y)


@pytest.mark.parametrize("variant", [0, 2])
def test_window_zero_subnormal_and_nonfinite(qw, variant):
    rng = np.random.default_rng(3)
    n = 64
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    x[5] = 0; y[9] = 0; x[10, 1] = np.uint64(1) << np.uint64(63); x[10, 0] = 0   # +0, -0
    # subnormal operands against huge ones: products land in the normal range
    x[20] = (np.uint64(12345678901234567), np.uint64(0x0000_0000_1234_5678))
    y[20, 1] = (y[20, 1] & np.uint64(0x8000_FFFF_FFFF_FFFF)) | (np.uint64(0x7F00) << np.uint64(48))
    x[21] = (np.uint64(1), np.uint64(0)); y[21] = (np.uint64(3), np.uint64(0))      # subnormal * subnormal -> underflows
    r, bad = _dot(qw, x, y, 3, variant=variant)
    assert bad == 0
    _check(r, x, y)
    # all-zero input and n = 0 give +0
    z = np.zeros((8, 2), dtype=np.uint64)
    r, _ = _dot(qw, z, y[:8], variant=variant); assert int(r[0]) == 0 and int(r[1]) == 0
    r, _ = _dot(qw, z, z, n=0, variant=variant); assert int(r[0]) == 0 and int(r[1]) == 0
    # non-finite classes
    inf = np.array([0, 0x7FFF << 48], dtype=np.uint64); ninf = np.array([0, 0xFFFF << 48], dtype=np.uint64)
    nan = np.array([1, 0x7FFF << 48], dtype=np.uint64)
    one = quad.from_double(np.array([1.0]))[0]
    xs = x.copy(); ys = y.copy(); xs[3] = inf; ys[3] = one
    r, _ = _dot(qw, xs, ys, 2, variant=variant); assert quad.same_bits(r, inf).all()
    xs[4] = one; ys[4] = ninf
    r, _ = _dot(qw, xs, ys, 2, variant=variant); assert quad.is_nan(r.reshape(1, 2)).all()        # Inf - Inf
    xs = x.copy(); ys = y.copy(); xs[3] = ninf; ys[3] = one
    r, _ = _dot(qw, xs, ys, 2, variant=variant); assert quad.same_bits(r, ninf).all()
    xs[3] = inf; ys[3] = 0
    r, _ = _dot(qw, xs, ys, 2, variant=variant); assert quad.is_nan(r.reshape(1, 2)).all()        # Inf * 0
    xs = x.copy(); xs[7] = nan
    r, _ = _dot(qw, xs, y, 2, variant=variant); assert quad.is_nan(r.reshape(1, 2)).all()


def test_window_merge_tree(qw):
    rng = np.random.default_rng(17)
    for kind in ("D113", "Dexp"):
        n = 257
        x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, kind)
        out = np.zeros((1, 2), dtype=np.uint64)
        qw.qwide_tree(n, x.ctypes.data, y.ctypes.data, out.ctypes.data)
        _check(out[0], x, y)


@pytest.mark.parametrize("kind,n,lanes", [("D113", 257, 257), ("Dexp", 257, 257), ("D113", 3000, 128), ("Dexp", 5000, 444), ("wide", 600, 128),
                                          ("D53", 1, 128), ("D113", 5, 128), ("cancel", 400, 128)])
def test_window_aligned_block_sum(qw, kind, n, lanes):
    """the block reduction of the level-1 kernels: every window shifted once to the largest anchor, then one 224-bit integer sum
    (qw_sum_aligned): the exact sum rounded once up to the window truncation, whatever the spread of the anchors; empty windows
    (more lanes than elements) are zeros; a sum that cancels keeps the small term that is left"""
    rng = np.random.default_rng(n + lanes)
    if kind == "wide":
        x = quad.random_quads(rng, n, emin=-3000, emax=3000); y = quad.random_quads(rng, n, emin=-200, emax=200)
    elif kind == "cancel":
        x = quad.random_quads(rng, n, "D113"); y = quad.random_quads(rng, n, "D113")
        x[n // 2:] = x[:n // 2]; y[n // 2:] = y[:n // 2]; y[n // 2:, 1] ^= np.uint64(1 << 63)      # the second half cancels the first ...
        x[-1] = quad.from_double(np.array([3.0]))[0]; y[-1] = quad.from_double(np.array([2.0 ** -90]))[0]   # ... except one small term
        x[n // 2 - 1] = 0
    else:
        x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, "D53" if kind == "D53" else "D113")
    out = np.zeros((1, 2), dtype=np.uint64)
    qw.qwide_blocksum(n, x.ctypes.data, y.ctypes.data, lanes, out.ctypes.data)
    _check(out[0], x, y)
    if kind == "cancel":       # every lane floors once when it moves to the common anchor: at most `lanes` units of the window's last bit
        assert abs(_frac(out[0]) - Fraction(3, 2 ** 90)) <= Fraction(lanes, 2 ** 133)


@pytest.mark.parametrize("variant", [0, 2])
def test_window_overflow_and_underflow_round(qw, variant):
    """sum beyond the binary128 range rounds to Inf; tiny sums round through the subnormal range"""
    big = np.array([[0, 0x7FFE << 48]], dtype=np.uint64)            # 2^16383
    two = quad.from_double(np.array([2.0]))
    r, bad = _dot(qw, np.repeat(big, 2, 0), np.repeat(two, 2, 0), variant=variant)
    assert bad == 0 and int(r[1]) == 0x7FFF << 48 and int(r[0]) == 0
    tiny = np.array([[0, 0x0001 << 48]], dtype=np.uint64)           # 2^-16382
    half = quad.from_double(np.array([0.5]))
    r, _ = _dot(qw, tiny, half, variant=variant)                                      # 2^-16383: subnormal, exact
    assert int(r[1]) == 0x0000_8000_0000_0000 and int(r[0]) == 0

"""CPU checks of the sliced FP64 accumulate of the fast-mode qgemv (qblas_b200/csrc/qslice.cuh, host/device dual source, built with
g++ through tests/host/qwide_host.cpp) against exact rational arithmetic.  Every ACCEPTED row must satisfy the fast-mode contract
|s^ - s| <= gamma_n * sum |a_j||x_j| — in fact the much tighter bound the header derives — and the rows the
acceptance test rejects (the kernel recomputes those with the window accumulator) must be exactly the ones the header says."""
import ctypes as C
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from qblas_b200 import quad

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QBIAS = 16383


@pytest.fixture(scope="module")
def qs(tmp_path_factory):
    so = tmp_path_factory.mktemp("qslice") / "libqwide_host.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=gnu++17", "-shared", "-fPIC", "-o", str(so),
                    os.path.join(ROOT, "tests", "host", "qwide_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.qslice_dot.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.qslice_dot.restype = C.c_int
    lib.qslice_sumsq.argtypes = [C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
    lib.qslice_sumsq.restype = C.c_int
    lib.qslice_dot2.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.qslice_dot2.restype = C.c_int
    return lib


def _dot(lib, a, x, lanes):
    a = np.ascontiguousarray(a); x = np.ascontiguousarray(x)
    out = np.zeros((1, 2), dtype=np.uint64); info = np.zeros(4, dtype=np.int32)
    rc = lib.qslice_dot(len(a), a.ctypes.data, x.ctypes.data, lanes, out.ctypes.data, info.ctypes.data)
    return rc, out[0], info


def _frac(q):
    return quad.to_fraction(int(q[1]), int(q[0]))


def _exact(a, x):
    tot = Fraction(0); sab = Fraction(0)
    for p, q in zip(a, x):
        t = _frac(p) * _frac(q)
        tot += t; sab += abs(t)
    return tot, sab


def _check_contract(r, a, x, info):
    tot, sab = _exact(a, x)
    n = len(a)
    err = abs(_frac(r) - tot)
    u = Fraction(1, 2 ** 113)
    assert err <= n * u / (1 - n * u) * sab, (float(err), float(sab))
    # the header's own bound: one rounding + (n + 64) * 2^-127 * 2^(anc - QBIAS) * 2^(EX - QBIAS)
    anc, EX = int(info[0]), int(info[1])
    unit = Fraction(2) ** (anc + EX - 2 * QBIAS - 127)
    assert err <= abs(tot) * u + (n + 64) * unit, (float(err), float(tot), float(unit))


@pytest.mark.parametrize("kind", ["D113", "D53", "Dexp"])
@pytest.mark.parametrize("lanes", [1, 3, 32])
def test_sliced_dot_vs_exact(qs, kind, lanes):
    rng = np.random.default_rng(len(kind) * 100 + lanes)
    n = 700
    a = quad.random_quads(rng, n, kind); x = quad.random_quads(rng, n, "D113" if kind != "D53" else "D53")   # D53: the reference benchmark's data
    rc, r, info = _dot(qs, a, x, lanes)
    assert rc == 1, info
    _check_contract(r, a, x, info)


def test_sliced_dot_rounds_the_exact_sum_once(qs):
    """without cancellation the result is the correctly rounded exact sum (bitwise), as with the window accumulator"""
    rng = np.random.default_rng(5)
    n = 500
    a = quad.random_quads(rng, n); x = quad.random_quads(rng, n)
    a[:, 1] &= np.uint64((1 << 63) - 1); x[:, 1] &= np.uint64((1 << 63) - 1)
    rc, r, info = _dot(qs, a, x, 4)
    assert rc == 1
    tot, _ = _exact(a, x)
    hi, lo = quad.from_fraction(tot)
    assert (int(r[1]), int(r[0])) == (hi, lo)


def test_sliced_dot_growing_elements_move_the_anchor(qs):
    """every element larger than all before it: the window is re-anchored each time (by 1 .. 300 bits), nothing is lost"""
    rng = np.random.default_rng(6)
    n = 200
    a = quad.random_quads(rng, n); x = quad.random_quads(rng, n)
    e = (QBIAS - 3000 + np.cumsum(rng.integers(1, 30, n))).astype(np.uint64)
    e[50] += np.uint64(300); e[51:] += np.uint64(300)
    a[:, 1] = (a[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (e << np.uint64(48))
    x[:, 1] = (x[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (np.uint64(QBIAS) << np.uint64(48))
    for lanes in (1, 7):
        rc, r, info = _dot(qs, a, x, lanes)
        assert rc == 1 and int(info[0]) == int(e[-1])
        _check_contract(r, a, x, info)


def test_sliced_dot_wide_exponents_accepted_or_rejected_correctly(qs):
    """exponents spread over 200 bits on both sides: accepted rows keep the contract; a row is rejected exactly when no product
    comes within 2^-12 of (largest |a|) x (largest |x|)"""
    rng = np.random.default_rng(7)
    seen = set()
    for trial in range(40):
        n = 256
        a = quad.random_quads(rng, n, emin=-100, emax=100); x = quad.random_quads(rng, n, emin=-100, emax=100)
        rc, r, info = _dot(qs, a, x, 8)
        ea = (a[:, 1] >> np.uint64(48)).astype(np.int64) & 0x7fff; ex = (x[:, 1] >> np.uint64(48)).astype(np.int64) & 0x7fff
        want = int((ea + ex).max() >= ea.max() + ex.max() - 12)
        assert rc == want, (trial, info)
        seen.add(rc)
        if rc == 1:
            _check_contract(r, a, x, info)
    assert seen == {0, 1} or seen == {0}


def test_sliced_dot_zeros_subnormals_specials(qs):
    rng = np.random.default_rng(8)
    n = 300
    a = quad.random_quads(rng, n); x = quad.random_quads(rng, n)
    a[::3] = 0                                      # zeros in the row are skipped
    a[5, 1] = np.uint64(1 << 63); a[5, 0] = 0       # -0
    x[::7] = 0
    rc, r, info = _dot(qs, a, x, 5)
    assert rc == 1
    _check_contract(r, a, x, info)
    b = a.copy(); b[10, 1] = np.uint64(0x0000_0000_0000_0001); b[10, 0] = 0     # a subnormal in the row: recomputed elsewhere
    assert _dot(qs, b, x, 5)[0] == 0
    b = a.copy(); b[10, 1] = np.uint64(0x7fff) << np.uint64(48)                  # Inf in the row
    assert _dot(qs, b, x, 5)[0] == 0
    y = x.copy(); y[11, 1] = np.uint64(0x7fff) << np.uint64(48)                  # Inf in x: the whole call goes to the window kernel
    assert _dot(qs, a, y, 5)[0] == -1
    z = np.zeros_like(a)
    rc, r, info = _dot(qs, z, x, 5)                 # an all-zero row: +0 here, and left to the window kernel (which gives the same)
    assert rc == 0 and int(r[0]) == 0 and int(r[1]) == 0


def test_sliced_dot_exact_cancellation(qs):
    """a term and its negative processed at the same anchor cancel exactly (the slices are the same, only the sign of the x copy
    differs); across a move of the anchor what remains is within the per-element bound"""
    rng = np.random.default_rng(9)
    n = 64
    a = quad.random_quads(rng, n); x = quad.random_quads(rng, n)
    a2 = np.repeat(a, 2, axis=0); x2 = np.repeat(x, 2, axis=0)
    x2[1::2, 1] ^= np.uint64(1 << 63)
    rc, r, info = _dot(qs, a2, x2, 1)
    assert rc == 1 and int(r[0]) == 0 and (int(r[1]) & ((1 << 63) - 1)) == 0
    a3 = np.concatenate([a, a]); x3 = np.concatenate([x, x])
    x3[n:, 1] ^= np.uint64(1 << 63)
    rc, r, info = _dot(qs, a3, x3, 4)
    assert rc == 1
    _check_contract(r, a3, x3, info)


def _sumsq(lib, x, lanes):
    x = np.ascontiguousarray(x)
    out = np.zeros((1, 2), dtype=np.uint64)
    rc = lib.qslice_sumsq(len(x), x.ctypes.data, lanes, out.ctypes.data)
    return rc, out[0]


@pytest.mark.parametrize("kind", ["D113", "D53", "Dexp", "wide"])
@pytest.mark.parametrize("lanes", [1, 5, 32])
def test_sliced_sum_of_squares_vs_exact(qs, kind, lanes):
    """qnrm2's sum of squares: the exact sum rounded once up to n 2^-125 of it (every term is positive: no acceptance test needed)"""
    rng = np.random.default_rng(len(kind) + lanes)
    n = 900
    x = quad.random_quads(rng, n, emin=-150, emax=150) if kind == "wide" else quad.random_quads(rng, n, kind)
    rc, r = _sumsq(qs, x, lanes)
    assert rc == 1
    tot = sum(_frac(q) ** 2 for q in x)
    err = abs(_frac(r) - tot)
    assert err <= tot / 2 ** 113 + n * tot / 2 ** 124
    if kind != "wide":
        hi, lo = quad.from_fraction(tot)
        assert (int(r[1]), int(r[0])) == (hi, lo)           # in fact the exact sum rounded once


def test_sliced_sum_of_squares_specials(qs):
    rng = np.random.default_rng(3)
    x = quad.random_quads(rng, 200)
    x[::4] = 0
    rc, r = _sumsq(qs, x, 3)
    assert rc == 1
    tot = sum(_frac(q) ** 2 for q in x)
    hi, lo = quad.from_fraction(tot)
    assert (int(r[1]), int(r[0])) == (hi, lo)
    y = x.copy(); y[7, 1] = np.uint64(0x7fff) << np.uint64(48)
    assert _sumsq(qs, y, 3)[0] == 0
    y = x.copy(); y[7, 1] = np.uint64(0); y[7, 0] = np.uint64(9)
    assert _sumsq(qs, y, 3)[0] == 0
    rc, r = _sumsq(qs, np.zeros_like(x), 3)
    assert rc == 1 and int(r[0]) == 0 and int(r[1]) == 0


def test_sliced_dot_extreme_exponents_match_the_window_accumulator(qs):
    """overflow to Inf, gradual underflow and the smallest normal results go through the same final rounding (qw_round) as the
    window accumulator: same bits on data without cancellation"""
    lib = qs
    lib.qwide_dot.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(12)
    n = 300
    for ea, ex in ((16000, 383), (16000, 300), (-8000, -8300), (-8000, -8382), (-8191, -8191), (0, 16383), (16383, 0)):
        a = quad.random_quads(rng, n, emin=ea - 3, emax=ea); x = quad.random_quads(rng, n, emin=ex - 3, emax=ex)
        a[:, 1] &= np.uint64((1 << 63) - 1); x[:, 1] &= np.uint64((1 << 63) - 1)        # positive: nothing cancels
        rc, r, info = _dot(qs, a, x, 4)
        assert rc == 1, (ea, ex, info)
        out = np.zeros((1, 2), dtype=np.uint64); bad = np.zeros(1, dtype=np.uint32)
        lib.qwide_dot(n, a.ctypes.data, 1, x.ctypes.data, 1, 4, 0, out.ctypes.data, bad.ctypes.data)
        assert (int(r[0]), int(r[1])) == (int(out[0][0]), int(out[0][1])), (ea, ex)


def _dot2(lib, x, y, lanes):
    x = np.ascontiguousarray(x); y = np.ascontiguousarray(y)
    out = np.zeros((1, 2), dtype=np.uint64)
    rc = lib.qslice_dot2(len(x), x.ctypes.data, y.ctypes.data, lanes, out.ctypes.data)
    return rc, out[0]


@pytest.mark.parametrize("kind", ["D113", "D53", "Dexp"])
@pytest.mark.parametrize("lanes", [1, 5, 32])
def test_sliced_dot_of_two_vectors_vs_exact(qs, kind, lanes):
    """qdot with both factors sliced on the fly: accepted results inside the contract (and the exact sum rounded once when nothing
    cancels badly); growing elements move both anchors"""
    rng = np.random.default_rng(len(kind) * 7 + lanes)
    n = 800
    x = quad.random_quads(rng, n, kind); y = quad.random_quads(rng, n, "D113" if kind != "D53" else "D53")
    rc, r = _dot2(qs, x, y, lanes)
    tot, sab = _exact(x, y)
    u = Fraction(1, 2 ** 113)
    if rc == 1:
        assert abs(_frac(r) - tot) <= n * u / (1 - n * u) * sab
        assert abs(_frac(r) - tot) <= abs(tot) * u + n * sab / 2 ** 120
    assert rc == 1 or kind == "Dexp"
    # positive data: the exact sum rounded once
    xp = x.copy(); yp = y.copy(); xp[:, 1] &= np.uint64((1 << 63) - 1); yp[:, 1] &= np.uint64((1 << 63) - 1)
    rc, r = _dot2(qs, xp, yp, lanes)
    if rc == 1 and kind != "Dexp":
        tot, _ = _exact(xp, yp)
        hi, lo = quad.from_fraction(tot)
        assert (int(r[1]), int(r[0])) == (hi, lo)


def test_sliced_dot_of_two_vectors_anchors_and_specials(qs):
    rng = np.random.default_rng(77)
    n = 300
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    ex = (QBIAS - 2000 + np.cumsum(rng.integers(0, 25, n))).astype(np.uint64); ey = (QBIAS + 1500 - np.cumsum(rng.integers(0, 25, n))).astype(np.uint64)
    x[:, 1] = (x[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (ex << np.uint64(48)); y[:, 1] = (y[:, 1] & np.uint64(0x8000FFFFFFFFFFFF)) | (ey << np.uint64(48))
    for lanes in (1, 6):
        rc, r = _dot2(qs, x, y, lanes)
        tot, sab = _exact(x, y)
        u = Fraction(1, 2 ** 113)
        if rc == 1:
            assert abs(_frac(r) - tot) <= n * u / (1 - n * u) * sab
    x = quad.random_quads(rng, n); y = quad.random_quads(rng, n)
    x[::5] = 0; y[::7] = 0
    rc, r = _dot2(qs, x, y, 3)
    assert rc == 1
    tot, sab = _exact(x, y)
    assert abs(_frac(r) - tot) <= abs(tot) / 2 ** 113 + n * sab / 2 ** 120
    z = x.copy(); z[9, 1] = np.uint64(0x7fff) << np.uint64(48)
    assert _dot2(qs, z, y, 3)[0] == 0 and _dot2(qs, y, z, 3)[0] == 0
    z = x.copy(); z[9, 1] = np.uint64(0); z[9, 0] = np.uint64(5)
    assert _dot2(qs, z, y, 3)[0] == 0

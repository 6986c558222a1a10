"""Host-side helpers for binary128 bit patterns held in numpy uint64 arrays of shape (..., 2)
([..., 0] = low 64 mantissa bits, [..., 1] = sign | exponent | high 48 mantissa bits), i.e. the
memory layout of Sleef_quad / __float128 on little-endian machines.  Pure integer code; no
arithmetic on quads happens here (that is the CUDA library's job).
"""
from fractions import Fraction

import numpy as np

BIAS = 16383
_M48 = (1 << 48) - 1


def empty(shape):
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    return np.zeros(shape + (2,), dtype=np.uint64)


def from_double(x):
    """Exact widening of float64 (any shape) -> quad bits (Sleef_cast_from_doubleq1 semantics)."""
    d = np.ascontiguousarray(x, dtype=np.float64)
    b = d.view(np.uint64)
    sign = b & np.uint64(1 << 63)
    e = ((b >> np.uint64(52)) & np.uint64(0x7FF)).astype(np.int64)
    m = b & np.uint64((1 << 52) - 1)
    out = np.zeros(d.shape + (2,), dtype=np.uint64)
    normal = (e > 0) & (e < 0x7FF)
    qe = np.where(normal, e - 1023 + BIAS, 0).astype(np.uint64)
    hi = sign | (qe << np.uint64(48)) | (m >> np.uint64(4))
    lo = m << np.uint64(60)
    infnan = e == 0x7FF
    hi = np.where(infnan, sign | np.uint64(0x7FFF << 48) | (m >> np.uint64(4)), hi)
    sub = (e == 0) & (m != 0)
    if np.any(sub):
        for idx in np.argwhere(sub):
            idx = tuple(idx)
            h, l = from_fraction(Fraction(float(d[idx])))
            hi[idx] = np.uint64(h)
            lo[idx] = np.uint64(l)
            if np.signbit(d[idx]):
                hi[idx] |= np.uint64(1 << 63)
    zero = (e == 0) & (m == 0)
    hi = np.where(zero, sign, hi)
    lo = np.where(zero | infnan & (m == 0), np.uint64(0), lo)
    out[..., 0] = lo
    out[..., 1] = hi
    return out


def from_fraction(fr):
    """Round-to-nearest-even conversion of an exact rational to quad bits -> (hi, lo) ints.
    Handles subnormals, overflow to Inf and signed results (sign of zero is +)."""
    fr = Fraction(fr)
    if fr == 0:
        return 0, 0
    sign = 1 if fr < 0 else 0
    a = -fr if sign else fr
    # find e with 2^e <= a < 2^(e+1)
    e = a.numerator.bit_length() - a.denominator.bit_length()
    if Fraction(2) ** e > a:
        e -= 1
    elif Fraction(2) ** (e + 1) <= a:
        e += 1
    emin = 1 - BIAS
    ee = max(e, emin)
    # scaled = a / 2^(ee-112) ; integer part is the mantissa (incl. implicit bit when normal)
    scaled = a / (Fraction(2) ** (ee - 112))
    m = scaled.numerator // scaled.denominator
    rem = scaled - m
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (m & 1)):
        m += 1
    if e >= emin:
        if m == (1 << 113):
            m >>= 1
            ee += 1
        if ee + BIAS >= 0x7FFF:
            return (sign << 63) | (0x7FFF << 48), 0
        bits = ((ee + BIAS) << 112) | (m - (1 << 112))
    else:
        bits = m  # subnormal (m may have reached 2^112 -> min normal, encoded naturally)
    bits |= sign << 127
    return bits >> 64, bits & ((1 << 64) - 1)


def to_fraction(hi, lo):
    """Exact rational value of a finite quad."""
    hi = int(hi)
    lo = int(lo)
    sign = -1 if hi >> 63 else 1
    e = (hi >> 48) & 0x7FFF
    m = ((hi & _M48) << 64) | lo
    if e == 0x7FFF:
        raise ValueError("inf/nan has no rational value")
    if e == 0:
        return sign * Fraction(m) * Fraction(2) ** (1 - BIAS - 112)
    return sign * Fraction(m | (1 << 112)) * Fraction(2) ** (e - BIAS - 112)


def is_nan(q):
    hi = q[..., 1] & np.uint64((1 << 63) - 1)
    return (hi > np.uint64(0x7FFF << 48)) | ((hi == np.uint64(0x7FFF << 48)) & (q[..., 0] != 0))


def same_bits(a, b):
    """Elementwise bitwise equality, with any NaN == any NaN (payloads are outside the contract)."""
    eq = (a[..., 0] == b[..., 0]) & (a[..., 1] == b[..., 1])
    return eq | (is_nan(a) & is_nan(b))


def to_float(q):
    """Lossy float64 view (for messages only)."""
    out = np.empty(q.shape[:-1], dtype=np.float64)
    for idx in np.ndindex(*q.shape[:-1]):
        hi, lo = int(q[idx + (1,)]), int(q[idx + (0,)])
        e = (hi >> 48) & 0x7FFF
        if e == 0x7FFF:
            out[idx] = float("nan") if (hi & _M48) | lo else (float("-inf") if hi >> 63 else float("inf"))
        else:
            try:
                out[idx] = float(to_fraction(hi, lo))
            except OverflowError:
                out[idx] = float("-inf") if hi >> 63 else float("inf")
    return out


def random_quads(rng, shape, kind="D113", emin=-8, emax=0):
    """Random quads.  kind:
       D53  : doubles U(-1,1) cast to quad (the reference's test/benchmark inputs,
              /root/reference/test_quadblas.cpp:88-111, benchmarks/benchmark.cpp:8-29)
       D113 : full 112-bit random mantissas, random sign, exponent uniform in [emin, emax]
       Dexp : D113 with exponent uniform in [-40, 40]
    """
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    if kind == "D53":
        return from_double(rng.uniform(-1.0, 1.0, size=shape))
    if kind == "Dexp":
        emin, emax = -40, 40
    out = np.zeros(shape + (2,), dtype=np.uint64)
    out[..., 0] = rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)
    mh = rng.integers(0, 1 << 48, size=shape, dtype=np.uint64)
    e = rng.integers(emin, emax + 1, size=shape).astype(np.int64) + BIAS
    s = rng.integers(0, 2, size=shape, dtype=np.uint64)
    out[..., 1] = (s << np.uint64(63)) | (e.astype(np.uint64) << np.uint64(48)) | mh
    return out

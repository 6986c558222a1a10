"""Pins the oracle's scalar arithmetic (and, through oracle/_ref, the shimmed reference's) against
exact rational arithmetic + a hand-written RNE: the `_u05` SLEEF functions the reference calls
 are correctly rounded, so the bits are unique."""
from fractions import Fraction

import numpy as np
import pytest

import qgen
from qblas_b200 import quad


def _finite(q):
    return ((int(q[1]) >> 48) & 0x7FFF) != 0x7FFF


def _exact_fma(a, b, c):
    fa, fb, fc = (quad.to_fraction(int(v[1]), int(v[0])) for v in (a, b, c))
    return fa * fb + fc


@pytest.mark.parametrize("regime", ["any", "similar", "cancel", "gaps", "tiny", "subres", "nearovf"])
def test_oracle_fma_exact(oracle, regime):
    rng = np.random.default_rng(hash(regime) & 0xFFFF)
    n = 250
    a, b, c = qgen.triples(rng, n, regime)
    got = oracle.fma(a, b, c)
    checked = 0
    for i in range(n):
        if not (_finite(a[i]) and _finite(b[i]) and _finite(c[i])):
            continue
        ex = _exact_fma(a[i], b[i], c[i])
        if ex == 0:
            continue  # sign-of-zero rules are covered by the specials test
        hi, lo = quad.from_fraction(ex)
        assert (int(got[i, 1]), int(got[i, 0])) == (hi, lo), f"{regime}[{i}]"
        checked += 1
    assert checked > n // 2


def test_oracle_add_mul_sqrt_exact(oracle):
    rng = np.random.default_rng(7)
    a, b, _ = qgen.triples(rng, 300, "similar")
    s, p = oracle.add(a, b), oracle.mul(a, b)
    for i in range(300):
        fa, fb = (quad.to_fraction(int(v[1]), int(v[0])) for v in (a[i], b[i]))
        if fa + fb != 0:
            assert (int(s[i, 1]), int(s[i, 0])) == quad.from_fraction(fa + fb)
        assert (int(p[i, 1]), int(p[i, 0])) == quad.from_fraction(fa * fb)
    # sqrt: r = RN(sqrt(x))  <=>  (r - ulp/2)^2 <= x <= (r + ulp/2)^2 (never a tie for sqrt)
    x = qgen.mk(np.zeros(200, dtype=np.uint64), qgen.BIAS - 30 + rng.integers(0, 60, size=200), *qgen.mantissas(rng, 200))
    r = oracle.sqrt(x)
    for i in range(200):
        fx = quad.to_fraction(int(x[i, 1]), int(x[i, 0]))
        fr = quad.to_fraction(int(r[i, 1]), int(r[i, 0]))
        e = ((int(r[i, 1]) >> 48) & 0x7FFF) - qgen.BIAS
        half = Fraction(2) ** (e - 113)
        assert (fr - half) ** 2 <= fx <= (fr + half) ** 2


def test_mul_add_are_fma_forms(oracle):
    """mul(a,b) == fma(a,b,-0) and add(a,b) == fma(a,1,b), signed zeros included.
    The product library relies on this identity for its epilogues."""
    rng = np.random.default_rng(11)
    for regime in ["similar", "specials", "tiny"]:
        a, b, _ = qgen.triples(rng, 2000, regime)
        negz = np.zeros_like(a); negz[:, 1] = np.uint64(1 << 63)
        one = np.zeros_like(a); one[:, 1] = np.uint64(0x3FFF << 48)
        assert quad.same_bits(oracle.mul(a, b), oracle.fma(a, b, negz)).all()
        assert quad.same_bits(oracle.add(a, b), oracle.fma(a, one, b)).all()


def test_ref_scalar_matches_oracle(oracle, ref):
    rng = np.random.default_rng(3)
    for regime in qgen.REGIMES:
        a, b, c = qgen.triples(rng, 3000, regime)
        assert quad.same_bits(ref.fma(a, b, c), oracle.fma(a, b, c)).all(), regime

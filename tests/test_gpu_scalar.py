"""GPU parity of the arithmetic core: every scalar op the reference takes from SLEEF
(Sleef_{fma,mul,add,sqrt}q1_u05, casts), bit for bit against the oracle, through the
C ABI (qb_elementwise_dev)."""
import numpy as np
import pytest
import torch

import qgen
from gpu_util import to_dev, to_host
from qblas_b200 import quad

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("regime", qgen.REGIMES)
@pytest.mark.parametrize("op", [0, 1])
def test_fma_bitexact(qb, oracle, regime, op):
    rng = np.random.default_rng(1000 + qgen.REGIMES.index(regime))
    n = 200_000
    a, b, c = qgen.triples(rng, n, regime)
    out = torch.empty((n, 2), dtype=torch.int64, device="cuda")
    qb.elementwise(op, to_dev(a), to_dev(b), to_dev(c), out)
    got = to_host(out)
    exp = oracle.fma(a, b, c)
    bad = ~quad.same_bits(got, exp)
    assert not bad.any(), f"{bad.sum()} mismatches, first at {np.argmax(bad)}: a={a[np.argmax(bad)]} b={b[np.argmax(bad)]} c={c[np.argmax(bad)]}"


@pytest.mark.parametrize("regime", ["similar", "cancel", "tiny", "specials", "overflow"])
def test_mul_add_sqrt_casts(qb, oracle, regime):
    rng = np.random.default_rng(77)
    n = 100_000
    a, b, _ = qgen.triples(rng, n, regime)
    out = torch.empty((n, 2), dtype=torch.int64, device="cuda")
    da, db = to_dev(a), to_dev(b)
    assert quad.same_bits(to_host(qb.elementwise(2, da, db, None, out)), oracle.mul(a, b)).all()
    assert quad.same_bits(to_host(qb.elementwise(3, da, db, None, out)), oracle.add(a, b)).all()
    assert quad.same_bits(to_host(qb.elementwise(4, da, None, None, out)), oracle.sqrt(a)).all()


def test_cast_roundtrip(qb, oracle):
    rng = np.random.default_rng(5)
    n = 50_000
    a = quad.random_quads(rng, n, "D113", -1100, 1030)
    out = torch.empty((n, 2), dtype=torch.int64, device="cuda")
    got = to_host(qb.elementwise(5, to_dev(a), None, None, out))
    exp = np.stack([oracle.from_double(oracle.to_double(v)) for v in a[:3000]])
    assert quad.same_bits(got[:3000], exp).all()
